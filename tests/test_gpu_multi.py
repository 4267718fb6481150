"""Several GPUs behind ONE caller (include/phyfoo.h phyfoo_multi_*): the library shards the sample, replicates the models and
all-reduces [ll, w0, w1] / [f, f_1..f_dim] over NCCL.  One call on N devices against the same call on one device: scores,
gamma, arg-max calls and the selected windows bit-identical, the all-reduced sums to 1e-12 relative.  Needs >= 2 GPUs
(`gpurun --gpus 2`); with one device the single-rank path of the same entry points is checked."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_devices():
    import torch
    return torch.cuda.device_count()


def _setup(order, n_aln=24, length=90, W=7):
    from phyfoo_b200 import _abi, synth
    from phyfoo_b200.newick import parse_newick
    newick = synth.random_binary_newick(8, 5)
    tree = parse_newick(newick)
    pwm = synth.random_pwm(W, order, 17)
    smp, _ = synth.make_sample(newick, n_aln, length, 23, pwm=pwm, gap_rate=0.01, length_jitter=25)
    return _abi, tree, pwm, smp, W


def _models(core, _abi, mctx, tree, pwm, W, order):
    fg = core.MultiModel(mctx, _abi.MODEL_FG, W, order, tree, np.tile(tree.branch, (W, 1)), pwm)
    bg = core.MultiModel(mctx, _abi.MODEL_BG, 0, 0, tree, np.tile(tree.branch, (1, 1)), [np.full((1, 4), 0.25)])
    return fg, bg


@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("n_dev", [1, 2, 4])
def test_multi_estep_and_gradient_equal_single_device(order, n_dev):
    from phyfoo_b200 import core
    if n_dev > _n_devices():
        pytest.skip(f"needs {n_dev} GPUs")
    _abi, tree, pwm, smp, W = _setup(order)
    logw = np.log([0.9, 0.1])
    res = {}
    for devs in ([0], list(range(n_dev))):
        mctx = core.MultiContext(devs)
        mds = core.MultiDataset(mctx, smp.col_offsets, smp.codes)
        fg, bg = _models(core, _abi, mctx, tree, pwm, W, order)
        e = core.multi_zoops_estep(mctx, mds, fg, bg, logw, strand_logw=np.log([0.5, 0.5]))
        fe = core.MultiFESet(mctx, mds, fg)                          # selects on the gamma left on each device
        pos = 1
        lam = np.log(pwm[pos].ravel()) + 0.03
        f, g = fe.grad(pos, lam)
        f2 = fe.eval(pos, lam)
        res[len(devs)] = (e, f, g, f2, len(fe), list(mds.bounds))
        fe.free(); fg.free(); bg.free(); mds.free(); mctx.close()
    (e1, f1, g1, f21, n1, _), (eN, fN, gN, f2N, nN, bounds) = res[1], res[n_dev]
    assert bounds[0] == 0 and bounds[-1] == smp.n_alignments and all(b > a for a, b in zip(bounds, bounds[1:]))
    assert np.array_equal(e1["gamma"], eN["gamma"])
    assert np.array_equal(e1["argmax_off"], eN["argmax_off"]) and np.array_equal(e1["argmax_strand"], eN["argmax_strand"])
    assert e1["stats"]["sweeps_fg"] == eN["stats"]["sweeps_fg"] and e1["stats"]["n_windows"] == eN["stats"]["n_windows"]
    assert abs(e1["ll"] - eN["ll"]) <= 1e-12 * abs(e1["ll"]) and np.allclose(e1["w"], eN["w"], rtol=1e-12, atol=0)
    assert n1 == nN
    assert abs(f1 - fN) <= 1e-12 * abs(f1) and abs(f21 - f2N) <= 1e-12 * abs(f21)
    assert np.all(np.abs(g1 - gN) <= 1e-9 * np.abs(g1) + 64 * np.finfo(float).eps * abs(f1) / 1e-4)


def test_multi_matches_single_gpu_api():
    """The multi entry points on one device against the plain single-GPU entry points (phyfoo_zoops_estep)."""
    from phyfoo_b200 import core
    _abi, tree, pwm, smp, W = _setup(0)
    logw = np.log([0.9, 0.1])
    mctx = core.MultiContext([0])
    mds = core.MultiDataset(mctx, smp.col_offsets, smp.codes)
    fg, bg = _models(core, _abi, mctx, tree, pwm, W, 0)
    e = core.multi_zoops_estep(mctx, mds, fg, bg, logw)
    ctx = core.default_context()
    sfg = core.DeviceModel(ctx, _abi.MODEL_FG, W, 0, tree, np.tile(tree.branch, (W, 1)), pwm)
    sbg = core.DeviceModel(ctx, _abi.MODEL_BG, 0, 0, tree, np.tile(tree.branch, (1, 1)), [np.full((1, 4), 0.25)])
    ref = core.zoops_estep(ctx, smp.device(ctx), sfg, sbg, logw)
    assert np.array_equal(e["gamma"], ref["gamma"]) and e["ll"] == ref["ll"] and np.array_equal(e["argmax_off"], ref["argmax_off"])


def test_multi_errors_do_not_hang():
    from phyfoo_b200 import core
    from phyfoo_b200.lib import PhyfooError
    n = min(_n_devices(), 2)
    _abi, tree, pwm, smp, W = _setup(0, n_aln=6, length=40)
    mctx = core.MultiContext(list(range(n)))
    with pytest.raises(PhyfooError):
        core.MultiDataset(mctx, smp.col_offsets[:1 + max(n - 1, 0)], smp.codes[: smp.col_offsets[max(n - 1, 0)]])   # fewer alignments than devices (or none)
    mds = core.MultiDataset(mctx, smp.col_offsets, smp.codes)
    wide = [np.full((1, 4), 0.25)] * 60
    fg = core.MultiModel(mctx, _abi.MODEL_FG, 60, 0, tree, np.tile(tree.branch, (60, 1)), wide)     # wider than the alignments
    bg = core.MultiModel(mctx, _abi.MODEL_BG, 0, 0, tree, np.tile(tree.branch, (1, 1)), [np.full((1, 4), 0.25)])
    with pytest.raises(PhyfooError):
        core.multi_zoops_estep(mctx, mds, fg, bg, np.log([0.9, 0.1]))
    with pytest.raises(PhyfooError):
        core.MultiContext([0, 0])
