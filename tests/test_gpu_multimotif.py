"""ZOOPS mixture of several motif models (BASELINE.json configs[3]: 3 motifs + homogeneous background).  No reference
counterpart (jstacs SingleHiddenMotifMixture has one motif component); the specification is DESIGN.md "multi-motif mixture",
restated in oracle/oracle.py::estep_multi on the oracle's own window / column scores.  K = 1 pins that restatement and the
kernel against the one-motif E-step."""
import numpy as np
import pytest

from util import orc, oracle_bg, oracle_fg, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    from phyfoo_b200 import core
    return core.default_context()


def _case(ctx, widths, orders, gap_rate=0.03):
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel
    newick = synth.random_binary_newick(7, 11)
    pwms = [synth.random_pwm(W, o, 40 + k) for k, (W, o) in enumerate(zip(widths, orders))]
    smp, _ = synth.make_sample(newick, 9, 50, 13, pwm=pwms[0], gap_rate=gap_rate, length_jitter=12)
    fgs = []
    for W, o, p in zip(widths, orders, pwms):
        m = PhyloBayesModel(W, o, newick, ctx)
        m.setCondProbs(p)
        fgs.append(m)
    bg = PhyloBackground(0, newick, ctx)
    ofgs = [oracle_fg(newick, p, order=o) for p, o in zip(pwms, orders)]
    return newick, smp, fgs, bg, ofgs, oracle_bg(newick, 0)


def test_one_component_equals_single_motif_estep(ctx):
    from phyfoo_b200 import core
    from phyfoo_b200.models import SingleHiddenMotifMixture
    newick, smp, fgs, bg, ofgs, obg = _case(ctx, [6], [0])
    ds = smp.device(ctx)
    res = core.zoops_estep_multi(ctx, ds, [fgs[0].device_model()], bg.device_model(), np.log([0.7, 0.3]))
    one = SingleHiddenMotifMixture(fgs[0], bg, zoops=0.7).getNewWeights(smp)
    assert rel_err(res["gamma"][0], one["gamma"]) < 1e-14 and abs(res["ll"] - one["ll"]) <= 1e-14 * abs(one["ll"])
    assert np.array_equal(res["argmax_off"], one["argmax_off"]) and np.allclose(res["w"], one["w"], rtol=1e-14, atol=0)
    ref = orc.estep(ofgs[0], obg, smp.col_offsets, smp.codes, np.log([0.7, 0.3]))
    spec = orc.estep_multi(ofgs, obg, smp.col_offsets, smp.codes, np.log([0.7, 0.3]))
    assert rel_err(spec["gamma"][0], ref["gamma"]) < 1e-12 and abs(spec["ll"] - ref["ll"]) <= 1e-12 * abs(ref["ll"])
    assert np.array_equal(spec["argmax_off"], ref["argmax_off"])


@pytest.mark.parametrize("widths,orders", [([6, 6, 6], [0, 0, 0]), ([4, 9, 6], [0, 1, 0])])
def test_three_motif_estep_matches_specification(ctx, widths, orders):
    from phyfoo_b200 import core
    newick, smp, fgs, bg, ofgs, obg = _case(ctx, widths, orders)
    ds = smp.device(ctx)
    logw = np.log([0.3, 0.25, 0.25, 0.2])
    aw = np.linspace(0.5, 1.5, smp.n_alignments)
    res = core.zoops_estep_multi(ctx, ds, [m.device_model() for m in fgs], bg.device_model(), logw, aw)
    ref = orc.estep_multi(ofgs, obg, smp.col_offsets, smp.codes, logw, aw)
    for k in range(3):
        assert np.max(np.abs(res["gamma"][k] - ref["gamma"][k])) < TOL
    assert abs(res["ll"] - ref["ll"]) <= TOL * abs(ref["ll"]) and np.allclose(res["w"], ref["w"], rtol=1e-9, atol=0)
    assert np.array_equal(res["argmax_motif"], ref["argmax_motif"]) and np.array_equal(res["argmax_off"], ref["argmax_off"])
    assert abs(res["w"].sum() - aw.sum()) <= 1e-9 * aw.sum()                  # the component posteriors of an alignment sum to 1
    # M-step data of component 1: the selector on its resident gamma = the selector on the gamma copied out
    sa, so, sw = core.select_multi(ctx, ds, 1)
    ra, ro, rw = core.select(ctx, ds, widths[1], res["gamma"][1])
    assert np.array_equal(sa, ra) and np.array_equal(so, ro) and np.array_equal(sw, rw)


def test_three_motif_em_recovers_planted_motifs(ctx):
    """EM of the 3-motif mixture on alignments with one of three planted motifs each: the log-likelihood increases and every
    planted site is called by the component that was trained towards it."""
    from phyfoo_b200 import synth
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import MultiMotifMixture, PhyloBackground, PhyloBayesModel
    newick = synth.caterpillar_newick(5)
    W = 7
    cons = [[0, 0, 1, 2, 3, 3, 1], [2, 1, 0, 3, 0, 2, 2], [3, 2, 2, 0, 1, 1, 0]]       # three consensus words that are not shifts of each other
    truth = [np.stack([np.eye(4)[c] * 0.88 + 0.03 for c in cons[k]])[:, None, :] for k in range(3)]
    parts, planted, kind = [], [], []
    for k in range(3):
        smp_k, pl = synth.make_sample(newick, 60, 60, 100 + k, pwm=[t for t in truth[k]])
        parts += [smp_k.getElementAt(i) for i in range(smp_k.n_alignments)]
        planted += list(pl); kind += [k] * smp_k.n_alignments
    smp = PhyloSample.from_alignments(parts)
    rng = np.random.default_rng(5)
    motifs = []
    for k in range(3):
        m = PhyloBayesModel(W, 0, newick, ctx)
        m.setCondProbs([0.6 * t + 0.1 for t in truth[k]])                    # start near its own motif
        motifs.append(m)
    mix = MultiMotifMixture(motifs, PhyloBackground(0, newick, ctx), weights=[0.3, 0.3, 0.3, 0.1])
    hist = mix.train(smp, max_iterations=6, rng=rng)
    assert hist[-1] > hist[0]
    res = mix.getNewWeights(smp)
    assert (res["argmax_motif"] == np.array(kind)).mean() > 0.85
    assert (res["argmax_off"] == np.array(planted)).mean() > 0.8
