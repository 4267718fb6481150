"""GPU parity at the EXACT shapes bench.py times (BASELINE.json configs[2] and configs[4]): the C3 tree
(12 species, seed 0x5EED0003) with an order-1 motif of width 12 on 1 kb alignments, and the C5 tree (30 species) with an
order-0 motif of width 12 -- a handful of alignments each, both strands, without gaps and with 5 % unobserved leaves.
Window scores against the oracle at 1e-9 relative, the total mean-field sweep count and the arg-max calls exact."""
import os

import numpy as np
import pytest

from util import orc, oracle_bg, oracle_fg, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    from phyfoo_b200 import core
    return core.default_context()


def _check(ctx, name, n_aln, length, gap_rate, order1_kernel=None):
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture, StrandModel
    cfg = dict(synth.CONFIGS[name])
    newick = synth.config_newick(cfg)
    pwm = synth.random_pwm(cfg["W"], cfg["order"], cfg["seed"] + 1)
    smp, planted = synth.make_sample(newick, n_aln, length, cfg["seed"] + 2, pwm=pwm, gap_rate=gap_rate)
    fg = PhyloBayesModel(cfg["W"], cfg["order"], newick, ctx)
    fg.setCondProbs(pwm)
    bg = PhyloBackground(0, newick, ctx)
    shm = SingleHiddenMotifMixture(StrandModel(fg, 0.5), bg, zoops=0.9)
    if order1_kernel is not None:
        ctx.set_option("order1_kernel", order1_kernel)
    try:
        res = shm.getNewWeights(smp, want_gamma=True)
        fwd, sw_f = fg.getLogProbFor(smp, strand=0, want_sweeps=True)
        rev, sw_r = fg.getLogProbFor(smp, strand=1, want_sweeps=True)
    finally:
        if order1_kernel is not None:
            ctx.set_option("order1_kernel", 2)
    ref = orc.estep(oracle_fg(newick, pwm, order=cfg["order"]), oracle_bg(newick, 0), smp.col_offsets, smp.codes,
                    np.log([0.9, 0.1]), np.log([0.5, 0.5]), threads=os.cpu_count() or 1)
    assert rel_err(fwd, ref["fwd"]) < TOL and rel_err(rev, ref["rev"]) < TOL
    assert int(sw_f.sum()) + int(sw_r.sum()) == ref["sweeps_fg"]
    assert res["stats"]["sweeps_fg"] == ref["sweeps_fg"]
    assert np.array_equal(res["argmax_off"], ref["argmax_off"]) and np.array_equal(res["argmax_strand"], ref["argmax_strand"])
    assert np.max(np.abs(res["gamma"] - ref["gamma"])) < TOL
    assert abs(res["ll"] - ref["ll"]) <= TOL * abs(ref["ll"])
    # the planted sites are where the arg-max calls land (sanity of the synthetic data, not a parity statement)
    assert np.mean(res["argmax_off"] == planted) > 0.5


def test_c3_shape_order1_width12(ctx):
    """10k x 12 species x 1 kb, order 1, W = 12 (north_star's target): 4 alignments of the exact length.  (The oracle needs
    ~25 s per 1 kb alignment and thread, which is why the variants below use more, shorter alignments.)"""
    _check(ctx, "c3", 4, 1000, 0.0)


def test_c3_shape_with_gaps(ctx):
    """5 % unobserved leaves: every window has gaps, all of them on the wavefront kernel."""
    _check(ctx, "c3", 8, 300, 0.05)


def test_c3_shape_sparse_gaps_mixed_path(ctx):
    """C3 with a 0.05 % gap rate: most windows on the node-per-thread kernel, the rest through its list."""
    _check(ctx, "c3", 8, 300, 0.0005)


def test_c3_shape_wavefront_kernel(ctx):
    _check(ctx, "c3", 8, 300, 0.0, order1_kernel=1)


@pytest.mark.parametrize("gap_rate", [0.0, 0.05])
def test_c3_data_order0_width12(ctx, gap_rate):
    """C3's data with an order-0 motif (`c3o0`): 12 species, the direct kernel with its state in global memory."""
    _check(ctx, "c3o0", 4, 1000, gap_rate)


def test_c5_shape_30_species_width12(ctx):
    """Genome-scan configuration: 30-species tree, W = 12, order 0, 2 kb alignments (two of them)."""
    _check(ctx, "c5", 2, 2000, 0.0)


def test_c5_shape_with_gaps(ctx):
    _check(ctx, "c5", 8, 300, 0.05)


def _reverse_complement(smp):
    """Every alignment read from its other end with complemented bases (gaps stay gaps)."""
    from phyfoo_b200.data import PhyloSample
    parts = []
    for i in range(smp.n_alignments):
        a = smp.getElementAt(i)[::-1]
        parts.append(np.where(a < 4, 3 - a, a).astype(np.uint8))
    return PhyloSample(np.concatenate(parts), smp.col_offsets)


@pytest.mark.parametrize("name,n_aln", [("c3", None), ("c3o0", None), ("c5", 1500)])
def test_full_size_properties_of_the_benched_configurations(ctx, name, n_aln):
    """The configurations bench.py times, at their full size (C5: 1500 of its 10^6 alignments, 3 * 10^6 windows), through
    properties that need no oracle:
      * ZOOPS weight 1: the posterior over the offsets of an alignment sums to 1, w = (number of alignments, 0);
      * the reverse strand of the reverse-complemented data is the forward strand of the data, window for window and bit for bit
        (the kernels read the same columns with the same symbols), and the sweep counts agree;
      * an E-step over two halves of the sample adds up to the E-step over the whole (ll to rounding, gamma bit for bit);
      * the same call twice returns the same bits;
      * the arg-max calls land on the planted sites."""
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    cfg, newick, pwm, smp, planted = synth.make_config(name, n_aln=n_aln)
    W = cfg["W"]
    fg = PhyloBayesModel(W, cfg["order"], newick, ctx); fg.setCondProbs(pwm)
    bg = PhyloBackground(0, newick, ctx)
    shm = SingleHiddenMotifMixture(fg, bg, zoops=1.0)
    res = shm.getNewWeights(smp)
    off = smp.device(ctx).window_offsets(W)
    per_aln = np.add.reduceat(res["gamma"], off[:-1])
    assert np.isfinite(res["ll"]) and np.max(np.abs(per_aln - 1.0)) < 1e-9
    assert abs(res["w"][0] - smp.n_alignments) < 1e-6 * smp.n_alignments and abs(res["w"][1]) < 1e-9
    assert (res["argmax_off"] == planted).mean() > (0.6 if cfg["order"] == 1 else 0.9)
    res2 = shm.getNewWeights(smp)
    assert res2["ll"] == res["ll"] and np.array_equal(res2["gamma"], res["gamma"]) and np.array_equal(res2["argmax_off"], res["argmax_off"])
    # halves
    h = smp.n_alignments // 2
    a, b = shm.getNewWeights(smp.slice(0, h)), shm.getNewWeights(smp.slice(h, smp.n_alignments))
    assert abs((a["ll"] + b["ll"]) - res["ll"]) <= 1e-12 * abs(res["ll"])
    assert np.array_equal(np.concatenate([a["gamma"], b["gamma"]]), res["gamma"])
    assert np.array_equal(np.concatenate([a["argmax_off"], b["argmax_off"]]), res["argmax_off"])
    # strands
    fwd, sw_f = fg.getLogProbFor(smp, strand=0, want_sweeps=True)
    rc = _reverse_complement(smp)
    rev, sw_r = fg.getLogProbFor(rc, strand=1, want_sweeps=True)
    for i in (0, 1, smp.n_alignments // 3, smp.n_alignments - 1):
        assert np.array_equal(rev[off[i]:off[i + 1]][::-1], fwd[off[i]:off[i + 1]])
    mirrored = np.concatenate([rev[off[i]:off[i + 1]][::-1] for i in range(smp.n_alignments)])
    assert np.array_equal(mirrored, fwd)
    assert int(sw_f.sum()) == int(sw_r.sum())
    del rc
