"""Order-0 scoring kernel compiled at run time for a model's tree (csrc/jit_o0.hpp, option order0_kernel = 2): the same pass
as the precompiled F81-structured kernel (order0_kernel = 1) written out as straight-line code for the tree.  Scores and sweep
counts must equal that kernel's bit for bit (same arithmetic, same helpers), and the oracle's within 1e-9 / exactly."""
import os

import numpy as np
import pytest

from util import orc, oracle_fg, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from phyfoo_b200 import core
    return core.default_context()


def _trees():
    from phyfoo_b200 import synth
    return {"star5": synth.star_newick(5), "cat5": synth.caterpillar_newick(5), "rand12": synth.random_binary_newick(12, 0x5EED0003),
            "rand30": synth.random_binary_newick(30, 0x5EED0005), "multif6": "((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.4):0.25):0.1,F:0.3)",
            "rand8": synth.random_binary_newick(8, 7)}


@pytest.mark.parametrize("tree", ["star5", "cat5", "rand8", "multif6", "rand12", "rand30"])
@pytest.mark.parametrize("gap_rate", [0.0, 0.06])
@pytest.mark.parametrize("W", [12, 5])
def test_compiled_kernel_equals_precompiled_kernel_and_oracle(ctx, tree, gap_rate, W):
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBayesModel
    newick = _trees()[tree]
    pwm = synth.random_pwm(W, 0, 5)
    smp, _ = synth.make_sample(newick, 6, 70, 9, pwm=pwm, gap_rate=gap_rate, length_jitter=20)
    fg = PhyloBayesModel(W, 0, newick, ctx)
    fg.setCondProbs(pwm)
    out = {}
    ctx.set_option("pattern_memo", 0)                # small trees would take the pattern-table path
    try:
        for kernel in (1, 2):
            ctx.set_option("order0_kernel", kernel)
            for strand in (0, 1):
                out[kernel, strand] = fg.getLogProbFor(smp, strand=strand, want_sweeps=True)
                names = [n for n, _ in ctx.last_kernel_times()]
                assert names == ["score_o0j_kernel" if kernel == 2 else "score_o0f_kernel"], (names, ctx.jit_note())
    finally:
        ctx.set_option("order0_kernel", -1)
        ctx.set_option("pattern_memo", 1)
    ofg = oracle_fg(newick, pwm)
    for strand in (0, 1):
        f, sf = out[1, strand]
        j, sj = out[2, strand]
        assert np.array_equal(sf, sj)
        assert np.array_equal(f, j), float(np.max(np.abs(f - j)))
        ref = np.concatenate([ofg.score_windows(smp.getElementAt(i), bool(strand))[0] for i in range(smp.n_alignments)])
        assert rel_err(j, ref) < 1e-9


def test_compiled_kernel_is_the_default_for_large_trees_and_is_cached(ctx):
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    newick = _trees()["rand12"]
    pwm = synth.random_pwm(9, 0, 3)
    smp, _ = synth.make_sample(newick, 5, 60, 4, pwm=pwm)
    fg = PhyloBayesModel(9, 0, newick, ctx)
    fg.setCondProbs(pwm)
    bg = PhyloBackground(0, newick, ctx)
    res = SingleHiddenMotifMixture(fg, bg, zoops=0.9).getNewWeights(smp)
    names = [n for n, _ in ctx.last_kernel_times()]
    assert names.count("score_o0j_kernel") == 2, (names, ctx.jit_note())           # background columns and motif windows
    ref = orc.estep(oracle_fg(newick, pwm), orc.Model(orc.BG, 0, 0, newick), smp.col_offsets, smp.codes, np.log([0.9, 0.1]),
                    threads=os.cpu_count() or 1)
    assert abs(res["ll"] - ref["ll"]) <= 1e-9 * abs(ref["ll"]) and np.array_equal(res["argmax_off"], ref["argmax_off"])
    assert res["stats"]["sweeps_fg"] == ref["sweeps_fg"]
    # a second model with the same tree and width reuses the loaded kernel (no second compilation: well under a second)
    import time
    fg2 = PhyloBayesModel(9, 0, newick, ctx)
    fg2.setCondProbs(synth.random_pwm(9, 0, 4))
    t0 = time.perf_counter()
    fg2.getLogProbFor(smp)
    assert time.perf_counter() - t0 < 0.5


@pytest.mark.parametrize("tree", ["cat5", "rand12"])
@pytest.mark.parametrize("W", [1, 20, 40])
def test_compiled_kernel_other_widths_and_f81beta(ctx, tree, W):
    """Group shapes of 1, 5 and 5 warps (W = 1 is the background model's shape), F81beta tables, ragged lengths with alignments
    shorter than the motif."""
    from phyfoo_b200 import _abi, synth
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBayesModel
    from phyfoo_b200.newick import parse_newick
    from util import random_codes
    newick = _trees()[tree]
    S = parse_newick(newick).n_species
    rng = np.random.default_rng(5 + W)
    pwm = synth.random_pwm(W, 0, 6)
    lens = [W, 3 * W + 7, max(W - 1, 1), 150, W + 1]
    smp = PhyloSample(random_codes(rng, sum(lens), S, 0.03), np.concatenate([[0], np.cumsum(lens)]))
    ctx.set_option("pattern_memo", 0)
    try:
        for evol in (_abi.EVOL_F81ALPHA, _abi.EVOL_F81BETA):
            PhyloBayesModel.EVOL_MODEL = evol
            fg = PhyloBayesModel(W, 0, newick, ctx)
            fg.setCondProbs(pwm)
            res = {}
            for kern in (1, 2):
                ctx.set_option("order0_kernel", kern)
                res[kern] = [fg.getLogProbFor(smp, strand=s, want_sweeps=True) for s in (0, 1)]
                assert [n for n, _ in ctx.last_kernel_times()] == ["score_o0j_kernel" if kern == 2 else "score_o0f_kernel"], ctx.jit_note()
            for s in (0, 1):
                assert np.array_equal(res[1][s][1], res[2][s][1]) and np.array_equal(res[1][s][0], res[2][s][0])
    finally:
        PhyloBayesModel.EVOL_MODEL = _abi.EVOL_F81ALPHA
        ctx.set_option("order0_kernel", -1)
        ctx.set_option("pattern_memo", 1)
