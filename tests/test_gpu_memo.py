"""The pattern-table path (csrc/memo.cuh) against the generic direct kernel (order0_kernel = 0; the F81-structured and the
run-time compiled kernels re-associate the sums and agree with it to rounding, tests/test_gpu_jit.py): bit-identical scores, sweep counts, gamma and
arg-max calls (same per-tree arithmetic, same summation order), and both against the oracle."""
import numpy as np
import pytest

from util import orc, oracle_bg, oracle_fg, random_codes, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ctx():
    from phyfoo_b200 import core
    c = core.default_context()
    c.set_option("order0_kernel", 0)      # "direct" here = the generic kernel, whose per-tree arithmetic the pattern table shares
    yield c
    c.set_option("pattern_memo", 1)
    c.set_option("order0_kernel", -1)


def _both(ctx, fn):
    ctx.set_option("pattern_memo", 0)
    direct = fn()
    assert ctx.last_score_path() == 0
    ctx.set_option("pattern_memo", 2)
    memo = fn()
    return direct, memo


@pytest.mark.parametrize("newick_name", ["star5", "cat5", "mixed6", "bal6", "star3"])
@pytest.mark.parametrize("gap_rate", [0.0, 0.2])
def test_memo_equals_direct_bit_for_bit(ctx, newick_name, gap_rate):
    from phyfoo_b200 import synth
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel
    from phyfoo_b200.newick import parse_newick
    newick = {"star5": synth.star_newick(5), "cat5": synth.caterpillar_newick(5), "star3": synth.star_newick(3, 0.35),
              "mixed6": "((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.4):0.25):0.1,F:0.3)",
              # three internal nodes at depth 1: three node slots (16 lanes per tree) in the wavefront trajectory kernel
              "bal6": "((A:0.1,B:0.2):0.1,(C:0.3,D:0.1):0.2,(E:0.2,F:0.25):0.15)"}[newick_name]
    S = parse_newick(newick).n_species
    rng = np.random.default_rng(77)
    W = 7
    lens = [W, 60, 5, 131, 33]
    smp = PhyloSample(random_codes(rng, sum(lens), S, gap_rate), np.concatenate([[0], np.cumsum(lens)]))
    fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(synth.random_pwm(W, 0, 5))
    bg = PhyloBackground(0, newick, ctx); bg.setCondProb(np.array([[0.3, 0.2, 0.1, 0.4]]), 0)
    for strand in (0, 1):
        (d, dk), (m, mk) = _both(ctx, lambda: fg.getLogProbFor(smp, strand=strand, want_sweeps=True))
        assert ctx.last_score_path() == 1
        assert np.array_equal(d, m) and np.array_equal(dk, mk)
    (d, dk), (m, mk) = _both(ctx, lambda: bg.getColumnLogProbs(smp, want_sweeps=True))
    assert np.array_equal(d, m) and np.array_equal(dk, mk)
    PhyloBayesModel.CALC_LOGLIKELIHOOD = True
    try:
        d, m = _both(ctx, lambda: fg.getLogProbFor(smp))
        assert np.array_equal(d, m)
    finally:
        PhyloBayesModel.CALC_LOGLIKELIHOOD = False
    # and against the oracle
    om = oracle_fg(newick, fg.getCondProbs() and [p.reshape(1, 4) for p in fg.getCondProbs()])
    ref = np.concatenate([om.score_windows(smp.getElementAt(i))[0] for i in range(smp.n_alignments)])
    ctx.set_option("pattern_memo", 2)
    assert rel_err(fg.getLogProbFor(smp), ref) < 1e-9


def test_memo_estep_and_parameter_updates(ctx):
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture, StrandModel
    newick = synth.caterpillar_newick(5)
    W = 8
    pwm = synth.random_pwm(W, 0, 9)
    smp, planted = synth.make_sample(newick, 30, 90, 4, pwm=pwm, gap_rate=0.02, length_jitter=9)
    fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(pwm)
    bg = PhyloBackground(0, newick, ctx)
    for motif in (fg, StrandModel(fg, 0.4)):
        shm = SingleHiddenMotifMixture(motif, bg, zoops=0.9)
        d, m = _both(ctx, lambda: shm.getNewWeights(smp))
        assert d["ll"] == m["ll"] and np.array_equal(d["gamma"], m["gamma"]) and np.array_equal(d["w"], m["w"])
        assert np.array_equal(d["argmax_off"], m["argmax_off"]) and np.array_equal(d["argmax_strand"], m["argmax_strand"])
        assert d["stats"]["sweeps_fg"] == m["stats"]["sweeps_fg"] and d["stats"]["sweeps_bg"] == m["stats"]["sweeps_bg"]
    # new parameters are picked up by the tables (trajectories are rebuilt per call)
    fg.setCondProb(np.array([[0.7, 0.1, 0.1, 0.1]]), 3)
    shm = SingleHiddenMotifMixture(fg, bg, zoops=0.9)
    d, m = _both(ctx, lambda: shm.getNewWeights(smp))
    assert d["ll"] == m["ll"] and np.array_equal(d["gamma"], m["gamma"])
    ref = orc.estep(oracle_fg(newick, [p.reshape(1, 4) for p in fg.getCondProbs()]), oracle_bg(newick, 0), smp.col_offsets,
                    smp.codes, np.log([0.9, 0.1]))
    assert abs(m["ll"] - ref["ll"]) <= 1e-9 * abs(ref["ll"]) and np.array_equal(m["argmax_off"], ref["argmax_off"])


def test_memo_sweep_cap_fallback_is_bit_identical(ctx):
    """Windows whose stop rule needs more sweeps than the stored trajectories go to the direct kernel: same bits."""
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture, StrandModel
    newick = synth.caterpillar_newick(5)
    W = 9
    pwm = synth.random_pwm(W, 0, 19)
    smp, _ = synth.make_sample(newick, 25, 120, 8, pwm=pwm, gap_rate=0.03)
    fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(pwm)
    bg = PhyloBackground(0, newick, ctx)
    ctx.set_option("pattern_memo", 0)
    d, dk = fg.getLogProbFor(smp, want_sweeps=True)
    shm = SingleHiddenMotifMixture(StrandModel(fg, 0.5), bg, zoops=0.9)
    de = shm.getNewWeights(smp)
    ctx.set_option("pattern_memo", 2)
    try:
        for cap in (1, 3, 6, 100):
            ctx.set_option("pattern_memo_sweep_cap", cap)
            m, mk = fg.getLogProbFor(smp, want_sweeps=True)
            assert ctx.last_score_path() == 1
            assert np.array_equal(d, m) and np.array_equal(dk, mk), cap
            assert fg.last_stats["sweeps_fg"] == int(dk.sum())
            names = [n for n, _ in ctx.last_kernel_times()]
            assert "memo_traj_kernel" in names and "memo_window_kernel" in names
            assert ("score_o0_kernel(fallback list)" in names) == (cap < 100)
            me = shm.getNewWeights(smp)
            assert me["ll"] == de["ll"] and np.array_equal(me["gamma"], de["gamma"])
            assert me["stats"]["sweeps_fg"] == de["stats"]["sweeps_fg"] and me["stats"]["sweeps_bg"] == de["stats"]["sweeps_bg"]
        assert (dk > 6).any() and (dk <= 6).any()      # the cap of 6 really splits the windows between the two kernels
    finally:
        ctx.set_option("pattern_memo_sweep_cap", 64)


def test_memo_two_phase_trajectories_are_bit_identical(ctx):
    """"pattern_memo_split": the second trajectory phase runs on its own stream under the first gather pass; windows the first
    pass leaves are gathered again (and, past the sweep cap, go to the direct kernel).  Same bits for every setting."""
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture, StrandModel
    newick = synth.caterpillar_newick(5)
    W = 10
    pwm = synth.random_pwm(W, 0, 23)
    smp, _ = synth.make_sample(newick, 40, 150, 9, pwm=pwm, gap_rate=0.02, length_jitter=20)
    fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(pwm)
    bg = PhyloBackground(0, newick, ctx)
    shm = SingleHiddenMotifMixture(StrandModel(fg, 0.4), bg, zoops=0.8)
    ctx.set_option("pattern_memo", 0)
    d, dk = fg.getLogProbFor(smp, want_sweeps=True)
    de = shm.getNewWeights(smp)
    ctx.set_option("pattern_memo", 2)
    try:
        for cap, split, wave in ((64, 0, 1), (64, 4, 1), (64, 8, 0), (64, 16, 1), (23, 16, 1), (23, 20, 0), (11, 8, 1), (100, 32, 1),
                                 (64, 0, 0)):
            ctx.set_option("pattern_memo_sweep_cap", cap)
            ctx.set_option("pattern_memo_split", split)
            ctx.set_option("pattern_memo_wave", wave)
            m, mk = fg.getLogProbFor(smp, want_sweeps=True)
            assert ctx.last_score_path() == 1
            assert np.array_equal(d, m) and np.array_equal(dk, mk), (cap, split, wave)
            assert fg.last_stats["sweeps_fg"] == int(dk.sum())
            names = [n for n, _ in ctx.last_kernel_times()]
            two = split > 0 and split + 4 <= cap + 1
            assert ("memo_window_kernel(second pass)" in names) == two, (cap, split, names)
            assert ("memo_traj_kernel(second phase, concurrent)" in names) == two
            me = shm.getNewWeights(smp)
            assert me["ll"] == de["ll"] and np.array_equal(me["gamma"], de["gamma"]), (cap, split)
            assert np.array_equal(me["argmax_off"], de["argmax_off"]) and np.array_equal(me["argmax_strand"], de["argmax_strand"])
            assert me["stats"]["sweeps_fg"] == de["stats"]["sweeps_fg"] and me["stats"]["sweeps_bg"] == de["stats"]["sweeps_bg"]
        assert (dk > 8).any() and (dk < 8).any()
    finally:
        ctx.set_option("pattern_memo_sweep_cap", 64)
        ctx.set_option("pattern_memo_split", 0)
        ctx.set_option("pattern_memo_wave", 1)
    with pytest.raises(Exception):
        ctx.set_option("pattern_memo_split", 6)


@pytest.mark.parametrize("newick_name", ["cat5", "bal6"])
@pytest.mark.parametrize("gap_rate", [0.0, 0.15])
def test_memo_trajectories_past_their_fixed_point(ctx, newick_name, gap_rate):
    """The wavefront trajectory kernel stops a tree once its q repeats bit for bit (fixed point or two-cycle) and fills the
    rest of the row.  With a stop tolerance of 1e-300 a window only stops when its summed free energy repeats exactly, so
    the windows read the rows far past that point -- up to sweep 99 and 100 (both parities of the two-cycle fill)."""
    from phyfoo_b200 import synth
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBayesModel
    from phyfoo_b200.newick import parse_newick
    newick = {"cat5": synth.caterpillar_newick(5),
              "bal6": "((A:0.1,B:0.2):0.1,(C:0.3,D:0.1):0.2,(E:0.2,F:0.25):0.15)"}[newick_name]
    S = parse_newick(newick).n_species
    rng = np.random.default_rng(5)
    W = 6
    lens = [80, 47, 122]
    smp = PhyloSample(random_codes(rng, sum(lens), S, gap_rate), np.concatenate([[0], np.cumsum(lens)]))

    class Tight(PhyloBayesModel):
        TOL = 1e-300

    class Tight99(Tight):
        MAX_SWEEPS = 99

    try:
        seen = []
        for cls in (Tight, Tight99):
            fg = cls(W, 0, newick, ctx); fg.setCondProbs(synth.random_pwm(W, 0, 12))
            ctx.set_option("pattern_memo_sweep_cap", 100)
            ctx.set_option("pattern_memo_split", 0)
            ctx.set_option("pattern_memo", 0)
            d, dk = fg.getLogProbFor(smp, want_sweeps=True)
            ctx.set_option("pattern_memo", 2)
            for wave in (1, 0):
                ctx.set_option("pattern_memo_wave", wave)
                m, mk = fg.getLogProbFor(smp, want_sweeps=True)
                assert ctx.last_score_path() == 1
                assert np.array_equal(dk, mk) and np.array_equal(d, m), (cls.__name__, wave)
            seen.append(dk)
        assert seen[0].max() > 30                                        # rows are read far past the default stop rule
        if newick_name == "cat5" and gap_rate == 0.0:
            assert (seen[0] == 100).any() and (seen[1] == 99).any()      # windows in a two-cycle run to the limit
            assert (seen[0] < 60).any()                                  # and others stop at an exact fixed point
    finally:
        ctx.set_option("pattern_memo_sweep_cap", 64)
        ctx.set_option("pattern_memo_split", 0)
        ctx.set_option("pattern_memo_wave", 1)


def test_memo_does_not_apply_to_many_species_or_order1(ctx):
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBayesModel
    ctx.set_option("pattern_memo", 2)
    newick = synth.random_binary_newick(12, 3)
    smp, _ = synth.make_sample(newick, 2, 30, 1)
    fg = PhyloBayesModel(5, 0, newick, ctx)
    fg.getLogProbFor(smp)
    assert ctx.last_score_path() == 0
    n5 = synth.star_newick(5)
    smp5, _ = synth.make_sample(n5, 2, 30, 1)
    fg1 = PhyloBayesModel(5, 1, n5, ctx)
    fg1.getLogProbFor(smp5)
    assert ctx.last_score_path() == 2
    from phyfoo_b200.lib import IllegalArgumentException
    with pytest.raises(IllegalArgumentException):
        ctx.set_option("pattern_memo", 7)
    with pytest.raises(IllegalArgumentException):
        ctx.set_option("no_such_option", 1)
