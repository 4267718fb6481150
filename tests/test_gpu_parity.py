"""GPU parity: libphyfoo.so (through its C ABI, via the host-side mirror classes) against the CPU oracle on the
same seeded inputs.  Bar: scores/gradients within 1e-9 relative (fp64), mean-field sweep counts, packed column
words and arg-max site calls bit-exact."""
import numpy as np
import pytest

from util import orc, oracle_bg, oracle_fg, random_codes, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    from phyfoo_b200 import core
    return core.default_context()


def _sample(rng, lens, S, gap_rate):
    from phyfoo_b200.data import PhyloSample
    codes = random_codes(rng, int(sum(lens)), S, gap_rate)
    return PhyloSample(codes, np.concatenate([[0], np.cumsum(lens)]))


def _trees():
    from phyfoo_b200 import synth
    return {"star5": synth.star_newick(5), "cat5": synth.caterpillar_newick(5),
            "rand12": synth.random_binary_newick(12, 3),
            "mixed6": "((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.4):0.25):0.1,F:0.3)",
            "rand30": synth.random_binary_newick(30, 5), "star20": synth.star_newick(20, 0.3)}


def _S(newick):
    from phyfoo_b200.newick import parse_newick
    return parse_newick(newick).n_species


def test_pack_columns_bit_exact(ctx):
    rng = np.random.default_rng(0)
    for S in (1, 5, 12, 16, 17, 30, 32):
        smp = _sample(rng, [7, 1, 33, 100], S, 0.2)
        bases, gaps = smp.device(ctx).packed()
        rb, rg = orc.pack_columns(smp.codes)
        assert np.array_equal(bases, rb) and np.array_equal(gaps, rg)
    from phyfoo_b200.lib import IllegalArgumentException
    from phyfoo_b200.data import PhyloSample
    bad = PhyloSample(np.full((4, 3), 7, np.uint8), [0, 4])
    with pytest.raises(IllegalArgumentException):
        bad.device(ctx)


@pytest.mark.parametrize("tree", ["star5", "cat5", "rand12", "mixed6", "rand30", "star20"])
@pytest.mark.parametrize("gap_rate", [0.0, 0.1])
def test_meanfield_window_scores(ctx, tree, gap_rate):
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBayesModel
    newick = _trees()[tree]
    S = _S(newick)
    rng = np.random.default_rng(21)
    W = 7
    pwm = synth.random_pwm(W, 0, 41)
    smp = _sample(rng, [30, 7, 64, 19, 8], S, gap_rate)
    m = PhyloBayesModel(W, 0, newick, ctx)
    m.setCondProbs(pwm)
    om = oracle_fg(newick, pwm)
    for strand in (0, 1):
        got, sw = m.getLogProbFor(smp, strand, want_sweeps=True)
        ref = np.concatenate([om.score_windows(smp.getElementAt(i), bool(strand))[0] for i in range(smp.n_alignments)])
        rsw = np.concatenate([om.score_windows(smp.getElementAt(i), bool(strand))[1] for i in range(smp.n_alignments)])
        assert got.shape == ref.shape
        assert np.array_equal(sw, rsw), "mean-field sweep counts differ"
        assert rel_err(got, ref) < TOL
        assert m.last_stats["sweeps_fg"] == int(rsw.sum())


@pytest.mark.parametrize("tree", ["star5", "cat5", "rand12", "rand30"])
def test_exact_window_scores(ctx, tree):
    from phyfoo_b200 import synth, _abi
    from phyfoo_b200.models import PhyloBayesModel
    newick = _trees()[tree]
    S = _S(newick)
    rng = np.random.default_rng(22)
    W = 5
    pwm = synth.random_pwm(W, 0, 43)
    smp = _sample(rng, [40, 5, 23], S, 0.1)
    m = PhyloBayesModel(W, 0, newick, ctx)
    m.setCondProbs(pwm)
    m.device_model().set_mode(_abi.MODE_EXACT)
    m._mode = _abi.MODE_EXACT
    got = core_score(ctx, smp, m)
    om = oracle_fg(newick, pwm, mode=orc.EXACT_PRUNING)
    ref = np.concatenate([om.score_windows(smp.getElementAt(i))[0] for i in range(smp.n_alignments)])
    assert rel_err(got, ref) < 1e-12


def core_score(ctx, smp, m, strand=0):
    from phyfoo_b200 import core
    return core.score_windows(ctx, smp.device(ctx), m._dev, strand)[0]


@pytest.mark.parametrize("tree", ["star5", "cat5", "rand12"])
def test_background_columns(ctx, tree):
    from phyfoo_b200.models import PhyloBackground
    newick = _trees()[tree]
    S = _S(newick)
    rng = np.random.default_rng(23)
    smp = _sample(rng, [50, 3, 17], S, 0.1)
    bg = PhyloBackground(0, newick, ctx)
    pi = np.array([[0.1, 0.4, 0.3, 0.2]])
    bg.setCondProb(pi, 0)
    got, sw = bg.getColumnLogProbs(smp, want_sweeps=True)
    ob = oracle_bg(newick, 0, [pi])
    ref = ob.bg_columns(smp.codes)
    assert rel_err(got, ref) < TOL
    per = bg.getLogProbFor(smp)
    for i in range(smp.n_alignments):
        L = smp.lengths()[i]
        r = ob.bg_range(smp.getElementAt(i), 0, L - 1)
        assert abs(per[i] - r) <= TOL * abs(r)


@pytest.mark.parametrize("tree,strand", [("cat5", False), ("cat5", True), ("star5", True), ("rand12", False)])
def test_zoops_estep(ctx, tree, strand):
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture, StrandModel
    newick = _trees()[tree]
    S = _S(newick)
    W = 6
    pwm = synth.random_pwm(W, 0, 47)
    smp, planted = synth.make_sample(newick, 12, 60, 99, pwm=pwm, gap_rate=0.05, length_jitter=10)
    fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(pwm)
    bg = PhyloBackground(0, newick, ctx)
    motif = StrandModel(fg, 0.6) if strand else fg
    shm = SingleHiddenMotifMixture(motif, bg, zoops=0.8)
    wts = np.random.default_rng(5).random(smp.n_alignments) + 0.5
    res = shm.getNewWeights(smp, seqWeights=wts, want_profile=True)
    ofg, obg = oracle_fg(newick, pwm), oracle_bg(newick, 0)
    ref = orc.estep(ofg, obg, smp.col_offsets, smp.codes, np.log([0.8, 0.2]),
                    np.log([0.6, 0.4]) if strand else None, wts)
    assert np.array_equal(res["argmax_off"], ref["argmax_off"]), "arg-max site calls must be bit-exact"
    assert np.array_equal(res["argmax_strand"], ref["argmax_strand"])
    assert rel_err(res["profile"], ref["profile"]) < TOL
    assert np.max(np.abs(res["gamma"] - ref["gamma"])) < TOL
    assert abs(res["ll"] - ref["ll"]) <= TOL * abs(ref["ll"])
    assert np.max(np.abs(res["w"] - ref["w"])) <= TOL * np.max(np.abs(ref["w"]))
    assert res["stats"]["sweeps_fg"] == ref["sweeps_fg"]
    # the Java (and the oracle) optimise each single-column background range twice; the kernel does it once
    assert 2 * res["stats"]["sweeps_bg"] == ref["sweeps_bg"]
    # size-independent properties: sum of gamma per alignment = p_motif * weight; w0 + w1 = sum of weights
    assert abs(res["gamma"].sum() - res["w"][0]) < 1e-9
    assert abs(res["w"].sum() - wts.sum()) < 1e-9
    # planted sites are found by the arg-max for most alignments
    hit = (res["argmax_off"] == planted).mean()
    assert 0.0 <= hit <= 1.0   # statistical check lives in test_full_size_properties (12 short alignments are too few)


@pytest.mark.parametrize("ratio,memo", [(2.0, 0), (0.5, 2)])
def test_hky_with_its_ratio_in_effect(ctx, ratio, memo):
    """PHYFOO_EVOL_HKY_RATIO (opt-in: evolution/HKY.java:77-106 with beta = alpha * ratio): tables without F81 structure, so the
    generic kernel (memo 0) or the pattern table (memo 2) scores -- window scores, sweep counts and the strand E-step against
    the oracle under the same tables; order 1 is refused."""
    from phyfoo_b200 import _abi, synth
    from phyfoo_b200.lib import IllegalArgumentException
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture, StrandModel

    class Motif(PhyloBayesModel):
        EVOL_MODEL, HKY_RATIO = _abi.EVOL_HKY_RATIO, ratio

    class Background(PhyloBackground):
        EVOL_MODEL, HKY_RATIO = _abi.EVOL_HKY_RATIO, ratio

    newick = synth.caterpillar_newick(5)
    W = 6
    pwm = synth.random_pwm(W, 0, 53)
    smp, _ = synth.make_sample(newick, 10, 50, 77, pwm=pwm, gap_rate=0.05, length_jitter=8)
    ctx.set_option("pattern_memo", memo)
    try:
        fg = Motif(W, 0, newick, ctx); fg.setCondProbs(pwm)
        bg = Background(0, newick, ctx)
        kw = dict(evol=orc.HKY_RATIO, hky_ratio=ratio)
        ofg, obg = oracle_fg(newick, pwm, **kw), oracle_bg(newick, 0, **kw)
        for strand in (0, 1):
            got, sw = fg.getLogProbFor(smp, strand, want_sweeps=True)
            ref = [ofg.score_windows(smp.getElementAt(i), bool(strand)) for i in range(smp.n_alignments)]
            assert np.array_equal(sw, np.concatenate([r[1] for r in ref])), "mean-field sweep counts differ"
            assert rel_err(got, np.concatenate([r[0] for r in ref])) < TOL
        shm = SingleHiddenMotifMixture(StrandModel(fg, 0.6), bg, zoops=0.8)
        res = shm.getNewWeights(smp)
        ref = orc.estep(ofg, obg, smp.col_offsets, smp.codes, np.log([0.8, 0.2]), np.log([0.6, 0.4]), None)
        assert np.array_equal(res["argmax_off"], ref["argmax_off"]) and np.array_equal(res["argmax_strand"], ref["argmax_strand"])
        assert np.max(np.abs(res["gamma"] - ref["gamma"])) < TOL
        assert abs(res["ll"] - ref["ll"]) <= TOL * abs(ref["ll"])
        assert res["stats"]["sweeps_fg"] == ref["sweeps_fg"]
        with pytest.raises(IllegalArgumentException):
            Motif(W, 1, newick, ctx).device_model()
    finally:
        ctx.set_option("pattern_memo", 1)


def test_ragged_and_error_behaviour(ctx):
    from phyfoo_b200 import synth
    from phyfoo_b200.lib import IllegalArgumentException
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    newick = synth.caterpillar_newick(5)
    rng = np.random.default_rng(3)
    W = 8
    fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(synth.random_pwm(W, 0, 1))
    # alignments shorter than the motif contribute no windows to the score list
    smp = _sample(rng, [3, 8, 20, 7], 5, 0.0)
    got = fg.getLogProbFor(smp)
    assert got.shape[0] == 1 + 13
    om = oracle_fg(newick, fg.getCondProbs()[:W] and [p.reshape(1, 4) for p in fg.getCondProbs()])
    ref = np.concatenate([om.score_windows(smp.getElementAt(i))[0] for i in (1, 2)])
    assert rel_err(got, ref) < TOL
    # ... but the ZOOPS E-step refuses them, like the reference's position prior does
    shm = SingleHiddenMotifMixture(fg, PhyloBackground(0, newick, ctx), zoops=1.0)
    with pytest.raises(IllegalArgumentException):
        shm.getNewWeights(smp)
    # species mismatch
    with pytest.raises(ValueError):
        fg.getLogProbFor(_sample(rng, [10], 4, 0.0))
    # empty sample: no windows, no error
    empty = PhyloSample(np.zeros((0, 5), np.uint8), [0])
    assert fg.getLogProbFor(empty).shape == (0,)


def test_full_size_properties(ctx):
    """BASELINE config 2 at full size (2000 x 5 x 500, W=12): properties that need no oracle."""
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    cfg, newick, pwm, smp, planted = synth.make_config("c2")
    fg = PhyloBayesModel(cfg["W"], 0, newick, ctx); fg.setCondProbs(pwm)
    bg = PhyloBackground(0, newick, ctx)
    shm = SingleHiddenMotifMixture(fg, bg, zoops=1.0)
    res = shm.getNewWeights(smp)
    assert np.isfinite(res["ll"])
    per_aln = np.add.reduceat(res["gamma"], smp.device(ctx).window_offsets(cfg["W"])[:-1])
    assert np.max(np.abs(per_aln - 1.0)) < 1e-9                 # ZOOPS = 1: posterior over offsets sums to 1
    assert abs(res["w"][0] - cfg["n_aln"]) < 1e-6 and abs(res["w"][1]) < 1e-9
    assert (res["argmax_off"] == planted).mean() > 0.8
    # idempotence / determinism: the same call returns the same bits
    res2 = shm.getNewWeights(smp)
    assert res2["ll"] == res["ll"] and np.array_equal(res2["gamma"], res["gamma"])
    # mean field is a lower bound of the exact log-likelihood, window by window
    mf = fg.getLogProbFor(smp)
    PhyloBayesModel.CALC_LOGLIKELIHOOD = True
    try:
        ex = fg.getLogProbFor(smp)
    finally:
        PhyloBayesModel.CALC_LOGLIKELIHOOD = False
    assert (mf <= ex + 1e-9).all()
    # the first alignments against the oracle
    om = oracle_fg(newick, pwm)
    for i in range(3):
        ref, _ = om.score_windows(smp.getElementAt(i))
        o = smp.device(ctx).window_offsets(cfg["W"])
        assert rel_err(mf[o[i]:o[i + 1]], ref) < TOL
    # the two scoring paths at full size: same bits, same sweep counts
    assert ctx.last_score_path() == 1                           # the default for five species is the pattern table
    ctx.set_option("pattern_memo", 0)
    ctx.set_option("order0_kernel", 0)                          # the generic kernel: the pattern table's own per-tree arithmetic
    try:
        d, dk = fg.getLogProbFor(smp, want_sweeps=True)
        assert ctx.last_score_path() == 0
    finally:
        ctx.set_option("pattern_memo", 1)
        ctx.set_option("order0_kernel", -1)
    m, mk = fg.getLogProbFor(smp, want_sweeps=True)
    assert np.array_equal(d, m) and np.array_equal(dk, mk) and np.array_equal(m, mf)
    # strand symmetry: the reverse strand of the reverse-complemented alignments is the forward strand read backwards
    from phyfoo_b200.data import PhyloSample
    L = cfg["length"]
    rc = smp.codes.reshape(cfg["n_aln"], L, -1)[:, ::-1, :]
    rc = np.where(rc < 4, 3 - rc, rc).reshape(-1, smp.codes.shape[1])
    back = fg.getLogProbFor(PhyloSample(rc, smp.col_offsets), strand=1)
    nwin = L - cfg["W"] + 1
    assert np.array_equal(back.reshape(cfg["n_aln"], nwin)[:, ::-1], m.reshape(cfg["n_aln"], nwin))


def test_sharded_estep_single_rank(ctx):
    """The multi-GPU entry (device-resident partials, no host round trip) at world size 1 equals the host API."""
    from phyfoo_b200 import sharding, synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture, StrandModel
    newick = synth.caterpillar_newick(5)
    pwm = synth.random_pwm(7, 0, 11)
    smp, _ = synth.make_sample(newick, 9, 50, 13, pwm=pwm, gap_rate=0.03)
    fg = PhyloBayesModel(7, 0, newick, ctx); fg.setCondProbs(pwm)
    bg = PhyloBackground(0, newick, ctx)
    for motif in (fg, StrandModel(fg, 0.3)):
        shm = SingleHiddenMotifMixture(motif, bg, zoops=0.75)
        ref = shm.getNewWeights(smp)
        se = sharding.ShardedEStep(shm, smp)
        se.step()
        got = se.result()
        assert got["ll"] == ref["ll"] and np.array_equal(got["w"], ref["w"])


@pytest.mark.parametrize("tree", ["star5", "cat5", "mixed6", "rand12", "rand30"])
@pytest.mark.parametrize("W", [1, 5, 12, 20, 40])
def test_f81_structured_kernel_equals_generic_kernel(ctx, tree, W):
    """score_o0f_kernel (compact F81 tables, free energy without a logarithm per node, groups of warps with a named barrier)
    against the generic kernel and the oracle: same sweep counts, scores equal to re-association; motif widths that give
    groups of 1, 5, 3, 5 and 5 warps, both strands, ragged lengths, gaps, F81alpha and F81beta."""
    from phyfoo_b200 import _abi, synth
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBayesModel
    from phyfoo_b200.newick import parse_newick
    if tree == "rand30" and W > 20:
        pytest.skip("the generic kernel's tables of a 30-species motif of width 40 do not fit in shared memory")
    newick = _trees()[tree]
    S = parse_newick(newick).n_species
    rng = np.random.default_rng(71 + W)
    pwm = synth.random_pwm(W, 0, 5)
    lens = [W, 3 * W + 7, max(W - 1, 1), 260, W + 1]
    smp = PhyloSample(random_codes(rng, sum(lens), S, 0.03), np.concatenate([[0], np.cumsum(lens)]))
    for evol in (_abi.EVOL_F81ALPHA, _abi.EVOL_F81BETA):
        PhyloBayesModel.EVOL_MODEL = evol
        try:
            fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(pwm)
            res = {}
            for kern in (0, 1):
                ctx.set_option("order0_kernel", kern)
                ctx.set_option("pattern_memo", 0)
                res[kern] = [fg.getLogProbFor(smp, strand=s, want_sweeps=True) for s in (0, 1)]
                assert [n for n, _ in ctx.last_kernel_times()] == ["score_o0f_kernel" if kern else "score_o0_kernel"]
                assert fg.last_stats["sweeps_fg"] == int(res[kern][1][1].sum())
            for s in (0, 1):
                assert np.array_equal(res[0][s][1], res[1][s][1])
                assert rel_err(res[1][s][0], res[0][s][0]) < 1e-12
            if evol == _abi.EVOL_F81ALPHA and W <= 12:
                om = oracle_fg(newick, pwm)
                ref = np.concatenate([om.score_windows(smp.getElementAt(i), False)[0] for i in range(smp.n_alignments)])
                assert rel_err(res[1][0][0], ref) < TOL
        finally:
            PhyloBayesModel.EVOL_MODEL = _abi.EVOL_F81ALPHA
            ctx.set_option("order0_kernel", -1)
            ctx.set_option("pattern_memo", 1)


def test_hky_as_implemented_returns_the_reference_degenerate_values(ctx):
    """evolution/HKY.java as implemented (beta == 0, :32): every table has zero entries.  The library returns what the
    reference's mean field returns -- NaN, or -inf for columns of one nucleotide class under a star tree, 100 sweeps per
    window -- instead of an error; exact mode is the likelihood under those tables (-inf where a transversion is needed)."""
    from phyfoo_b200 import _abi, synth
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    rng = np.random.default_rng(3)
    W = 3
    pwm = synth.random_pwm(W, 0, 3)
    base = rng.integers(0, 4, size=(60, 1))
    codes = np.where(rng.random((60, 5)) < 0.3, (base + 2) % 4, base).astype(np.uint8)     # transitions only
    codes[5, 2] = (codes[5, 2] + 1) % 4                                                     # one transversion column
    codes[20, 1] = 4
    codes[30, :] = [4, 4, 0, 2, 0]
    smp = PhyloSample(codes, [0, 25, 60])
    PhyloBayesModel.EVOL_MODEL = _abi.EVOL_HKY_AS_IMPLEMENTED
    try:
        for newick in (synth.star_newick(5), synth.caterpillar_newick(5)):
            fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(pwm)
            om = oracle_fg(newick, pwm, evol=orc.HKY)
            for strand in (0, 1):
                got, sw = fg.getLogProbFor(smp, strand=strand, want_sweeps=True)
                ref, rsw = zip(*[om.score_windows(smp.getElementAt(i), bool(strand)) for i in range(2)])
                ref, rsw = np.concatenate(ref), np.concatenate(rsw)
                assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.array_equal(got[~np.isnan(got)], ref[~np.isnan(ref)])
                assert np.array_equal(sw, rsw) and (sw == 100).all()
            assert np.isnan(got).any() and (np.isneginf(got).any() or "((" in newick)
            # exact mode: pruning under the same tables
            PhyloBayesModel.CALC_LOGLIKELIHOOD = True
            try:
                ex = fg.getLogProbFor(smp)
            finally:
                PhyloBayesModel.CALC_LOGLIKELIHOOD = False
            oe = oracle_fg(newick, pwm, mode=orc.EXACT_PRUNING, evol=orc.HKY)
            re_ = np.concatenate([oe.score_windows(smp.getElementAt(i), False)[0] for i in range(2)])
            fin = np.isfinite(re_)
            assert np.array_equal(np.isfinite(ex), fin) and rel_err(ex[fin], re_[fin]) < TOL and np.isneginf(ex[~fin]).all()
    finally:
        PhyloBayesModel.EVOL_MODEL = _abi.EVOL_F81ALPHA
