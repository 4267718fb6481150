"""GPU parity of the order-1 phylogenetic Bayesian network motif (BASELINE config 3's model; reference:
models/PhyloBayesModel.java:356-370 with model.fg.order = 1) against the CPU oracle: window scores within 1e-9
relative, mean-field sweep counts and arg-max site calls bit-exact, with and without gaps, both strands."""
import numpy as np
import pytest

from util import orc, oracle_bg, oracle_fg, random_codes, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    from phyfoo_b200 import core
    return core.default_context()


def _trees():
    from phyfoo_b200 import synth
    return {"star5": synth.star_newick(5), "cat5": synth.caterpillar_newick(5),
            "rand12": synth.random_binary_newick(12, 3),
            "mixed6": "((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.4):0.25):0.1,F:0.3)"}


@pytest.mark.parametrize("tree", ["star5", "cat5", "mixed6", "rand12"])
@pytest.mark.parametrize("gap_rate", [0.0, 0.15])
def test_order1_window_scores(ctx, tree, gap_rate):
    from phyfoo_b200 import synth
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBayesModel
    from phyfoo_b200.newick import parse_newick
    newick = _trees()[tree]
    S = parse_newick(newick).n_species
    rng = np.random.default_rng(41)
    W = 5 if tree == "rand12" else 6
    pwm = synth.random_pwm(W, 1, 61)
    lens = [W, 23, 3, 40]
    smp = PhyloSample(random_codes(rng, sum(lens), S, gap_rate), np.concatenate([[0], np.cumsum(lens)]))
    fg = PhyloBayesModel(W, 1, newick, ctx)
    fg.setCondProbs(pwm)
    om = oracle_fg(newick, pwm, order=1)
    for strand in (0, 1):
        got, sw = fg.getLogProbFor(smp, strand=strand, want_sweeps=True)
        ref, rsw = [], []
        for i in range(smp.n_alignments):
            r, k = om.score_windows(smp.getElementAt(i), bool(strand))
            ref.append(r); rsw.append(k)
        ref, rsw = np.concatenate(ref), np.concatenate(rsw)
        assert got.shape == ref.shape
        assert np.array_equal(sw, rsw), (sw[:10], rsw[:10])
        assert rel_err(got, ref) < TOL
        assert fg.last_stats["sweeps_fg"] == int(rsw.sum())


def test_order1_zoops_estep(ctx):
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture, StrandModel
    newick = synth.caterpillar_newick(5)
    W = 6
    pwm = synth.random_pwm(W, 1, 67)
    smp, planted = synth.make_sample(newick, 10, 45, 3, pwm=pwm, gap_rate=0.04, length_jitter=6)
    fg = PhyloBayesModel(W, 1, newick, ctx); fg.setCondProbs(pwm)
    bg = PhyloBackground(0, newick, ctx)
    for strand in (False, True):
        shm = SingleHiddenMotifMixture(StrandModel(fg, 0.5) if strand else fg, bg, zoops=0.85)
        res = shm.getNewWeights(smp)
        ref = orc.estep(oracle_fg(newick, pwm, order=1), oracle_bg(newick, 0), smp.col_offsets, smp.codes, np.log([0.85, 0.15]),
                        np.log([0.5, 0.5]) if strand else None)
        assert np.array_equal(res["argmax_off"], ref["argmax_off"]) and np.array_equal(res["argmax_strand"], ref["argmax_strand"])
        assert np.max(np.abs(res["gamma"] - ref["gamma"])) < TOL
        assert abs(res["ll"] - ref["ll"]) <= TOL * abs(ref["ll"])
        assert res["stats"]["sweeps_fg"] == ref["sweeps_fg"]


def test_order1_unsupported_paths_fail_loudly(ctx):
    from phyfoo_b200 import synth
    from phyfoo_b200.lib import PhyfooError
    from phyfoo_b200.models import PhyloBayesModel
    newick = synth.star_newick(5)
    smp, _ = synth.make_sample(newick, 2, 20, 1)
    fg = PhyloBayesModel(4, 1, newick, ctx)
    PhyloBayesModel.CALC_LOGLIKELIHOOD = True
    try:
        with pytest.raises(PhyfooError):
            fg.getLogProbFor(smp)
    finally:
        PhyloBayesModel.CALC_LOGLIKELIHOOD = False


@pytest.mark.parametrize("tree", ["cat5", "mixed6", "rand12"])
@pytest.mark.parametrize("gap_rate", [0.0, 0.12])
def test_order1_fixed_q_objective_and_gradient(ctx, tree, gap_rate):
    """M-step objective of the order-1 motif: f(lambda) and its forward-difference gradient (dim 4 at position 0, 16 = four
    context rows elsewhere) against the oracle's FESet on the same selected windows, incl. reverse-complement windows."""
    from phyfoo_b200 import core, synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    from phyfoo_b200.newick import parse_newick
    EPS = 1e-4
    newick = _trees()[tree]
    S = parse_newick(newick).n_species
    W = 4 if tree == "rand12" else 5
    pwm = synth.random_pwm(W, 1, 71)
    smp, _ = synth.make_sample(newick, 14, 40, 9, pwm=pwm, gap_rate=gap_rate, length_jitter=5)
    fg = PhyloBayesModel(W, 1, newick, ctx); fg.setCondProbs(pwm)
    bg = PhyloBackground(0, newick, ctx)
    res = SingleHiddenMotifMixture(fg, bg, zoops=0.9).getNewWeights(smp)
    ds = smp.device(ctx)
    sa, so, sw = core.select(ctx, ds, W, res["gamma"])
    strand = (np.arange(sa.size) % 4 == 1).astype(np.uint8)
    fe = core.FESet(ctx, ds, fg.device_model(), sa, so, sw, strand)
    om = oracle_fg(newick, pwm, order=1)
    cols = np.stack([smp.getElementAt(a)[o:o + W] for a, o in zip(sa, so)])
    rc = np.where(cols[:, ::-1, :] < 4, 3 - cols[:, ::-1, :], 4).astype(np.uint8)
    cols = np.where(strand[:, None, None] == 1, rc, cols)
    ofe = orc.FESet(om, cols, sw)
    rng = np.random.default_rng(5)

    def tol(g, f0):
        return TOL * np.abs(g) + 64 * np.finfo(float).eps * abs(f0) / EPS

    for pos in (0, W - 1, 1):
        rows = 1 if pos == 0 else 4
        lam = (np.log(rng.dirichlet(np.ones(4), size=rows)) + rng.normal(size=(rows, 1))).ravel()
        f, rf = fe.eval(pos, lam), ofe.eval(pos, lam)
        assert abs(f - rf) <= TOL * abs(rf), (pos, f, rf)
        f0, g = fe.grad(pos, lam, EPS)
        r0, rg = ofe.grad(pos, lam, EPS)
        assert abs(f0 - r0) <= TOL * abs(r0)
        assert np.all(np.abs(g - rg) <= tol(rg, r0)), (pos, g, rg)
        assert np.allclose(fg.device_model().get_pi(pos, 4 * rows), om.get_pi(pos), rtol=0, atol=1e-15)
    lam_all = np.concatenate([np.log(rng.dirichlet(np.ones(4), size=(1 if t == 0 else 4))).ravel() for t in range(W)])
    f0, g = fe.grad(-1, lam_all, EPS)
    r0, rg = ofe.grad(-1, lam_all, EPS)
    assert abs(f0 - r0) <= TOL * abs(r0)
    assert np.all(np.abs(g - rg) <= tol(rg, r0))
    assert abs(fe.eval(-1, lam_all) - r0) <= TOL * abs(r0)
    # back at the parameters the q were optimised for: f = sum_u w_u score_u
    for t in range(W):
        fe.eval(t, np.log(pwm[t].ravel())); ofe.eval(t, np.log(pwm[t].ravel()))
    direct = sum(w * om.score_window(c)[0] for w, c in zip(sw, cols))
    assert abs(fe.eval(0, np.log(pwm[0].ravel())) - direct) <= 1e-9 * abs(direct)
    # the same objective from sufficient statistics (fe1_count_kernel + host dot products inside the library): value and
    # forward-difference gradient equal the per-window evaluation along a walk over the positions, also after branch changes
    ss = core.SSObjective(fe)
    assert abs(ss.value() - direct) <= 1e-10 * abs(direct)
    for pos in (1, 0, W - 1, 1):
        rows = 1 if pos == 0 else 4
        lam = (np.log(rng.dirichlet(np.ones(4), size=rows)) + rng.normal(size=(rows, 1))).ravel()
        f, rf = ss.eval(pos, lam), fe.eval(pos, lam)
        assert abs(f - rf) <= 1e-11 * abs(rf), (pos, f, rf)
        f0, g = ss.grad(pos, lam, EPS)
        r0, rg = fe.grad(pos, lam, EPS)
        assert abs(f0 - r0) <= 1e-11 * abs(r0) and np.all(np.abs(g - rg) <= tol(rg, r0))
        assert np.allclose(ss.get_pi(pos, 4 * rows), fg.device_model().get_pi(pos, 4 * rows), rtol=0, atol=1e-15)
    with pytest.raises(ValueError):
        ss.eval(1, np.zeros(4))                                    # an order-1 position t >= 1 has 16 parameters
    ss.free()


@pytest.mark.parametrize("tree", ["star5", "cat5", "mixed6", "rand12"])
def test_order1_wavefront_equals_node_by_node_kernel(ctx, tree):
    """The wavefront kernel (list-scheduled node updates, 16 lanes per window) against the quad kernel that walks the nodes
    one by one: identical sweep counts, free energies equal to re-association, both strands, gaps, the M-step objective."""
    from phyfoo_b200 import core, synth
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBayesModel
    from phyfoo_b200.newick import parse_newick
    newick = _trees()[tree]
    S = parse_newick(newick).n_species
    rng = np.random.default_rng(43)
    W = 7
    pwm = synth.random_pwm(W, 1, 91)
    lens = [W, 61, 2, 95, W + 1]
    smp = PhyloSample(random_codes(rng, sum(lens), S, 0.1), np.concatenate([[0], np.cumsum(lens)]))
    fg = PhyloBayesModel(W, 1, newick, ctx); fg.setCondProbs(pwm)
    try:
        res = {}
        for kern in (0, 1):
            ctx.set_option("order1_kernel", kern)
            res[kern] = [fg.getLogProbFor(smp, strand=s, want_sweeps=True) for s in (0, 1)]
            names = [n for n, _ in ctx.last_kernel_times()]
            assert names == ["score_o1w_kernel" if kern else "score_o1_kernel"]
            n = res[kern][0][0].size
            sa, so = np.repeat(np.arange(smp.n_alignments), np.maximum(smp.lengths() - W + 1, 0))[: min(n, 40)], None
            off = smp.device(ctx).window_offsets(W)
            so = (np.arange(n)[: sa.size] - off[sa]).astype(np.int32)
            fe = core.FESet(ctx, smp.device(ctx), fg.device_model(), sa.astype(np.int32), so, np.linspace(0.5, 1.5, sa.size))
            res[kern].append(fe.grad(2, np.log(pwm[2].ravel()) + 0.05, 1e-4))
            fe.eval(2, np.log(pwm[2].ravel()))
        for s in (0, 1):
            assert np.array_equal(res[0][s][1], res[1][s][1])
            assert rel_err(res[1][s][0], res[0][s][0]) < 1e-12
        (f0, g0), (f1, g1) = res[0][2], res[1][2]
        assert abs(f0 - f1) <= 1e-12 * abs(f0) and np.max(np.abs(g0 - g1)) <= 1e-9 * np.max(np.abs(g0)) + 64 * np.finfo(float).eps * abs(f0) / 1e-4
    finally:
        ctx.set_option("order1_kernel", 2)


@pytest.mark.parametrize("qglobal", [0, 8, 3])
@pytest.mark.parametrize("tree,W", [("star5", 4), ("cat5", 7), ("mixed6", 1), ("mixed6", 6), ("rand12", 12)])
def test_order1_node_per_thread_kernel(ctx, tree, W, qglobal):
    """The node-per-thread kernel (net1n.cuh, default): windows without gaps on it, the few windows with an unobserved leaf
    through its device-side list to the wavefront kernel; q rows in shared memory (qglobal = 0) or in global memory behind
    L1 with 8 or 3 warps per SM.  Against the oracle (scores 1e-9 relative, sweep counts exact) and
    against the wavefront kernel alone (same sweep counts, scores equal to re-association); both strands, ragged lengths."""
    from phyfoo_b200 import synth
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBayesModel
    from phyfoo_b200.newick import parse_newick
    newick = _trees()[tree]
    S = parse_newick(newick).n_species
    rng = np.random.default_rng(53 + W)
    pwm = synth.random_pwm(W, 1, 97)
    lens = [W, 150, max(W - 1, 1), 333, W + 1, 77]
    codes = random_codes(rng, sum(lens), S)
    gap_cols = rng.choice(codes.shape[0], 6, replace=False)          # a handful of columns with one unobserved leaf
    codes[gap_cols, rng.integers(0, S, gap_cols.size)] = 4
    smp = PhyloSample(codes, np.concatenate([[0], np.cumsum(lens)]))
    fg = PhyloBayesModel(W, 1, newick, ctx); fg.setCondProbs(pwm)
    om = oracle_fg(newick, pwm, order=1)
    try:
        res = {}
        ctx.set_option("order1_qglobal", qglobal)
        for kern in (2, 1):
            ctx.set_option("order1_kernel", kern)
            res[kern] = [fg.getLogProbFor(smp, strand=s, want_sweeps=True) for s in (0, 1)]
            names = [n for n, _ in ctx.last_kernel_times()]
            assert names == (["score_o1n_kernel", "score_o1w_kernel(windows with gaps)"] if kern == 2 else ["score_o1w_kernel"])
            if kern == 2:
                assert fg.last_stats["sweeps_fg"] == int(res[2][1][1].sum())
        for s in (0, 1):
            ref, rsw = [], []
            for i in range(smp.n_alignments):
                r, k = om.score_windows(smp.getElementAt(i), bool(s))
                ref.append(r); rsw.append(k)
            ref, rsw = np.concatenate(ref), np.concatenate(rsw)
            assert np.array_equal(res[2][s][1], rsw)
            assert rel_err(res[2][s][0], ref) < TOL
            assert np.array_equal(res[2][s][1], res[1][s][1])
            assert rel_err(res[2][s][0], res[1][s][0]) < 1e-12
    finally:
        ctx.set_option("order1_kernel", 2)
        ctx.set_option("order1_qglobal", -1)


@pytest.mark.parametrize("tree", ["cat5", "rand12"])
@pytest.mark.parametrize("gap_rate", [0.0, 0.1])
def test_order1_background_ranges(ctx, tree, gap_rate):
    """PhyloBackground of order 1 (models/PhyloBackground.java:241-305): whole-alignment scores and sub-ranges of at least two
    columns against the oracle's bg_range; single-column ranges and too short alignments are refused."""
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.lib import IllegalArgumentException
    from phyfoo_b200.models import PhyloBackground
    from phyfoo_b200.newick import parse_newick
    newick = _trees()[tree]
    S = parse_newick(newick).n_species
    rng = np.random.default_rng(47)
    lens = [2, 31, 9, 50]
    smp = PhyloSample(random_codes(rng, sum(lens), S, gap_rate), np.concatenate([[0], np.cumsum(lens)]))
    pi = [rng.dirichlet(np.ones(4), size=1), rng.dirichlet(np.ones(4), size=4)]
    bg = PhyloBackground(1, newick, ctx)
    bg.setCondProbs(pi)
    ob = oracle_bg(newick, 1, pi)
    per = bg.getLogProbFor(smp)
    for i, L in enumerate(lens):
        r = ob.bg_range(smp.getElementAt(i), 0, L - 1)
        assert abs(per[i] - r) <= TOL * abs(r), (i, per[i], r)
    for (i, s, e) in ((1, 3, 20), (3, 0, 1), (3, 17, 49), (2, 7, 8)):
        r = ob.bg_range(smp.getElementAt(i), s, e)
        assert abs(bg.getLogProbForRange(smp, i, s, e) - r) <= TOL * abs(r)
    assert bg.getLogProbForRange(smp, 1, 5, 4) == 0.0
    with pytest.raises(ValueError):
        bg.getLogProbForRange(smp, 1, 5, 5)
    short = PhyloSample(random_codes(rng, 4, S), [0, 1, 4])
    with pytest.raises(IllegalArgumentException):
        bg.getLogProbFor(short)


@pytest.mark.parametrize("tree", ["cat5", "rand12"])
@pytest.mark.parametrize("fg_order", [0, 1])
@pytest.mark.parametrize("gap_rate", [0.0, 0.06])
def test_zoops_estep_with_order1_background(ctx, tree, fg_order, gap_rate):
    """SingleHiddenMotifMixture.getNewWeights with a PhyloBackground of order 1 (jstacs SingleHiddenMotifMixture.java:536-540
    with models/PhyloBackground.java:241-305): per offset bg(s, e), bg(s, j-1) and bg(j+W, e) with s = j-1, e = j+W.  The two
    flank ranges are single columns, for which the reference's two-tree net keeps the observation of the previous call in its
    second tree (MeanFieldForBayesNet.java:84) -- always column e here, because bg(s, e) is scored just before; the oracle
    keeps that state like the Java object graph does.  gamma / ll at 1e-9, arg-max calls exact, ragged lengths incl. L = W."""
    from phyfoo_b200 import synth
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture, StrandModel
    from phyfoo_b200.newick import parse_newick
    newick = _trees()[tree]
    S = parse_newick(newick).n_species
    rng = np.random.default_rng(59)
    W = 5
    pwm = synth.random_pwm(W, fg_order, 3)
    lens = [W, W + 1, W + 2, 31, 64]
    smp = PhyloSample(random_codes(rng, sum(lens), S, gap_rate), np.concatenate([[0], np.cumsum(lens)]))
    pi = [rng.dirichlet(np.ones(4), size=1), rng.dirichlet(np.ones(4), size=4)]
    fg = PhyloBayesModel(W, fg_order, newick, ctx); fg.setCondProbs(pwm)
    bg = PhyloBackground(1, newick, ctx); bg.setCondProbs(pi)
    aw = rng.uniform(0.5, 2.0, len(lens))
    for strand in (False, True):
        shm = SingleHiddenMotifMixture(StrandModel(fg, 0.4) if strand else fg, bg, zoops=0.8)
        res = shm.getNewWeights(smp, seqWeights=aw, want_profile=True)
        ref = orc.estep(oracle_fg(newick, pwm, order=fg_order), oracle_bg(newick, 1, pi), smp.col_offsets, smp.codes, np.log([0.8, 0.2]),
                        np.log([0.4, 0.6]) if strand else None, aw)
        assert rel_err(res["profile"], ref["profile"]) < TOL
        assert np.array_equal(res["argmax_off"], ref["argmax_off"]) and np.array_equal(res["argmax_strand"], ref["argmax_strand"])
        assert np.max(np.abs(res["gamma"] - ref["gamma"])) < TOL
        assert abs(res["ll"] - ref["ll"]) <= TOL * abs(ref["ll"]) and np.allclose(res["w"], ref["w"], rtol=1e-9, atol=0)
