"""GPU parity of the M-step path: data selection (bit-exact indices), the fixed-q free-energy objective and its
forward-difference gradient, through the C ABI against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from util import orc, oracle_bg, oracle_fg, random_codes

pytestmark = pytest.mark.gpu

TOL = 1e-9
EPS = 1e-4   # optimizing/AbstractFreeEnergyOptimizer.java:58-60


@pytest.fixture(scope="module")
def ctx():
    from phyfoo_b200 import core
    return core.default_context()


def _oracle_select(smp, W, gamma, t=0.8, q=0.0):
    lens = smp.lengths()
    nwin = np.maximum(lens - W + 1, 0)
    off = np.concatenate([[0], np.cumsum(nwin)])
    sa, so, sw = [], [], []
    for a in range(smp.n_alignments):
        if nwin[a] == 0:
            continue
        idx, ww = orc.select_group(gamma[off[a]:off[a + 1]], t, q)
        sa += [a] * len(idx); so += list(idx); sw += list(ww)
    return np.array(sa, np.int32), np.array(so, np.int32), np.array(sw)


def _grad_tol(g, f0):
    # both sides subtract two sums of magnitude |f|: cancellation noise of f / eps on top of the 1e-9 bar
    return TOL * np.abs(g) + 64 * np.finfo(float).eps * abs(f0) / EPS


def test_select_bit_exact(ctx):
    from phyfoo_b200 import core
    from phyfoo_b200.data import PhyloSample
    rng = np.random.default_rng(31)
    lens = [40, 5, 9, 200, 1500, 12]
    W = 9
    smp = PhyloSample(random_codes(rng, sum(lens), 5), np.concatenate([[0], np.cumsum(lens)]))
    ds = smp.device(ctx)
    nw = ds.windows(W)
    for kind in ("peaked", "flat", "ties", "zeros"):
        if kind == "peaked":
            g = rng.random(nw) ** 12
        elif kind == "flat":
            g = np.full(nw, 0.01) + rng.random(nw) * 1e-6
        elif kind == "ties":
            g = rng.integers(0, 4, nw).astype(np.float64)
        else:
            g = np.zeros(nw); g[::7] = rng.random(g[::7].size)
        for t, q in ((0.8, 0.0), (0.5, 0.3), (1.0, 0.0)):
            sa, so, sw = core.select(ctx, ds, W, g, t, q)
            ra, ro, rw = _oracle_select(smp, W, g, t, q)
            assert np.array_equal(sa, ra) and np.array_equal(so, ro), (kind, t, q)
            assert np.array_equal(sw, rw)


@pytest.mark.parametrize("tree", ["star5", "cat5", "rand12", "mixed6"])
@pytest.mark.parametrize("gap_rate", [0.0, 0.1])
def test_fixed_q_objective_and_gradient(ctx, tree, gap_rate):
    from phyfoo_b200 import core, synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    from test_gpu_parity import _trees, _S
    newick = _trees()[tree]
    S = _S(newick)
    W = 6
    pwm = synth.random_pwm(W, 0, 53)
    smp, _ = synth.make_sample(newick, 40, 70, 7, pwm=pwm, gap_rate=gap_rate, length_jitter=8)
    fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(pwm)
    bg = PhyloBackground(0, newick, ctx)
    res = SingleHiddenMotifMixture(fg, bg, zoops=0.9).getNewWeights(smp)
    ds = smp.device(ctx)
    # selection from the gamma that is still resident on the device == selection from the host copy == oracle
    sa, so, sw = core.select(ctx, ds, W, None)
    sa2, so2, sw2 = core.select(ctx, ds, W, res["gamma"])
    assert np.array_equal(sa, sa2) and np.array_equal(so, so2) and np.array_equal(sw, sw2)
    ra, ro, rw = _oracle_select(smp, W, res["gamma"])
    assert np.array_equal(sa, ra) and np.array_equal(so, ro) and np.array_equal(sw, rw)
    assert 1 <= sa.size <= 40 * 70

    strand = (np.arange(sa.size) % 3 == 0).astype(np.uint8)      # mix in reverse-complement windows
    fe = core.FESet(ctx, ds, fg.device_model(), sa, so, sw, strand)
    assert len(fe) == sa.size
    om = oracle_fg(newick, pwm)
    cols = np.stack([smp.getElementAt(a)[o:o + W] for a, o in zip(sa, so)])
    rc = cols[:, ::-1, :].copy()
    rc = np.where(rc < 4, 3 - rc, 4).astype(np.uint8)
    cols = np.where(strand[:, None, None] == 1, rc, cols)
    ofe = orc.FESet(om, cols, sw)
    rng = np.random.default_rng(3)
    for pos in (0, W - 1, 2):
        lam = np.log(rng.dirichlet(np.ones(4))) + rng.normal()
        f = fe.eval(pos, lam)
        rf = ofe.eval(pos, lam)
        assert abs(f - rf) <= TOL * abs(rf)
        f0, g = fe.grad(pos, lam, EPS)
        r0, rg = ofe.grad(pos, lam, EPS)
        assert abs(f0 - r0) <= TOL * abs(r0)
        assert np.all(np.abs(g - rg) <= _grad_tol(rg, r0)), (g, rg)
        # the model is left at lambda on both sides
        assert np.allclose(fg.device_model().get_pi(pos, 4), om.get_pi(pos), rtol=0, atol=1e-15)
    # all positions at once (MotifParamOptimizer)
    lam_all = np.log(rng.dirichlet(np.ones(4), size=W)).ravel()
    f0, g = fe.grad(-1, lam_all, EPS)
    r0, rg = ofe.grad(-1, lam_all, EPS)
    assert abs(f0 - r0) <= TOL * abs(r0)
    assert np.all(np.abs(g - rg) <= _grad_tol(rg, r0))
    # parameters changed behind the FE set's back are picked up
    fg.setCondProb(pwm[1], 1); om.set_pi(1, pwm[1].ravel())
    lam = np.log(pwm[3].ravel())
    assert abs(fe.eval(3, lam) - ofe.eval(3, lam)) <= TOL * abs(ofe.eval(3, lam))
    # at the parameters the q were optimised for, f = sum_u w_u * score_u  (only position 3 was left elsewhere)
    for t in range(W):
        fe.eval(t, np.log(pwm[t].ravel())); ofe.eval(t, np.log(pwm[t].ravel()))
    direct = sum(w * om.score_window(c)[0] for w, c in zip(sw, cols))
    assert abs(fe.eval(0, np.log(pwm[0].ravel())) - direct) <= 1e-9 * abs(direct)


def test_background_training_objective(ctx):
    """PhyloBackground.train's objective: every column of every alignment, per-alignment weights."""
    from phyfoo_b200 import core, synth
    from phyfoo_b200.models import PhyloBackground
    newick = synth.caterpillar_newick(5)
    smp, _ = synth.make_sample(newick, 6, 30, 5, gap_rate=0.1, length_jitter=5)
    wts = np.random.default_rng(2).random(smp.n_alignments) + 0.2
    bg = PhyloBackground(0, newick, ctx)
    pi0 = np.array([[0.2, 0.3, 0.1, 0.4]])
    bg.setCondProb(pi0, 0)
    fe = core.FESet(ctx, smp.device(ctx), bg.device_model(), aln_weights=wts)
    assert len(fe) == smp.codes.shape[0]
    ob = oracle_bg(newick, 0, [pi0])
    colw = np.repeat(wts, smp.lengths())
    ofe = orc.FESet(ob, smp.codes[:, None, :], colw)
    lam = np.log([0.3, 0.2, 0.25, 0.25])
    f0, g = fe.grad(0, lam, EPS)
    r0, rg = ofe.grad(0, lam, EPS)
    assert abs(f0 - r0) <= TOL * abs(r0)
    assert np.all(np.abs(g - rg) <= _grad_tol(rg, r0))


def test_fe_errors_and_empty(ctx):
    from phyfoo_b200 import core, synth
    from phyfoo_b200.lib import IllegalArgumentException
    from phyfoo_b200.models import PhyloBayesModel
    newick = synth.star_newick(5)
    smp, _ = synth.make_sample(newick, 3, 20, 1)
    fg = PhyloBayesModel(4, 0, newick, ctx)
    ds = smp.device(ctx)
    with pytest.raises(IllegalArgumentException):
        core.FESet(ctx, ds, fg.device_model(), [0], [18], [1.0])       # window leaves its alignment
    with pytest.raises(IllegalArgumentException):
        core.FESet(ctx, ds, fg.device_model(), [3], [0], [1.0])        # no such alignment
    fe = core.FESet(ctx, ds, fg.device_model(), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0))
    assert len(fe) == 0 and fe.eval(0, np.zeros(4)) == 0.0
    fe2 = core.FESet(ctx, ds, fg.device_model(), [0, 1], [0, 3], [1.0, 0.5])
    with pytest.raises(IllegalArgumentException):
        fe2.eval(0, np.zeros(3))
    with pytest.raises(IllegalArgumentException):
        fe2.eval(7, np.zeros(4))


@pytest.mark.parametrize("tree", ["cat5", "rand12", "mixed6"])
@pytest.mark.parametrize("evol", ["alpha", "beta"])
def test_sufficient_statistics_objective(ctx, tree, evol):
    """One device pass -> expected counts; then f(theta) for ANY pi or branch lengths is a host-side dot product.  Checked
    against the faithful device evaluation (fe.eval) and against the oracle with changed pi and changed branch lengths."""
    from phyfoo_b200 import _abi, core, optimize, synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    from test_gpu_parity import _trees
    newick = _trees()[tree]
    W = 5
    pwm = synth.random_pwm(W, 0, 83)
    smp, _ = synth.make_sample(newick, 30, 60, 17, pwm=pwm, gap_rate=0.08, length_jitter=6)
    ev_id, ev_orc = (_abi.EVOL_F81ALPHA, orc.F81ALPHA) if evol == "alpha" else (_abi.EVOL_F81BETA, orc.F81BETA)
    old = PhyloBayesModel.EVOL_MODEL
    PhyloBayesModel.EVOL_MODEL = ev_id
    try:
        fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(pwm)
        bg = PhyloBackground(0, newick, ctx)
        res = SingleHiddenMotifMixture(fg, bg, zoops=0.9).getNewWeights(smp)
        ds = smp.device(ctx)
        sa, so, sw = core.select(ctx, ds, W, None)
        fe = core.FESet(ctx, ds, fg.device_model(), sa, so, sw)
        obj = optimize.SuffStatObjective(fe, fg._pi, fg._branch, ev_id)
    finally:
        PhyloBayesModel.EVOL_MODEL = old
    om = oracle_fg(newick, pwm, evol=ev_orc)
    cols = np.stack([smp.getElementAt(a)[o:o + W] for a, o in zip(sa, so)])
    ofe = orc.FESet(om, cols, sw)
    assert abs(obj.value() - ofe.sum()) <= TOL * abs(ofe.sum())
    rng = np.random.default_rng(7)
    # a new stationary distribution at one position: host dot product == device evaluation == oracle
    lam = np.log(rng.dirichlet(np.ones(4))) + 0.7
    f_host = obj.eval_lambda(2, lam)
    f_dev, f_orc = fe.eval(2, lam), ofe.eval(2, lam)
    assert abs(f_host - f_orc) <= TOL * abs(f_orc) and abs(f_dev - f_orc) <= TOL * abs(f_orc)
    f0, g = obj.grad(lambda z: obj.eval_lambda(2, z), lam, EPS)
    obj.eval_lambda(2, lam)                          # (the gradient leaves the object at its last perturbed point, like the Java)
    r0, rg = ofe.grad(2, lam, EPS)
    assert np.all(np.abs(g - rg) <= _grad_tol(rg, r0))
    # new branch lengths at one position (EdgeOptimizer's parameters): oracle with the same branch lengths set
    N = fg.tree.n_nodes
    x = rng.normal(-1.0, 0.7, N - 1)
    f_edges = obj.eval_edges(3, x)
    from phyfoo_b200 import evolution
    alpha = evolution.branch_from_logit(x)
    for k in range(1, N):
        om.set_branch(3, k, alpha[k - 1])
    assert abs(f_edges - ofe.sum()) <= TOL * abs(ofe.sum())
    # (the reference's parametrisation is not an exact inverse pair: alpha = (e^x + 0.0005) / (1 + e^x) but x0 = logit(alpha))
    assert abs(evolution.logit_from_branch(alpha) - x).max() < 0.05
    # counts are non-negative and each position's root counts sum to the total weight (plus the gap-leaf share)
    assert (obj.counts >= 0).all() and (obj.counts[:, :4].sum(axis=1) >= sw.sum() - 1e-9).all()


def test_library_suffstat_objective_follows_the_device_objective_along_an_optimisation(ctx):
    """phyfoo_ss_* (expected counts from one device pass, evaluations on the host inside the library) against phyfoo_fe_eval
    (a device pass over the selected windows per evaluation): every point a whole position-wise BFGS visits, 1e-12 relative;
    then the one-call M-step (phyfoo_ss_optimize) against the same optimiser driven through phyfoo_fe_eval / phyfoo_fe_grad."""
    from phyfoo_b200 import core, optimize, synth
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    newick = synth.random_binary_newick(7, 4)
    W = 6
    pwm = synth.random_pwm(W, 0, 5)
    smp, _ = synth.make_sample(newick, 60, 80, 6, pwm=pwm, gap_rate=0.03, length_jitter=10)
    fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(synth.random_pwm(W, 0, 77))     # start away from the optimum
    bg = PhyloBackground(0, newick, ctx)
    res = SingleHiddenMotifMixture(fg, bg, zoops=0.9).getNewWeights(smp)
    ds = smp.device(ctx)
    sa, so, sw = core.select(ctx, ds, W, res["gamma"])
    fe = core.FESet(ctx, ds, fg.device_model(), sa, so, sw)
    ss = core.SSObjective(fe)
    visited = []
    for pos in (3, 0, 5):
        def fun(z, pos=pos):
            visited.append((pos, np.array(z)))
            return -ss.eval(pos, z)
        x, f, it, ev = optimize.quasi_newton_bfgs(fun, lambda z: tuple(-v for v in ss.grad(pos, z)), np.log(ss.get_pi(pos)),
                                                  optimize.TC_MOTIF_SEPARATED_POSITIONS())
        ss.eval(pos, x)
        fe.eval(pos, x)                                   # the device model follows
    assert len(visited) > 30
    ss2 = core.SSObjective(fe)                            # counts do not depend on the parameters: same object, fresh copy of the model
    for pos, z in visited[-40:]:
        a, b = ss2.eval(pos, z), fe.eval(pos, z)
        assert abs(a - b) <= 1e-12 * abs(b), (pos, a, b)
    # one-call M-step vs the same optimiser on the device objective
    fg.setCondProbs(synth.random_pwm(W, 0, 77))
    fe2 = core.FESet(ctx, ds, fg.device_model(), sa, so, sw)
    ss3 = core.SSObjective(fe2)
    order = np.arange(W, dtype=np.int32)[::-1].copy()
    fb, fa, ev = ss3.optimize(order)
    for pos in order:
        pos = int(pos)
        lam0 = np.log(fg.getCondProb(pos))
        x, f, it, n = optimize.quasi_newton_bfgs(lambda z: -fe2.eval(pos, z), lambda z: tuple(-v for v in fe2.grad(pos, z)), lam0,
                                                 optimize.TC_MOTIF_SEPARATED_POSITIONS())
        f_dev = fe2.eval(pos, x)
    assert fa > fb and abs(fa - f_dev) <= 1e-7 * abs(fa)
    for f_ in (fe, fe2):
        f_.free()
