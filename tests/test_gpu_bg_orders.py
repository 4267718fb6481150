"""PhyloBackground of order k >= 2 (models/PhyloBackground.java:241-305, structure :381-398) through the generic net kernel
(csrc/netg.cuh): range terms, whole-alignment scores and the ZOOPS E-step against the oracle -- values at 1e-9 relative, mean-field
sweep counts and arg-max calls exact; order 1 through the same kernel against the order-1 network kernel.  The flank ranges of
the E-step are shorter than k + 1 columns: their trailing trees hold what the reference's previous call left there
(MeanFieldForBayesNet.java:84), which the oracle reproduces by keeping one model object like the Java does
(tests/test_emul_host.py checks the same bookkeeping on the host)."""
import numpy as np
import pytest

from util import orc, oracle_bg, oracle_fg, random_codes, rel_err
from test_gpu_parity import _trees

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    from phyfoo_b200 import core
    return core.default_context()


def _pis(rng, order):
    return [rng.dirichlet(np.ones(4), size=4 ** min(order, t)) for t in range(order + 1)]


@pytest.mark.parametrize("tree,order", [("cat5", 2), ("mixed6", 3), ("rand12", 2), ("star5", 1)])
@pytest.mark.parametrize("gap_rate", [0.0, 0.1])
def test_background_range_terms_of_any_order(ctx, tree, order, gap_rate):
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.lib import IllegalArgumentException
    from phyfoo_b200.models import PhyloBackground
    from phyfoo_b200.newick import parse_newick
    newick = _trees()[tree]
    S = parse_newick(newick).n_species
    rng = np.random.default_rng(61 + order)
    lens = [order + 1, 23, order + 2, 40]
    smp = PhyloSample(random_codes(rng, sum(lens), S, gap_rate), np.concatenate([[0], np.cumsum(lens)]))
    pi = _pis(rng, order)
    bg = PhyloBackground(order, newick, ctx); bg.setCondProbs(pi)
    ob = oracle_bg(newick, order, pi)
    PhyloBackground.GENERIC_NET = True
    try:
        per = bg.getLogProbFor(smp)
        cond, first, sw = bg.getRangeTerms(smp, want_sweeps=True)
        assert [n for n, _ in ctx.last_kernel_times()] == ["score_netg_kernel"]
    finally:
        PhyloBackground.GENERIC_NET = False
    for i, L in enumerate(lens):
        r = ob.bg_range(smp.getElementAt(i), 0, L - 1)
        assert abs(per[i] - r) <= TOL * abs(r), (i, per[i], r)
    for (i, s, e) in ((1, 3, 20), (3, 0, order), (3, 17, 39), (2, 0, order + 1), (1, 22 - order, 22)):
        r = ob.bg_range(smp.getElementAt(i), s, e)
        assert abs(bg.getLogProbForRange(smp, i, s, e) - r) <= TOL * abs(r)
    # sweeps of every (order + 1)-column window: the oracle's mean field on that window alone
    w = 0
    for i, L in enumerate(lens):
        a = smp.getElementAt(i)
        for u in range(L - order):
            ob.reset_stats()
            ob.bg_range(a[u:u + order + 1], 0, order)
            # bg_range optimises the first window twice (models/PhyloBackground.java:271-273 and :284-287)
            assert 2 * sw[w] == ob.sweeps_total(), (i, u)
            w += 1
    assert w == sw.size
    assert bg.getLogProbForRange(smp, 1, 5, 4) == 0.0
    with pytest.raises(ValueError):
        bg.getLogProbForRange(smp, 1, 5, 5 + order - 1)
    if order == 1:                                             # the order-1 network kernel gives the same terms
        c1, f1 = bg.getRangeTerms(smp)
        assert rel_err(c1, cond) < 1e-12 and rel_err(f1, first) < 1e-12
    short = PhyloSample(random_codes(rng, order + 3, S), [0, order, order + 3])
    with pytest.raises(IllegalArgumentException):
        bg.getLogProbFor(short)


@pytest.mark.parametrize("tree,order", [("cat5", 2), ("mixed6", 3), ("cat5", 1)])
@pytest.mark.parametrize("fg_order", [0, 1])
@pytest.mark.parametrize("gap_rate", [0.0, 0.06])
def test_zoops_estep_with_background_of_any_order(ctx, tree, order, fg_order, gap_rate):
    from phyfoo_b200 import synth
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture, StrandModel
    from phyfoo_b200.newick import parse_newick
    newick = _trees()[tree]
    S = parse_newick(newick).n_species
    rng = np.random.default_rng(67 + order)
    W = 5
    pwm = synth.random_pwm(W, fg_order, 3)
    lens = [W, W + 1, W + 2, W + order, W + 2 * order + 1, 31]
    smp = PhyloSample(random_codes(rng, sum(lens), S, gap_rate), np.concatenate([[0], np.cumsum(lens)]))
    pi = _pis(rng, order)
    fg = PhyloBayesModel(W, fg_order, newick, ctx); fg.setCondProbs(pwm)
    bg = PhyloBackground(order, newick, ctx); bg.setCondProbs(pi)
    aw = rng.uniform(0.5, 2.0, len(lens))
    ctx.set_option("bg_generic", 1)
    try:
        for strand in (False, True):
            shm = SingleHiddenMotifMixture(StrandModel(fg, 0.4) if strand else fg, bg, zoops=0.8)
            res = shm.getNewWeights(smp, seqWeights=aw, want_profile=True)
            assert "score_netg_kernel(left flanks)" in [n for n, _ in ctx.last_kernel_times()]
            ref = orc.estep(oracle_fg(newick, pwm, order=fg_order), oracle_bg(newick, order, pi), smp.col_offsets, smp.codes,
                            np.log([0.8, 0.2]), np.log([0.4, 0.6]) if strand else None, aw)
            assert rel_err(res["profile"], ref["profile"]) < TOL
            assert np.array_equal(res["argmax_off"], ref["argmax_off"]) and np.array_equal(res["argmax_strand"], ref["argmax_strand"])
            assert np.max(np.abs(res["gamma"] - ref["gamma"])) < TOL
            assert abs(res["ll"] - ref["ll"]) <= TOL * abs(ref["ll"]) and np.allclose(res["w"], ref["w"], rtol=1e-9, atol=0)
    finally:
        ctx.set_option("bg_generic", 0)


@pytest.mark.parametrize("tree,order", [("cat5", 1), ("mixed6", 2)])
def test_background_training_of_any_order(ctx, tree, order):
    """PhyloBackground.train for order k >= 1 (models/PhyloBackground.java:111-176) on the expected counts of the generic net
    (phyfoo_ss_create_bg): the objective values the optimisers report equal the oracle's fixed-q objective under the same
    parameters -- q fixed under the starting model, window p weighted with the alignment that holds COLUMN p (the reference's
    indexing) --, training raises the objective and the log-probability of the data, and the device model follows."""
    from phyfoo_b200 import core
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBackground
    from phyfoo_b200.newick import parse_newick
    newick = _trees()[tree]
    tr = parse_newick(newick)
    rng = np.random.default_rng(83 + order)
    T = order + 1
    lens = [25, 40, T, 31]
    smp = PhyloSample(random_codes(rng, sum(lens), tr.n_species, 0.05), np.concatenate([[0], np.cumsum(lens)]))
    wts = rng.uniform(0.5, 2.0, len(lens))
    pi0 = _pis(rng, order)
    bg = PhyloBackground(order, newick, ctx); bg.setCondProbs(pi0)
    ll0 = bg.getLogProbFor(smp).sum()
    # the oracle's objective: every (k+1)-column window of every alignment, weights by the reference's (mis)indexing
    wins, ww = [], []
    col_w = np.repeat(wts, lens)
    for i, L in enumerate(lens):
        a = smp.getElementAt(i)
        for u in range(L - order):
            wins.append(a[u:u + T]); ww.append(col_w[len(ww)])
    ob = oracle_bg(newick, order, pi0)
    ofe = orc.FESet(ob, np.stack(wins), np.array(ww))
    # 1. the objective and its gradient through the C ABI
    obj = core.SSObjective(background=(ctx, smp.device(ctx), bg.device_model(), wts))
    assert [n for n, _ in ctx.last_kernel_times()] == ["netg_count_kernel"]
    assert abs(obj.value() - ofe.sum()) <= TOL * abs(ofe.sum())
    t = order                                                           # the position with all 4^k context rows
    lam = rng.normal(size=4 ** order * 4)
    f, g = obj.grad(t, lam, 1e-4)
    r0, rg = ofe.grad(t, lam, 1e-4)
    assert abs(f - r0) <= TOL * abs(r0)
    assert np.all(np.abs(g - rg) <= TOL * np.abs(rg) + 64 * np.finfo(float).eps * abs(r0) / 1e-4)
    with pytest.raises(ValueError):
        obj.eval(t, np.zeros(4)) if order else None
    obj.free()
    ofe.eval(t, np.log(pi0[t]).ravel())
    # 2. training: pi position by position, then the branch lengths
    PhyloBackground.EDGE_LEARNING = True
    try:
        st = bg.train(smp, wts)
    finally:
        PhyloBackground.EDGE_LEARNING = False
    assert st["edges_f_after"] >= st["edges_f_before"] >= st["f_before"] and st["edges_f_after"] > st["f_before"]
    assert abs(st["f_before"] - ofe.sum()) <= TOL * abs(ofe.sum())
    for tt in range(T):
        ofe.eval(tt, np.log(bg.getCondProb(tt)))
    assert abs(ofe.sum() - st["edges_f_before"]) <= TOL * abs(st["edges_f_before"])
    for tt in range(T):
        for k in range(1, tr.n_nodes):
            ob.set_branch(tt, k, bg._branch[tt, k])
    assert abs(ofe.sum() - st["edges_f_after"]) <= TOL * abs(st["edges_f_after"])
    # 3. the device model scores with the trained parameters (a fresh oracle under them agrees) and the data became likelier
    ob2 = oracle_bg(newick, order, [bg.getCondProb(tt) for tt in range(T)])
    for tt in range(T):
        for k in range(1, tr.n_nodes):
            ob2.set_branch(tt, k, bg._branch[tt, k])
    per = bg.getLogProbFor(smp)
    for i, L in enumerate(lens):
        r = ob2.bg_range(smp.getElementAt(i), 0, L - 1)
        assert abs(per[i] - r) <= TOL * abs(r)
    assert per.sum() > ll0
