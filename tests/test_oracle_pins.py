"""Pins for the CPU oracle beside the executed reference statements of tests/test_reference_statements.py (the reference ships
no tests, golden vectors or fixtures and no JVM exists here):
  * an independently written numpy restatement (oracle/np_restatement.py) agreeing to 1e-12;
  * the closed form for star trees under F81alpha;
  * brute-force enumeration of the joint for exact mode, and the literal frontier elimination;
  * the variational bound  -F <= log P  with equality on star trees.
"""
import math

import numpy as np
import pytest

from util import orc, oracle_bg, oracle_fg, random_codes, rel_err

from oracle import np_restatement as npr
from phyfoo_b200 import synth

STAR = synth.star_newick(5, 0.2)
CAT = synth.caterpillar_newick(5, 0.2)
MIXED = "((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.4):0.25):0.1,F:0.3)"


def _np_net(newick, pwm, order):
    net = npr.Net.motif(newick, len(pwm), order)
    for t, p in enumerate(pwm):
        net.set_pi(t, p)
    net.build_tables()
    return net


@pytest.mark.parametrize("newick", [STAR, CAT, MIXED])
@pytest.mark.parametrize("gap_rate", [0.0, 0.2])
def test_meanfield_order0_matches_numpy_restatement(newick, gap_rate):
    rng = np.random.default_rng(1)
    W = 4
    pwm = synth.random_pwm(W, 0, 7)
    S = npr.parse_newick(newick)[3].__len__()
    m = oracle_fg(newick, pwm)
    net = _np_net(newick, pwm, 0)
    for _ in range(6):
        cols = random_codes(rng, W, S, gap_rate)
        for rc in (False, True):
            v, k = m.score_window(cols, rc)
            v2, k2 = npr.score_window(net, cols, rc)
            assert k == k2
            assert abs(v - v2) <= 1e-12 * abs(v2)


@pytest.mark.parametrize("newick", [STAR, CAT])
@pytest.mark.parametrize("gap_rate", [0.0, 0.25])
def test_meanfield_order1_matches_numpy_restatement(newick, gap_rate):
    rng = np.random.default_rng(2)
    W = 3
    pwm = synth.random_pwm(W, 1, 11)
    m = oracle_fg(newick, pwm, order=1)
    net = _np_net(newick, pwm, 1)
    for _ in range(4):
        cols = random_codes(rng, W, 5, gap_rate)
        v, k = m.score_window(cols, False)
        v2, k2 = npr.score_window(net, cols, False)
        assert k == k2
        assert abs(v - v2) <= 1e-12 * abs(v2)


def test_star_closed_form_equals_exact_and_meanfield():
    rng = np.random.default_rng(3)
    pwm = synth.random_pwm(1, 0, 5)
    alphas = [0.2] * 5
    for mode in (orc.EXACT_PRUNING, orc.EXACT_ELIMINATION, orc.MEANFIELD):
        m = oracle_fg(STAR, pwm, mode=mode)
        for _ in range(10):
            col = random_codes(rng, 1, 5, 0.15)
            v, _ = m.score_window(col)
            ref = npr.star_closed_form(pwm[0][0], alphas, col[0])
            tol = 1e-12 if mode != orc.MEANFIELD else 1e-9   # mean field is exact on a star (one hidden node) up to its stop rule
            if mode == orc.MEANFIELD and (col == 4).any():
                continue   # a gap leaf is a second hidden node: the bound is no longer tight
            assert abs(v - ref) <= tol * abs(ref)


@pytest.mark.parametrize("newick", [CAT, MIXED])
def test_exact_modes_agree_with_bruteforce(newick):
    rng = np.random.default_rng(4)
    S = len(npr.parse_newick(newick)[3])
    pwm = synth.random_pwm(1, 0, 9)
    prune = oracle_fg(newick, pwm, mode=orc.EXACT_PRUNING)
    elim = oracle_fg(newick, pwm, mode=orc.EXACT_ELIMINATION)
    net = _np_net(newick, pwm, 0)
    for _ in range(8):
        col = random_codes(rng, 1, S, 0.2)
        a, _ = prune.score_window(col)
        b, _ = elim.score_window(col)
        c = prune.exact_bruteforce(col)
        d = npr.exact_loglik_order0(net, col)
        assert abs(a - c) <= 1e-12 * abs(c)
        assert abs(b - c) <= 1e-12 * abs(c)
        assert abs(d - c) <= 1e-12 * abs(c)


def test_variational_bound():
    rng = np.random.default_rng(5)
    pwm = synth.random_pwm(3, 0, 13)
    mf = oracle_fg(CAT, pwm)
    ex = oracle_fg(CAT, pwm, mode=orc.EXACT_PRUNING)
    gaps = []
    for _ in range(20):
        cols = random_codes(rng, 3, 5)
        v, k = mf.score_window(cols)
        e, _ = ex.score_window(cols)
        assert v <= e + 1e-9
        assert 2 <= k <= 100
        gaps.append(e - v)
    assert max(gaps) > 1e-3   # the caterpillar bound is not tight


def test_background_order0_and_order1_ranges():
    rng = np.random.default_rng(6)
    codes = random_codes(rng, 12, 5, 0.1)
    bg0 = oracle_bg(CAT, 0)
    net0 = npr.Net.background(CAT, 0)
    net0.build_tables()
    v = bg0.bg_range(codes, 2, 9)
    v2 = npr.bg_range_order0(net0, codes, 2, 9)
    assert abs(v - v2) <= 1e-12 * abs(v2)
    cols = bg0.bg_columns(codes)
    assert abs(cols[2:10].sum() - v) <= 1e-12 * abs(v)
    assert bg0.bg_range(codes, 5, 4) == 0.0
    # order 1: conditional terms of a sliding 2-tree window; the first position comes from the first window
    pi1 = [np.full((1, 4), 0.25), np.random.default_rng(0).dirichlet(np.ones(4), size=4)]
    bg1 = oracle_bg(STAR, 1, pi1)
    net1 = npr.Net.background(STAR, 1)
    for t, p in enumerate(pi1):
        net1.set_pi(t, p)
    net1.build_tables()
    total = 0.0
    start, end = 1, 6
    net1.observe(codes[start:start + 2])
    mf = npr.MeanField(net1, 2); mf.optimize()
    total += mf.free_energy(0, 0)
    for u in range(start, end):
        net1.observe(codes[u:u + 2])
        mf = npr.MeanField(net1, 2); mf.optimize()
        total += mf.free_energy(1, 1)
    v1 = bg1.bg_range(codes, start, end)
    assert abs(v1 + total) <= 1e-12 * abs(total)


def test_zoops_alignment_against_numpy():
    rng = np.random.default_rng(7)
    W, L = 4, 30
    pwm = synth.random_pwm(W, 0, 17)
    fg, bg = oracle_fg(CAT, pwm), oracle_bg(CAT, 0)
    codes = random_codes(rng, L, 5)
    off = np.array([0, L], np.int64)
    logw = np.log([0.7, 0.3])
    for faithful in (False, True):
        res = orc.estep(fg, bg, off, codes, logw, faithful_cost=faithful)
        scores, _ = fg.score_windows(codes)
        cols = bg.bg_columns(codes)
        gam, w0, w1, ll, am = npr.zoops_alignment(scores, cols, W, logw)
        assert rel_err(res["gamma"], gam) < 1e-12
        assert abs(res["w"][0] - w0) < 1e-12 and abs(res["w"][1] - w1) < 1e-12
        assert abs(res["ll"] - ll) <= 1e-12 * abs(ll)
        assert res["argmax_off"][0] == am
        assert abs(res["gamma"].sum() - w0) < 1e-12      # sum of gamma per alignment = p_motif


def test_strand_model_and_threads():
    rng = np.random.default_rng(8)
    W = 3
    pwm = synth.random_pwm(W, 0, 19)
    fg, bg = oracle_fg(STAR, pwm), oracle_bg(STAR, 0)
    lens = [10, 14, 9]
    codes = random_codes(rng, sum(lens), 5, 0.1)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    logw = np.log([0.5, 0.5]); slw = np.log([0.6, 0.4])
    r1 = orc.estep(fg, bg, off, codes, logw, slw, threads=1)
    r4 = orc.estep(fg, bg, off, codes, logw, slw, threads=4)
    for k in ("gamma", "fwd", "rev", "profile"):
        assert np.array_equal(r1[k], r4[k])
    assert r1["ll"] == r4["ll"] and r1["sweeps"] == r4["sweeps"]
    # rev[j] is the forward score of the reverse-complemented window
    j = 2
    win = codes[off[1] + j: off[1] + j + W]
    rc = win[::-1].copy(); rc[rc < 4] = 3 - rc[rc < 4]
    v, _ = fg.score_window(rc)
    nwin0 = lens[0] - W + 1
    assert r1["rev"][nwin0 + j] == v
    mot = np.logaddexp(slw[0] + r1["fwd"], slw[1] + r1["rev"])
    bgc = bg.bg_columns(codes[off[1]:off[2]])
    a = -math.log(lens[1] - W + 1) + mot[nwin0 + j] - bgc[j:j + W].sum()
    assert abs(r1["profile"][nwin0 + j] - a) < 1e-12 * abs(a)


def test_selector_and_lambda():
    w = np.array([0.05, 0.5, 0.1, 0.3, 0.05])
    idx, ww = orc.select_group(w, 0.8, 0.0)
    assert list(idx) == [1, 3, 2]           # 0.5, 0.8 (not > 0.8), 0.9 -> stop after exceeding
    assert np.allclose(ww, [0.5, 0.3, 0.1])
    idx, _ = orc.select_group(np.array([1.0, 0.0, 0.0]), 0.8, 0.0)
    assert list(idx) == [0]
    lam = np.log([0.1, 0.2, 0.3, 0.4, 0.25, 0.25, 0.25, 0.25]) + 3.0
    assert np.allclose(orc.lambda2pi(lam), [0.1, 0.2, 0.3, 0.4, 0.25, 0.25, 0.25, 0.25])


def test_fixed_q_objective_and_forward_difference():
    rng = np.random.default_rng(9)
    W = 3
    pwm = synth.random_pwm(W, 0, 23)
    m = oracle_fg(CAT, pwm)
    n = 5
    cols = np.stack([random_codes(rng, W, 5, 0.1) for _ in range(n)])
    wts = rng.random(n)
    fs = orc.FESet(m, cols, wts)
    direct = sum(wts[u] * m.score_window(cols[u])[0] for u in range(n))
    assert abs(fs.sum() - direct) <= 1e-12 * abs(direct)    # at the trained theta, -sum w F == sum w score
    lam = np.log(pwm[1].ravel())
    f0, g = fs.grad(1, lam, 1e-4)
    assert abs(f0 - direct) <= 1e-10 * abs(direct)
    for i in range(4):
        lp = lam.copy(); lp[i] += 1e-4
        assert abs(g[i] - (fs.eval(1, lp) - f0) / 1e-4) < 1e-9
    fs.eval(1, lam)


def test_encoding_and_packing():
    codes = orc.encode_alignment(["ACgt-", "acGT-", "-----"])
    assert codes.tolist() == [[0, 0, 4], [1, 1, 4], [2, 2, 4], [3, 3, 4], [4, 4, 4]]
    assert orc.drop_allgap_columns(codes).shape[0] == 4
    with pytest.raises(ValueError):
        orc.encode_alignment(["ACNT"])
    rng = np.random.default_rng(10)
    c = random_codes(rng, 7, 20, 0.3)
    bases, gaps = orc.pack_columns(c)
    assert bases.shape == (2, 7) and gaps.shape == (1, 7)
    for col in range(7):
        for s in range(20):
            g = (gaps[s // 32, col] >> (s % 32)) & 1
            b = (bases[s // 16, col] >> (2 * (s % 16))) & 3
            assert (4 if g else b) == c[col, s]
            if g:
                assert b == 0


def test_term_counter_matches_closed_forms():
    """The flop model bench.py uses for the roofline equals the oracle's instrumented counter."""
    import bench
    rng = np.random.default_rng(12)
    for newick, S, I in ((STAR, 5, 1), (CAT, 5, 4), (synth.random_binary_newick(12, 3), 12, 11)):
        for W in (1, 3, 5):
            c0 = oracle_fg(newick, synth.random_pwm(W, 0, 1)).count_terms(random_codes(rng, W, S))
            upd, fe = bench.flops_order0(I, S)
            assert (c0["upd_flops"], c0["fe_flops"]) == (W * upd, W * fe)
            assert c0["upd_transc"] == 4 * I * W                      # 4 exp per hidden node and sweep
            c1 = oracle_fg(newick, synth.random_pwm(W, 1, 1), order=1).count_terms(random_codes(rng, W, S))
            assert (c1["upd_flops"], c1["fe_flops"]) == bench.flops_order1(I, S, W)


@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("gap_rate", [0.0, 0.1])
def test_background_of_higher_order_matches_numpy_restatement(order, gap_rate):
    """PhyloBackground of order k >= 2 (models/PhyloBackground.java:241-305, structure :381-398): the oracle's range value
    against the independently written numpy restatement -- the first k positions from the first (k+1)-tree window, every further
    position the last tree of its own window, each window's mean field optimised from the uniform start."""
    rng = np.random.default_rng(30 + order)
    codes = random_codes(rng, 11, 5, gap_rate)
    pis = [rng.dirichlet(np.ones(4), size=4 ** min(order, t)) for t in range(order + 1)]
    bg = oracle_bg(CAT, order, pis)
    net = npr.Net.background(CAT, order)
    for t, p in enumerate(pis):
        net.set_pi(t, p)
    net.build_tables()
    T = order + 1
    for start, end in ((0, 10), (2, 2 + order), (3, 9)):
        total = 0.0
        net.observe(codes[start:start + T])
        mf = npr.MeanField(net, T); mf.optimize()
        for i in range(order):
            total += mf.free_energy(i, i)
        for u in range(start, end - order + 1):
            net.observe(codes[u:u + T])
            mf = npr.MeanField(net, T); mf.optimize()
            total += mf.free_energy(order, order)
        v = bg.bg_range(codes, start, end)
        assert abs(v + total) <= 1e-12 * abs(total), (start, end, v, -total)


@pytest.mark.parametrize("order", [1, 2])
def test_short_background_ranges_keep_the_previous_observation(order):
    """A range shorter than order + 1 columns observes only its own columns; the trailing trees of the shared net keep what the
    call before left there (algorithm/MeanFieldForBayesNet.java:84), while the update still sweeps every
    hidden node and the stop rule and the returned free energy span the range's trees only (models/PhyloBackground.java:
    260-283).  The oracle keeps that state in its model object; here the same sequence of calls -- the three calls of one motif
    offset of the ZOOPS E-step (jstacs SingleHiddenMotifMixture.java:536-540) -- on the numpy restatement with an explicit
    observation map."""
    rng = np.random.default_rng(50 + order)
    L, W, j = 14, 4, 5
    codes = random_codes(rng, L, 5, 0.08)
    pis = [rng.dirichlet(np.ones(4), size=4 ** min(order, t)) for t in range(order + 1)]
    bg = oracle_bg(CAT, order, pis)
    net = npr.Net.background(CAT, order)
    for t, p in enumerate(pis):
        net.set_pi(t, p)
    net.build_tables()
    T = order + 1
    leaves = net.leaves

    def observe_trees(first_col, n_trees):                  # VirtualTree.setObservation for the first n_trees trees only
        for t in range(n_trees):
            for s, k in enumerate(leaves):
                c = int(codes[first_col + t][s])
                if c < 4:
                    net.obs[(t, k)] = c
                else:
                    net.obs.pop((t, k), None)

    def np_range(start, end):
        if end < start:
            return 0.0
        n = end - start + 1
        first_len = min(T, n)
        total = 0.0
        observe_trees(start, first_len)
        mf = npr.MeanField(net, first_len); mf.optimize()
        pos = start
        for i in range(order):
            if pos > end:
                break
            total += mf.free_energy(i, i); pos += 1
        u = start
        while pos <= end:
            wl = first_len if u == start else T
            observe_trees(u, wl)
            mf = npr.MeanField(net, wl); mf.optimize()
            total += mf.free_energy(order, order); pos += 1; u += 1
        return -total

    net.obs = {}
    s, e = max(j - order, 0), min(j + W + order - 1, L - 1)
    for (a, b) in ((0, L - 1), (s, e), (s, j - 1), (j + W, e)):           # logPSeq, then the three ranges of offset j
        v, ref = np_range(a, b), bg.bg_range(codes, a, b)
        assert abs(v - ref) <= 1e-12 * abs(ref), ((a, b), v, ref)
    # the value of the short range does depend on what came before: the same range after a different call differs
    fresh = oracle_bg(CAT, order, pis)
    fresh.bg_range(codes, 0, order)                                       # leaves the columns 0 .. order in the trees
    assert abs(fresh.bg_range(codes, s, j - 1) - bg.bg_range(codes, s, j - 1)) > 1e-9 or order == 0
