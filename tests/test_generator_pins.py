"""The oracle held against the only outputs of reference code that exist in this image: the bundled synthetic data
(/root/reference/data/synthetic_data), which the reference's own sampler wrote (bayesNet/BayesNetHandler.java:114-142 through
PhyloBackground / PhyloBayesModel.emitSample) from the parameters its description*.txt files state.  tests/golden/
generator_pin.npz holds the column-pattern histograms of those files (scripts/make_generator_pin_fixture.py); the tests ask
whether the oracle's exact column likelihood under the stated tree and distribution is the distribution the columns were drawn
from (Pearson's chi-square over the 4^5 patterns, 1023 degrees of freedom, ~260 000 columns per file).

This pins, against reference-generated data and to that statistical resolution: the Newick parser and species order, the
meaning of a branch length under F81alpha (a substitution PROBABILITY, evolution/FS81alpha.java:76-92: +-2.5 % is resolved),
the CPF look-up order and the exact likelihood -- not the mean-field numbers bit for bit (nothing can: no JVM), but the model
they approximate.  It is a statistical pin of the model and of the object graph (tree, species order, tables); the bit-level pin of the
arithmetic is tests/test_reference_statements.py, which executes the reference's own statements.
"""
import os

import numpy as np
import pytest

from util import orc, oracle_fg

from oracle import np_restatement as npr

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generator_pin.npz"))
PATTERNS = np.array([[(i // 4 ** s) % 4 for s in range(5)] for i in range(1024)], dtype=np.uint8)   # index = sum code[s] 4^s
UNIFORM = (0.25, 0.25, 0.25, 0.25)
DOF = 1023
ACCEPT = DOF + 5 * np.sqrt(2 * DOF)          # mean + 5 sigma of a chi-square with 1023 degrees of freedom = 1249
_cache = {}


def log_table(newick, pi=UNIFORM, mode=orc.EXACT_PRUNING, evol=orc.F81ALPHA, hky_ratio=0.0):
    """log P(pattern) of all 1024 gap-free columns of five species under a one-position model."""
    key = (newick, tuple(pi), mode, evol, hky_ratio)
    if key not in _cache:
        m = oracle_fg(newick, np.array([[pi]]), mode=mode, evol=evol, hky_ratio=hky_ratio)
        with np.errstate(all="ignore"):
            _cache[key] = np.array([m.score_window(PATTERNS[i:i + 1])[0] for i in range(1024)])
    return _cache[key]


def chi_square(hist, logp):
    expected = hist.sum() * np.exp(logp)
    return float(((hist - expected) ** 2 / expected).sum())


FILES = [("bg_tree1", "tree1"), ("bg_tree2", "tree2"), ("bg_tree3", "tree3"), ("bg_star", "star"),
         ("flanks_tree1", "tree1"), ("flanks_ratio", "star")]


@pytest.mark.parametrize("data,tree", FILES)
def test_bundled_backgrounds_are_draws_from_the_exact_likelihood(data, tree):
    hist = G[data]
    assert hist.sum() > 250_000
    p = log_table(str(G[tree]))
    assert abs(np.exp(p).sum() - 1) < 1e-12
    assert chi_square(hist, p) < ACCEPT
    # the same table from the literal node elimination and from the independent numpy restatement
    assert np.allclose(log_table(str(G[tree]), mode=orc.EXACT_ELIMINATION), p, rtol=1e-12, atol=0)
    net = npr.Net.motif(str(G[tree]), 1, 0)
    net.set_pi(0, np.array(UNIFORM))
    net.build_tables()
    for i in (0, 27, 400, 1023):
        assert abs(npr.exact_loglik_order0(net, PATTERNS[i:i + 1]) - p[i]) <= 1e-12 * abs(p[i])
    # the test has the power to tell the trees apart: every other topology of the bundle is rejected by a wide margin
    for other in ("tree1", "tree2", "tree3", "star"):
        if other != tree:
            assert chi_square(hist, log_table(str(G[other]))) > 20 * ACCEPT


def test_branch_length_is_a_substitution_probability():
    """F81alpha (the reference's default, evolution/FS81alpha.java:76-92): a child keeps its parent's base with probability
    1 - length, else draws from pi.  Lengths 5 % off the stated 0.2 are rejected, and so is the rate reading of the same
    Newick string (FS81beta.java:81-103, alpha = 1 - exp(-length * beta))."""
    t1 = str(G["tree1"])
    for data in ("bg_tree1", "flanks_tree1"):
        got = {s: chi_square(G[data], log_table(t1.replace("0.2", s))) for s in ("0.19", "0.2", "0.21")}
        assert got["0.2"] < ACCEPT < min(got["0.19"], got["0.21"])
    assert chi_square(G["bg_tree1"], log_table(t1, evol=orc.F81BETA)) > 5 * ACCEPT


def test_mean_field_scores_prefer_the_generating_tree():
    """The default scoring mode (mean-field free energy, models/PhyloBayesModel.java:222-267) on the bundled backgrounds:
    a lower bound of the exact likelihood for every pattern, and as a model-selection score it ranks the tree each file was
    generated from first."""
    trees = ("tree1", "tree2", "tree3", "star")
    for data, truth in (("bg_tree1", "tree1"), ("bg_tree2", "tree2"), ("bg_tree3", "tree3"), ("bg_star", "star")):
        hist = G[data]
        per_column = {}
        for t in trees:
            mf = log_table(str(G[t]), mode=orc.MEANFIELD)
            assert (mf <= log_table(str(G[t])) + 1e-9).all()
            per_column[t] = float((hist * mf).sum() / hist.sum())
        assert max(per_column, key=per_column.get) == truth


def test_planted_sites_on_the_star_tree_follow_the_motif_model():
    """ratio_positive_negative_testing/data_1.0.fg: 750 sites of a 15-column motif on the star tree at the annotated MOTIF_POS
    (0-based first column, io/PhyloSample.java annotations).  Per motif column the mean exact log-likelihood of the 750 site
    columns under the PWM row of description.txt is within 4 standard errors of its expectation; under the neighbouring row
    (a MOTIF_POS convention off by one) it is far outside for most columns."""
    pwm = G["pwm_ratio"] / G["pwm_ratio"].sum(axis=1, keepdims=True)
    star = str(G["star"])
    tables = [log_table(star, pi=tuple(pwm[p])) for p in range(15)]

    def z(p, row):
        lp = tables[row]
        q = np.exp(lp)
        mean = (q * lp).sum()
        var = (q * lp * lp).sum() - mean * mean
        h = G["sites_ratio"][p]
        return ((h * lp).sum() / h.sum() - mean) / np.sqrt(var / h.sum())

    assert all(G["sites_ratio"][p].sum() == 750 for p in range(15))
    assert max(abs(z(p, p)) for p in range(15)) < 4
    assert sum(abs(z(p, (p + 1) % 15)) > 4 for p in range(15)) >= 10


def _pair_keep(hist, i, j, pi):
    """(P(species i and j agree) - sum pi^2) / (1 - sum pi^2): the probability that no redraw separates the two leaves."""
    same = hist[(PATTERNS[:, i] == PATTERNS[:, j])].sum() / hist.sum()
    s2 = float((np.asarray(pi) ** 2).sum())
    return (same - s2) / (1 - s2)


def test_planted_sites_of_the_tree_data_sets_are_not_draws_from_the_described_model():
    """A property of the bundled data, recorded so that nobody tunes the oracle to it: in trees/data_tree1_1.0.fg the flanks
    follow the caterpillar tree (previous tests) but in the planted sites SPECIES_0 is not the sibling of SPECIES_1 -- it agrees
    with every other species as if its four branches to the root were drawn on their own (keep probability 0.8^4 times the
    other leaf's path to the root), while species 1..4 follow the tree.  The sites of the star-tree files are unaffected."""
    pwm = G["pwm_tree1"] / G["pwm_tree1"].sum(axis=1, keepdims=True)
    keep = np.zeros((5, 5))
    for i in range(5):
        for j in range(i + 1, 5):
            keep[i, j] = np.mean([_pair_keep(G["sites_tree1"][p], i, j, pwm[p]) for p in range(10)])
    depth = [4, 4, 3, 2, 1]                                    # branches between a leaf and the root of the caterpillar
    expected = {(1, 2): 0.8 ** 3, (1, 3): 0.8 ** 4, (1, 4): 0.8 ** 5, (2, 3): 0.8 ** 3, (2, 4): 0.8 ** 4, (3, 4): 0.8 ** 3}
    for (i, j), e in expected.items():
        assert abs(keep[i, j] - e) < 0.03                      # species 1..4: the described tree
    assert keep[0, 1] < 0.25                                   # described: 0.8^2 = 0.64
    for j in range(1, 5):
        assert abs(keep[0, j] - 0.8 ** 4 * 0.8 ** depth[j]) < 0.04
    # the flanks of the same file: SPECIES_0 and SPECIES_1 are siblings
    assert abs(_pair_keep(G["flanks_tree1"], 0, 1, UNIFORM) - 0.64) < 0.01


def _hky_linear(alpha, beta, pi=UNIFORM):
    """evolution/HKY.java:77-106 with the ratio the constructor was meant to store (beta = alpha * ratio)."""
    a, b = [alpha * x for x in pi], [beta * x for x in pi]
    return np.array([[1 - (a[2] + b[3] + b[1]), b[1], a[2], b[3]],
                     [b[0], 1 - (a[3] + b[0] + b[2]), b[2], a[3]],
                     [a[0], b[1], 1 - (a[0] + b[3] + b[1]), b[3]],
                     [b[0], a[1], b[2], 1 - (a[1] + b[0] + b[2])]])


def test_hky_as_implemented_gives_the_bundled_hky_data_probability_zero():
    """The bundled HKY files were written with the transition / transversion ratio in effect (star tree: P(column) =
    sum_r pi_r prod_s T[r][x_s] with beta = ratio * alpha fits both files).  The class as it stands never stores the ratio
    (evolution/HKY.java:32), so beta = 0: the faithful tables -- what the oracle restates and the library returns, flagged by
    phyfoo_model_is_degenerate -- give every column that holds a transversion probability zero, two thirds of the reference's
    own HKY columns."""
    for data, alpha, ratio in (("bg_hky2", 0.2, 2.0), ("bg_hky05", 0.4, 0.5)):      # HKY/description.txt:6-7,50-52
        T = _hky_linear(alpha, alpha * ratio)
        assert np.allclose(T.sum(axis=1), 1)
        p = np.array([sum(0.25 * np.prod(T[r][PATTERNS[i]]) for r in range(4)) for i in range(1024)])
        assert chi_square(G[data], np.log(p)) < ACCEPT
        # the opt-in model of the library and the oracle (PHYFOO_EVOL_HKY_RATIO) is that distribution
        tree = str(G["star04" if alpha == 0.4 else "star"])
        opt_in = log_table(tree, evol=orc.HKY_RATIO, hky_ratio=ratio)
        assert np.max(np.abs(opt_in - np.log(p)) / np.abs(np.log(p))) < 1e-12
        assert chi_square(G[data], log_table(tree, evol=orc.HKY_RATIO, hky_ratio=1.0)) > 20 * ACCEPT     # ratio 1 = F81
    # ... and on the three trees of HKY+trees/ (internal nodes: tables of every level)
    for data, tree, ratio in (("bg_tree1_hky2", "tree1", 2.0), ("bg_tree1_hky05", "tree1", 0.5), ("bg_tree3_hky2", "tree3", 2.0),
                              ("bg_tree2_hky05", "tree2", 0.5)):
        newick = str(G[tree]) if ratio == 2.0 else str(G[tree]).replace("0.2", "0.4")      # description_tree*.txt:5,50
        assert chi_square(G[data], log_table(newick, evol=orc.HKY_RATIO, hky_ratio=ratio)) < ACCEPT
        assert chi_square(G[data], log_table(newick, evol=orc.HKY_RATIO, hky_ratio=1.0)) > 20 * ACCEPT
    as_implemented = log_table(str(G["star"]), evol=orc.HKY)
    zero = ~np.isfinite(as_implemented)
    assert zero.sum() == 960 and abs(np.exp(as_implemented[~zero]).sum() - 1) < 1e-12
    assert G["bg_hky2"][zero].sum() > 0.6 * G["bg_hky2"].sum()


@pytest.mark.gpu
@pytest.mark.parametrize("data,tree,ratio", [("bg_tree1", "tree1", 0), ("bg_tree3", "tree3", 0), ("bg_star", "star", 0),
                                             ("bg_hky2", "star", 2.0), ("bg_hky05", "star04", 0.5)])
def test_device_scores_against_the_bundled_backgrounds(data, tree, ratio):
    """The same question put to libphyfoo.so: the exact-mode scores of the 1024 patterns (one alignment of 1024 columns, width
    1, through phyfoo_score_windows) are the distribution the reference's sampler drew the file from; the default mean-field
    scores of the same columns are the oracle's and stay below them."""
    from phyfoo_b200 import _abi, core
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBayesModel
    ctx = core.default_context()
    newick = str(G[tree])
    smp = PhyloSample(PATTERNS.copy(), [0, 1024])
    evol = dict(evol=orc.HKY_RATIO, hky_ratio=ratio) if ratio else {}

    class Model(PhyloBayesModel):                     # the reference selects the evolution model through a static (EvolModel.java:7)
        EVOL_MODEL = _abi.EVOL_HKY_RATIO if ratio else _abi.EVOL_F81ALPHA
        HKY_RATIO = float(ratio)

    m = Model(1, 0, newick, ctx)
    m.setCondProbs(np.array([[UNIFORM]]))
    mean_field = core.score_windows(ctx, smp.device(ctx), m.device_model(), 0)[0]
    ref = log_table(newick, mode=orc.MEANFIELD, **evol)
    assert np.max(np.abs(mean_field - ref) / np.abs(ref)) < 1e-9
    Model.CALC_LOGLIKELIHOOD = True                   # the reference's static switch (models/PhyloBayesModel.java:54); device_model()
    exact = core.score_windows(ctx, smp.device(ctx), m.device_model(), 0)[0]      # follows it with phyfoo_model_set_mode
    assert exact.shape == (1024,)
    assert np.max(np.abs(exact - log_table(newick, **evol)) / np.abs(log_table(newick, **evol))) < 1e-12
    assert abs(np.exp(exact).sum() - 1) < 1e-12
    assert chi_square(G[data], exact) < ACCEPT
    assert (mean_field <= exact + 1e-9).all()


# ---- the ZOOPS posterior on reference-generated positives ----------------------------------------------------------------
R = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ratio_star_first400.npz"))


def _calibration(gamma, argmax_off, W=15):
    """Two identities of an exact posterior over the planted position J of an alignment, in expectation over data drawn from
    the model: E[gamma_J] = E[sum_j gamma_j^2] and P(arg-max = J) = E[max_j gamma_j].  Returns their z statistics over the
    alignments, the hit rate and the mean posterior of the annotated position."""
    lengths = np.diff(R["col_offsets"])
    off = np.concatenate([[0], np.cumsum(lengths - W + 1)])
    pos, n = R["motif_pos"], lengths.size
    g_true = np.array([gamma[off[i] + pos[i]] for i in range(n)])
    g_sq = np.array([(gamma[off[i]:off[i + 1]] ** 2).sum() for i in range(n)])
    g_max = np.array([gamma[off[i]:off[i + 1]].max() for i in range(n)])
    hit = (argmax_off == pos).astype(np.float64)
    z = lambda d: float(d.mean() / (d.std(ddof=1) / np.sqrt(n)))
    return z(g_true - g_sq), z(hit - g_max), float(hit.mean()), float(g_true.mean())


def _oracle_estep(newick, pwm):
    from util import oracle_bg
    fg, bg = oracle_fg(newick, pwm[:, None, :]), oracle_bg(newick, 0)
    with np.errstate(divide="ignore"):
        return orc.estep(fg, bg, R["col_offsets"], R["codes"], np.log([1.0, 0.0]), None, None, threads=4)


def test_zoops_posterior_is_calibrated_on_the_reference_generated_positives():
    """ratio_positive_negative_testing/data_1.0.fg (first 400 alignments, ZOOPS 1.0, one site each at MOTIF_POS, star tree, where
    the mean field is exact): the E-step under the generating parameters (SingleHiddenMotifMixture.getNewWeights,
    jstacs!/.../SingleHiddenMotifMixture.java:499-566, uniform position prior) is the exact posterior of the planted position,
    so both calibration identities hold within 4 standard errors -- and fail under a slower tree or a sharpened / flattened
    PWM.  This holds the motif and background scores, their difference per offset, the position prior and the normalisation
    against positions the reference's sampler planted."""
    pwm = G["pwm_ratio"] / G["pwm_ratio"].sum(axis=1, keepdims=True)
    star = str(G["star"])
    res = _oracle_estep(star, pwm)
    z1, z2, hit, g_true = _calibration(res["gamma"], res["argmax_off"])
    assert abs(z1) < 4 and abs(z2) < 4
    assert hit > 0.8 and g_true > 0.75
    assert abs(res["gamma"].sum() - 400) < 1e-9
    sharp = pwm ** 2 / (pwm ** 2).sum(axis=1, keepdims=True)
    flat = np.sqrt(pwm) / np.sqrt(pwm).sum(axis=1, keepdims=True)
    for newick, p in ((star.replace("0.2", "0.5"), pwm), (star, sharp), (star, flat)):
        r = _oracle_estep(newick, p)
        w1, w2, _, _ = _calibration(r["gamma"], r["argmax_off"])
        assert abs(w1) > 4 and abs(w2) > 4


@pytest.mark.gpu
def test_device_zoops_posterior_on_the_reference_generated_positives():
    """The same E-step through libphyfoo.so: the oracle's posterior (1e-9), its arg-max calls bit for bit, calibrated."""
    from phyfoo_b200 import core
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    ctx = core.default_context()
    pwm = G["pwm_ratio"] / G["pwm_ratio"].sum(axis=1, keepdims=True)
    star = str(G["star"])
    smp = PhyloSample(R["codes"], R["col_offsets"])
    fg = PhyloBayesModel(15, 0, star, ctx)
    fg.setCondProbs(pwm[:, None, :])
    shm = SingleHiddenMotifMixture(fg, PhyloBackground(0, star, ctx), zoops=1.0)
    res = shm.getNewWeights(smp)
    ref = _oracle_estep(star, pwm)
    assert np.array_equal(res["argmax_off"], ref["argmax_off"])
    assert np.max(np.abs(res["gamma"] - ref["gamma"])) < 1e-9
    assert abs(res["ll"] - ref["ll"]) <= 1e-9 * abs(ref["ll"])
    z1, z2, hit, _ = _calibration(res["gamma"], res["argmax_off"])
    assert abs(z1) < 4 and abs(z2) < 4 and hit > 0.8
