"""Branch-length learning of the background model and the per-alignment learned topologies:
PhyloBackground.train with EDGE_LEARNING (models/PhyloBackground.java:160-176), TrainingUtil.learnTree
(training/TrainingUtil.java:72-80), PhyloSample.train (io/PhyloSample.java:243-270) and scoring / the E-step under an
alignment's learnedTopology (models/PhyloBackground.java:250-253,301-303; PhyloBayesModel.java:233-265).  The objective values
the optimiser reports are checked against the oracle's fixed-q objective under the same parameters; the optimiser's own path is
a stand-in."""
import numpy as np
import pytest

from util import orc, oracle_bg, oracle_fg

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    from phyfoo_b200 import core
    return core.default_context()


@pytest.mark.parametrize("one_length", [False, True])
def test_background_edge_learning_reports_the_oracles_objective(ctx, one_length):
    from phyfoo_b200 import synth
    from phyfoo_b200.models import PhyloBackground
    truth, start = synth.caterpillar_newick(5, 0.35), synth.caterpillar_newick(5, 0.1)
    smp, _ = synth.make_sample(truth, 12, 150, 23, gap_rate=0.05, length_jitter=20)
    wts = np.random.default_rng(4).random(smp.n_alignments) + 0.5
    bg = PhyloBackground(0, start, ctx)
    PhyloBackground.EDGE_LEARNING, PhyloBackground.EDGE_LEARNING_ONE_LENGTH = True, one_length
    try:
        st = bg.train(smp, np.concatenate([wts, [123.0, 456.0]]))       # extendWeights reads one weight per alignment only
    finally:
        PhyloBackground.EDGE_LEARNING, PhyloBackground.EDGE_LEARNING_ONE_LENGTH = False, False
    assert st["f_after"] >= st["edges_f_before"] > st["f_before"] and st["edges_f_after"] > st["edges_f_before"]
    # the oracle with q fixed under the STARTING model, evaluated at the learned pi and then at the learned branch lengths
    ob = oracle_bg(start, 0, [np.full(4, 0.25)])
    ofe = orc.FESet(ob, smp.codes[:, None, :], np.repeat(wts, smp.lengths()))
    assert abs(ofe.sum() - st["f_before"]) <= TOL * abs(st["f_before"])
    f_pi = ofe.eval(0, np.log(bg.getCondProb(0)))
    assert abs(f_pi - st["edges_f_before"]) <= TOL * abs(f_pi)
    for k in range(1, bg.tree.n_nodes):
        ob.set_branch(0, k, bg._branch[0, k])
    assert abs(ofe.sum() - st["edges_f_after"]) <= TOL * abs(st["edges_f_after"])
    learned = bg._branch[0, 1:]
    assert learned.mean() > 0.15                                         # moved from 0.1 towards the generating 0.35
    if one_length:
        assert np.allclose(learned, learned[0], rtol=0, atol=0)
    # the device model follows: scoring under the learned lengths equals the oracle under them
    ob2 = oracle_bg(start, 0, [bg.getCondProb(0)])
    for k in range(1, bg.tree.n_nodes):
        ob2.set_branch(0, k, bg._branch[0, k])
    cols = bg.getColumnLogProbs(smp)
    ref = ob2.bg_columns(smp.getElementAt(0))
    assert np.allclose(cols[:ref.size], ref, rtol=1e-9, atol=0)


def test_learn_tree_recovers_the_branch_scale(ctx):
    from phyfoo_b200 import synth, training
    from phyfoo_b200.newick import parse_newick
    truth, start = synth.star_newick(5, 0.3), synth.star_newick(5, 0.05)
    smp, _ = synth.make_sample(truth, 40, 200, 5)
    learned = training.learnTree(smp, start, ctx)
    tr = parse_newick(learned)
    assert tr.n_nodes == parse_newick(start).n_nodes
    b = tr.branch[1:]
    assert np.all(np.abs(b - 0.3) < 0.08), b
    assert all(len(x.split(".")[1]) <= 3 for x in __import__("re").findall(r":([0-9.]+)", learned))      # Util.round(x, 3)


def test_per_alignment_topologies_and_scoring_under_them(ctx):
    from phyfoo_b200 import synth, training
    from phyfoo_b200.data import PhyloSample, combineBranchLengths
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    from phyfoo_b200.newick import parse_newick
    base = synth.star_newick(4, 0.25)
    W = 5
    pwm = synth.random_pwm(W, 0, 12)
    slow, _ = synth.make_sample(synth.star_newick(4, 0.08), 3, 400, 1, pwm=pwm)
    fast, _ = synth.make_sample(synth.star_newick(4, 0.5), 3, 400, 2, pwm=pwm)
    smp = PhyloSample.from_alignments([slow.getElementAt(i) for i in range(3)] + [fast.getElementAt(i) for i in range(3)])
    learned = training.learnTopologies(smp, 0.7, base, ctx)
    assert smp.learnedTopology is learned and len(learned) == 6
    means = [parse_newick(t).branch[1:].mean() for t in learned]
    assert max(means[:3]) < min(means[3:])                               # the slowly evolving alignments get the shorter trees
    g = parse_newick(smp.globalTopology).branch[1:].mean()
    assert min(means) < g < max(means)

    # scoring: an alignment with a learnedTopology is scored under its own branch lengths (one batch call per run)
    bg = PhyloBackground(0, base, ctx)
    fg = PhyloBayesModel(W, 0, base, ctx); fg.setCondProbs(pwm)
    got_bg, got_fg = bg.getLogProbFor(smp), fg.getLogProbFor(smp)
    plain = PhyloSample(smp.codes, smp.col_offsets)
    nwin = np.concatenate([[0], np.cumsum(smp.lengths() - W + 1)])
    for i in (0, 4):
        ob = oracle_bg(learned[i], 0, [np.full(4, 0.25)])
        ref = ob.bg_columns(smp.getElementAt(i)).sum()
        assert abs(got_bg[i] - ref) <= TOL * abs(ref)
        of = oracle_fg(combineBranchLengths(learned[i], base, 1.0), pwm)
        ref_w = of.score_windows(smp.getElementAt(i))[0]
        assert np.allclose(got_fg[nwin[i]:nwin[i + 1]], ref_w, rtol=TOL, atol=0)
    assert np.abs(got_bg - bg.getLogProbFor(plain)).max() > 1e-3          # the topologies do change the scores
    # the reference resets the model from Newick strings: branch lengths come back rounded to three decimals
    assert np.allclose(bg._branch[0, 1:], 0.25) and np.allclose(fg._branch[:, 1:], 0.25)

    # E-step: ll and w add up over the runs; against the oracle's E-step of each alignment under its own trees
    shm = SingleHiddenMotifMixture(fg, bg, zoops=0.8)
    res = shm.getNewWeights(smp)
    ll, w = 0.0, np.zeros(2)
    for i in range(6):
        ob = oracle_bg(learned[i], 0, [np.full(4, 0.25)])
        of = oracle_fg(learned[i], pwm)
        a = smp.getElementAt(i)
        r = orc.estep(of, ob, np.array([0, a.shape[0]], np.int64), a, np.log([0.8, 0.2]))
        ll += r["ll"]; w += r["w"]
        assert res["argmax_off"][i] == r["argmax_off"][0]
        assert np.abs(res["gamma"][nwin[i]:nwin[i + 1]] - r["gamma"]).max() < TOL
    assert abs(res["ll"] - ll) <= TOL * abs(ll) and np.allclose(res["w"], w, rtol=TOL)
    # two neighbours with the same topology form one run
    smp.learnedTopology = [learned[0], learned[0], None, learned[3], learned[3], learned[3]]
    res2 = shm.getNewWeights(smp)
    assert res2["gamma"].size == res["gamma"].size and np.isfinite(res2["ll"])
    assert np.allclose(res2["gamma"][nwin[0]:nwin[1]], res["gamma"][nwin[0]:nwin[1]], rtol=0, atol=1e-12)


def test_tree_learning_on_the_references_heterogeneity_data(ctx):
    """The reference's bundled data/synthetic_data/heterogeneity (tests/golden/heterogeneity_bg_first60.npz,
    scripts/make_heterogeneity_fixture.py): data_1.0.bg was simulated on a star tree with every branch 0.2, data_rBG_1.0.bg with
    every branch of every alignment drawn from beta(3, 7) (mean 0.3).  TrainingUtil.learnTree recovers the two scales from a
    common start, and PhyloSample.train gives the divergent alignments of the heterogeneous set the longer trees."""
    import os
    from util import ROOT
    from phyfoo_b200 import training
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.newick import parse_newick
    z = np.load(os.path.join(ROOT, "tests", "golden", "heterogeneity_bg_first60.npz"))
    start = str(z["newick_star"]).replace("0.2", "0.1")
    homo = PhyloSample(z["codes_homogeneous"], z["col_offsets_homogeneous"])
    hetero = PhyloSample(z["codes_heterogeneous"], z["col_offsets_heterogeneous"])
    b_homo = parse_newick(training.learnTree(homo, start, ctx)).branch[1:]
    b_hetero = parse_newick(training.learnTree(hetero, start, ctx)).branch[1:]
    assert np.all(np.abs(b_homo - 0.2) < 0.06), b_homo
    assert b_hetero.mean() > b_homo.mean() + 0.04 and np.all(np.abs(b_hetero - 0.3) < 0.1), (b_homo, b_hetero)
    # per-alignment trees: the mean branch length of an alignment follows how often its species disagree
    sub = hetero.slice(0, 16)
    learned = training.learnTopologies(sub, 0.9, start, ctx)
    means = np.array([parse_newick(t).branch[1:].mean() for t in learned])
    disagree = []
    for i in range(sub.n_alignments):
        a = sub.getElementAt(i)
        disagree.append(np.mean([(a[:, s] != a[:, r]).mean() for s in range(5) for r in range(s)]))
    r = np.corrcoef(means, np.array(disagree))[0, 1]
    assert r > 0.6, (r, means, disagree)
    assert means.std() > 0.01
