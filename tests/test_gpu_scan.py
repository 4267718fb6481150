"""Genome-scan / classification front-end (classification/ClassificationUtil.java semantics): the device-side consumers of
the per-offset scores against the same computations in numpy on the host copy of the scores (bit-exact: thresholds are order
statistics, hits are comparisons), and the host mirrors against hand-checked cases.  The first test is about the scan LOGIC: its
scores are the GPU's own on both sides (self-referential for the scores by design); test_scan_on_oracle_scores closes that loop
with thresholds, hit counts and arg-max offsets computed from the oracle's window scores."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def setup():
    from phyfoo_b200 import core, synth
    from phyfoo_b200.models import PhyloBayesModel
    ctx = core.default_context()
    newick = synth.random_binary_newick(12, 3)
    W = 7
    pwm = synth.random_pwm(W, 0, 21)
    pos, planted = synth.make_sample(newick, 40, 90, 3, pwm=pwm, gap_rate=0.02, length_jitter=15)
    neg, _ = synth.make_sample(newick, 50, 80, 4, gap_rate=0.02, length_jitter=20)
    fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(pwm)
    return ctx, fg, pos, neg, planted


@pytest.mark.gpu
def test_device_scan_matches_host_semantics(setup):
    from phyfoo_b200 import classification as cl
    ctx, fg, pos, neg, planted = setup
    ps, ns = cl.getScores(fg, pos), cl.getScores(fg, neg)
    spec = [0.5, 0.9, 0.99, 0.999, 1.0, 0.0]
    thr_host = cl.determineThresholdsForSpecifitiesPerPos(ns, spec)
    thr_dev = cl.determineThresholdsForSpecifitiesPerPos_device(fg, neg, spec)
    assert np.array_equal(thr_host, thr_dev)
    one = cl.determineThresholdsForSpecifitiesPerPos_device(fg, neg, [0.95], rule="single")[0]
    assert one == cl.determineThresholdForSpecifity(0.95, ns)
    for t in thr_host[:4]:
        h = cl.scan_hits_device(fg, pos, t)
        assert np.array_equal(h["n_hits"], [int((x > t).sum()) for x in ps])
        assert np.array_equal(h["first_hit"], [int(np.argmax(x > t)) if (x > t).any() else -1 for x in ps])
        assert np.array_equal(h["arg_max"], [int(np.argmax(x)) for x in ps])
        assert np.array_equal(h["max_score"], [x.max() for x in ps])
        assert cl.determineSensitivityPerSeq_device(t, fg, pos) == cl.determineSensitivityPerSeq(t, ps)
        assert cl.determinePositionSet(t, ps) == {(i, u) for i, x in enumerate(ps) for u in np.nonzero(x > t)[0]}
    sens = cl.getSensitivitiesPerSeqForSpecifitiesPerPos(ps, ns, [0.99, 0.999])
    assert sens[0] >= sens[1] and sens[0] > 0.5          # planted sites stand out of the background
    assert (cl.scan_hits_device(fg, pos, -np.inf)["arg_max"] == planted).mean() >= 0.6
    # reverse strand scan runs too and differs
    assert not np.array_equal(cl.scan_hits_device(fg, pos, thr_host[1], strand=1)["max_score"],
                              cl.scan_hits_device(fg, pos, thr_host[1], strand=0)["max_score"])


def test_host_mirrors_hand_checked():
    from phyfoo_b200 import classification as cl
    neg = [np.array([0.0, 1.0, 2.0]), np.array([]), np.array([3.0, 4.0])]
    assert list(cl.determineThresholdsForSpecifitiesPerPos(neg, [0.0, 0.5, 0.99, 1.0])) == [0.0, 2.0, 3.0, 4.0]
    assert cl.determineThresholdForSpecifity(0.5, neg) == 2.0
    posm = [np.array([0.5, 3.5]), np.array([1.0]), np.array([5.0, 5.0])]
    assert cl.determineSensitivityPerSeq(3.0, posm) == 2 / 3
    assert cl.determinePositionSet(3.0, posm) == {(0, 1), (2, 0), (2, 1)}
    assert cl.calculateAUC([0.5, 1.0], [0.5, 1.0]) == 0.5 * 0.25 + 0.5 * 0.75
    assert cl.getDynamicClassificationRate([1.0, 2.0, 3.0], [0.0, 0.5, 2.5], thresh=0.75) == 5 / 6
    assert cl.getDynamicClassificationRate([1.0, 2.0, 3.0], [0.0, 0.5, 2.5]) == 5 / 6
    with pytest.raises(ValueError):
        cl.calculateAUC([1.0], [1.0, 2.0])


@pytest.mark.gpu
def test_scan_on_oracle_scores(setup):
    """The same consumers with the ORACLE's window scores as the reference: hit counts / first hits / arg-max offsets of the
    device scan equal those of numpy on the oracle's scores wherever no oracle score lies within 1e-9 relative of the
    threshold, and the device threshold for a specificity equals the oracle's order statistic to 1e-9."""
    from phyfoo_b200 import classification as cl
    from util import oracle_fg
    ctx, fg, pos, neg, planted = setup
    om = oracle_fg(fg.newick, fg.getCondProbs())
    ops = [om.score_windows(pos.getElementAt(i), False)[0] for i in range(pos.n_alignments)]
    ons = [om.score_windows(neg.getElementAt(i), False)[0] for i in range(neg.n_alignments)]
    spec = [0.5, 0.9, 0.99]
    thr_ref = cl.determineThresholdsForSpecifitiesPerPos(ons, spec)
    thr_dev = cl.determineThresholdsForSpecifitiesPerPos_device(fg, neg, spec)
    assert np.max(np.abs(thr_dev - thr_ref) / np.abs(thr_ref)) < 1e-9
    for t in thr_ref:
        h = cl.scan_hits_device(fg, pos, t)
        for i, x in enumerate(ops):
            assert h["arg_max"][i] == int(np.argmax(x)) and abs(h["max_score"][i] - x.max()) <= 1e-9 * abs(x.max())
            if np.min(np.abs(x - t)) > 1e-9 * abs(t):              # no score on the knife's edge of the threshold
                assert h["n_hits"][i] == int((x > t).sum())
                assert h["first_hit"][i] == (int(np.argmax(x > t)) if (x > t).any() else -1)


@pytest.mark.gpu
def test_classification_on_the_references_bundled_positives_and_negatives():
    """BASELINE configs[0] end to end through the classification front-end: the first 200 alignments of the reference's
    data/synthetic_data/trees/data_tree1_1.0.fg (one planted site each) against the first 200 of data_tree1_1.0.bg (none), scored
    under the generating PWM and tree (tests/golden/c1_tree1_first200.npz, c1_tree1_bg_first200.npz; scripts/make_c1_fixture.py).
    Thresholds from the negatives, sensitivity on the positives (classification/ClassificationUtil.java:231-278), the per-alignment
    classification by log-likelihood differences (:82-159); device flavours equal the host flavours on the same scores, and the
    scores of a few alignments are the oracle's."""
    import os
    from util import ROOT, oracle_fg, rel_err
    from phyfoo_b200 import classification as cl, core, synth
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    ctx = core.default_context()
    cfg, newick, pwm, pos, planted = synth.make_config("c1")
    z = np.load(os.path.join(ROOT, "tests", "golden", "c1_tree1_bg_first200.npz"))
    neg = PhyloSample(z["codes"], z["col_offsets"])
    W = cfg["W"]
    fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(pwm)
    bg = PhyloBackground(0, newick, ctx)
    ps, ns = cl.getScores(fg, pos), cl.getScores(fg, neg)
    om = oracle_fg(newick, pwm)
    for i in (0, 117):
        assert rel_err(ps[i], om.score_windows(pos.getElementAt(i))[0]) < 1e-9
        assert rel_err(ns[i], om.score_windows(neg.getElementAt(i))[0]) < 1e-9
    spec = [0.9, 0.99, 0.999, 0.9999]
    thr = cl.determineThresholdsForSpecifitiesPerPos(ns, spec)
    assert np.array_equal(thr, cl.determineThresholdsForSpecifitiesPerPos_device(fg, neg, spec))
    sens = cl.getSensitivitiesPerSeqForSpecifitiesPerPos(ps, ns, spec)
    assert np.all(np.diff(sens) <= 0) and sens[1] > 0.8             # at 99 % per-position specificity most planted sites are found
    for t, s in zip(thr, sens):
        assert cl.determineSensitivityPerSeq_device(t, fg, pos) == s
    hits = cl.scan_hits_device(fg, pos, thr[2])
    assert np.array_equal(hits["n_hits"], [int((x > thr[2]).sum()) for x in ps])
    assert np.array_equal(hits["arg_max"], [int(np.argmax(x)) for x in ps])
    # (the raw motif score -- not the ratio to the background the E-step uses -- peaks at the planted site in a minority of the
    # alignments only: conserved columns score high under any phylogenetic model)
    # false-positive rate per negative alignment at the same threshold stays near (1 - specificity) x windows per alignment
    fp = cl.determineSensitivityPerSeq_device(thr[2], fg, neg)
    assert fp < 0.5
    # ROC over the per-position thresholds and the per-alignment classification under the ZOOPS mixture
    grid = np.linspace(0.0, 1.0, 101)
    tg = cl.determineThresholdsForSpecifitiesPerPos(ns, grid)
    tpr = np.array([cl.determineSensitivityPerSeq(t, ps) for t in tg])
    fpr = np.array([cl.determineSensitivityPerSeq(t, ns) for t in tg])
    o = np.argsort(fpr, kind="stable")
    auc = cl.calculateAUC(np.append(fpr[o], 1.0), np.append(tpr[o], 1.0))
    assert 0.5 < auc <= 1.0                                       # (0.535: per-alignment ROC of the RAW motif score, see above)
    shm = SingleHiddenMotifMixture(fg, bg, zoops=0.9)

    def per_alignment_loglik(smp):
        # getLogProbFor(seq) of the mixture (jstacs SingleHiddenMotifMixture.java:520-550) from the E-step's profile a_j
        res = shm.getNewWeights(smp, want_gamma=False, want_profile=True)
        off = smp.device(ctx).window_offsets(W)
        z = np.array([np.logaddexp.reduce(res["profile"][off[i]:off[i + 1]]) for i in range(smp.n_alignments)])
        ll = bg.getLogProbFor(smp) + np.logaddexp(np.log(0.9) + z, np.log(0.1))
        assert abs(ll.sum() - res["ll"]) <= 1e-9 * abs(res["ll"])
        return ll
    d_pos = cl.getDifferences(per_alignment_loglik(pos), bg.getLogProbFor(pos))
    d_neg = cl.getDifferences(per_alignment_loglik(neg), bg.getLogProbFor(neg))
    rate = cl.getDynamicClassificationRate(d_pos, d_neg)
    assert rate > 0.75                                            # 0.8325: the mixture's likelihood ratio separates the two sets
