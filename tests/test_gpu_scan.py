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
