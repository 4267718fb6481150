"""A model written to and read back from the reference's XML format scores like a model with the same (three-decimal)
branch lengths: the file carries everything the scoring path needs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("order", [0, 1])
def test_reloaded_model_scores_identically(order):
    from phyfoo_b200 import core, synth, xmlio
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    ctx = core.default_context()
    newick = "((A:0.1234,B:0.3):0.15,(C:0.2,(D:0.05049,E:0.4):0.25):0.1,F:0.3)"
    rounded = "((A:0.123,B:0.3):0.15,(C:0.2,(D:0.05,E:0.4):0.25):0.1,F:0.3)"
    W = 6
    pwm = synth.random_pwm(W, order, 8)
    smp, _ = synth.make_sample(rounded, 12, 70, 3, pwm=pwm, gap_rate=0.05)
    a = PhyloBayesModel(W, order, newick, ctx); a.setCondProbs(pwm)
    b = PhyloBayesModel.fromXML(a.toXML(), ctx)
    c = PhyloBayesModel(W, order, rounded, ctx); c.setCondProbs(pwm)
    assert np.array_equal(b.getLogProbFor(smp), c.getLogProbFor(smp))
    assert not np.array_equal(a.getLogProbFor(smp), c.getLogProbFor(smp))       # the rounding on save is real
    bg = PhyloBackground(0, newick, ctx); bg.setCondProb(np.array([[0.3, 0.2, 0.1, 0.4]]), 0)
    bg2 = PhyloBackground.fromXML(bg.toXML(), ctx)
    bgc = PhyloBackground(0, rounded, ctx); bgc.setCondProb(np.array([[0.3, 0.2, 0.1, 0.4]]), 0)
    assert np.array_equal(bg2.getColumnLogProbs(smp), bgc.getColumnLogProbs(smp))
    if order == 0:
        r1 = SingleHiddenMotifMixture(b, bg2, zoops=0.8).getNewWeights(smp)
        r2 = SingleHiddenMotifMixture(c, bgc, zoops=0.8).getNewWeights(smp)
        assert r1["ll"] == r2["ll"] and np.array_equal(r1["gamma"], r2["gamma"])
