"""Host logic of phyfoo_b200/training.py and the learned-topology bookkeeping, without a device: combineBranchLengths
(io/PhyloSample.java:272-310), the runs of alignments sharing a learnedTopology, the trainModel loop
(training/TrainingUtil.java:21-43) on a stand-in model."""
import numpy as np
import pytest

from phyfoo_b200.data import PhyloSample, combineBranchLengths
from phyfoo_b200.models import topology_runs
from phyfoo_b200 import training


def test_combine_branch_lengths_is_the_weighted_mean_in_java_notation():
    a = "((A:0.1,B:0.2):0.05,C:0.3);"
    b = "((A:0.3,B:0.4):0.15,C:0.1);"
    assert combineBranchLengths(a, b, 1.0) == "((A:0.1,B:0.2):0.05,C:0.3);"
    assert combineBranchLengths(a, b, 0.0) == "((A:0.3,B:0.4):0.15,C:0.1);"
    got = combineBranchLengths(a, b, 0.25)
    vals = [float(x) for x in __import__("re").findall(r":([0-9.E-]+)", got)]
    assert np.allclose(vals, [0.25, 0.35, 0.125, 0.15], rtol=0, atol=1e-15)
    assert got.startswith("((A:0.24999999999999997,")                    # String.valueOf(double): shortest round-trip digits
    # lengths are matched by their order of appearance; the names and the structure are newick1's
    assert combineBranchLengths("(X:0.5,Y:0.25)", "(P:0.1,Q:0.05,R:0.3)", 0.5) == "(X:0.3,Y:0.15)"
    with pytest.raises(IndexError):
        combineBranchLengths(a, "(A:0.1,B:0.2)", 0.5)
    # an integer length has no '.', so the reference's pattern does not see it (kept)
    assert combineBranchLengths("(A:1,B:0.5)", "(A:0.25,B:0.125)", 0.5) == "(A:1,B:0.375)"
    assert PhyloSample.combineBranchLengths(a, b, 1.0) == a


def test_topology_runs_and_slices():
    codes = np.zeros((12, 3), np.uint8)
    smp = PhyloSample(codes, [0, 2, 5, 7, 10, 12])
    assert topology_runs(smp) is None
    smp.learnedTopology = [None] * 5
    assert topology_runs(smp) is None
    smp.learnedTopology = ["t1", "t1", None, "t2", "t1"]
    assert topology_runs(smp) == [(0, 2, "t1"), (2, 3, None), (3, 4, "t2"), (4, 5, "t1")]
    sub = smp.slice(1, 4)
    assert sub.learnedTopology == ["t1", None, "t2"] and sub.n_alignments == 3


class _FakeModel:
    """log-probability approaches 0 geometrically with every training round."""

    def __init__(self, ratio):
        self.rounds, self.ratio, self.seen_weights = 0, ratio, []

    def train(self, data, weights=None):
        self.rounds += 1
        self.seen_weights.append(weights)

    def getLogProbFor(self, data):
        return np.array([-100.0 * self.ratio ** self.rounds, 0.0])


def test_train_model_loop_stops_like_the_reference():
    m = _FakeModel(0.5)                       # differences 50, 25, 12.5, ... : the difference after round k is 100 / 2^k, first <= 0.1 at k = 10
    rounds, ll = training.trainModel(m, None, 1e-1)
    assert rounds == 10 and abs(ll + 100.0 * 0.5 ** 10) < 1e-12
    m = _FakeModel(0.99)                      # never converges: 20 rounds at most
    assert training.trainModel(m, None, 1e-9)[0] == 20
    m = _FakeModel(0.0)                       # converged at once: still two rounds (the first difference is against -2^31)
    assert training.trainModel(m, None, 1e-1)[0] == 2
    m = _FakeModel(0.0)
    assert training.trainModel(m, None, 1e-1, minSteps=7, weights=[1.0])[0] == 7 and m.seen_weights[0] == [1.0]


def test_scoring_under_a_learned_topology_swaps_and_restores_branch_lengths():
    """models/PhyloBackground.java:250-253,301-303 and PhyloBayesModel.java:233-265 without a device: inside the call the model
    carries the alignment's branch lengths, afterwards those of the Newick strings taken before -- i.e. rounded to three
    decimals (bayesNet/VirtualTree.java:208-234), the reference's side effect."""
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel
    base = "((A:0.12345,B:0.2):0.30049,C:0.4)"
    topo = "((A:0.5,B:0.25):0.125,C:0.0625);"
    bg = PhyloBackground(0, base)
    with bg._under_topology(None):
        assert np.allclose(bg._branch[0, 1:], [0.30049, 0.12345, 0.2, 0.4], rtol=0, atol=0)
    with bg._under_topology(topo):
        assert np.allclose(bg._branch[0, 1:], [0.125, 0.5, 0.25, 0.0625], rtol=0, atol=0)
    assert np.allclose(bg._branch[0, 1:], [0.3, 0.123, 0.2, 0.4], rtol=0, atol=1e-15)
    fg = PhyloBayesModel(3, 0, base)
    fg.setBranchLength(1, 2, 0.77777)
    with fg._under_topology(topo):
        assert np.allclose(fg._branch[:, 1:], np.tile([0.125, 0.5, 0.25, 0.0625], (3, 1)), rtol=0, atol=0)
    assert np.allclose(fg._branch[0, 1:], [0.3, 0.123, 0.2, 0.4], rtol=0, atol=1e-15)
    assert abs(fg._branch[1, 2] - 0.778) < 1e-15                  # per-position lengths come back too, rounded
    with pytest.raises(ValueError):
        with bg._under_topology("(A:0.1,B:0.2,C:0.3)"):            # another topology
            pass
