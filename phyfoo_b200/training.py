"""The reference's small training drivers around the background model's branch-length learning:
`TrainingUtil.trainModel` / `learnTree` (training/TrainingUtil.java:21-43,72-80, what tools/LearnTree.java runs) and the
per-alignment tree learning of `PhyloSample.train` (io/PhyloSample.java:243-270) with `combineBranchLengths` (:272-310).

Everything numerical happens behind the C ABI: one device pass per `PhyloBackground.train` call fixes the mean fields of every
column and collects the expected counts (phyfoo_fe_prepare_all, phyfoo_ss_create); the optimisers then evaluate host dot products
inside the library (phyfoo_ss_eval / phyfoo_ss_eval_branches).
"""
import numpy as np

from .data import PhyloSample, combineBranchLengths
from .models import PhyloBackground


def trainModel(model, data, eps, minSteps=0, weights=None):
    """training/TrainingUtil.java:21-43: train until the summed log-probability of the data changes by at most `eps` between
    two rounds (at least one round, at least `minSteps`, at most 20).  Returns (rounds, last summed log-probability)."""
    lastLog = curLog = float(-2 ** 31)               # Integer.MIN_VALUE
    i = 0
    while i == 0 or i < minSteps or (abs(lastLog - curLog) > eps and i < 20):
        model.train(data, weights)                   # train(data) = weights of 1 (models/AbstractPhyloModel.java:60-64)
        lastLog = curLog
        curLog = float(np.sum(model.getLogProbFor(data)))
        i += 1
    return i, curLog


def learnTree(data: PhyloSample, initialTree: str, ctx=None):
    """training/TrainingUtil.java:72-80: an order-0 background with EDGE_LEARNING on, trained by trainModel(bg, data, 1e-1, 0,
    null); the learned tree is the Newick string of its (only) position, branch lengths rounded to three decimals."""
    bg = PhyloBackground(0, initialTree, ctx)
    PhyloBackground.EDGE_LEARNING = True
    try:
        trainModel(bg, data, 1e-1, 0, None)
    finally:
        PhyloBackground.EDGE_LEARNING = False
        PhyloBackground.EDGE_LEARNING_ONE_LENGTH = False
    return bg.getNewickString(0)


def learnTopologies(sample: PhyloSample, weight: float, baseTopology: str, ctx=None, global_rounds=5):
    """`PhyloSample.train(sample, weight, baseTopology)` (io/PhyloSample.java:243-270): five rounds of background training with
    branch-length learning on the whole sample give the global tree; then for every alignment the branch lengths are
    re-initialised from `baseTopology`, the background is trained on that alignment alone, and the alignment's
    `learnedTopology` is the weighted mean of its own tree and the global one.  As in the reference the stationary
    distribution is NOT re-initialised between alignments, and the weight an alignment's columns carry is `weights[0]` of the
    vector built for it -- `weight` for alignment 0, (1 - weight) / n for every other one -- because extendWeights indexes the
    one-element sub-sample (models/PhyloBackground.java:178-186).  Sets and returns `sample.learnedTopology` (list of Newick
    strings)."""
    n = sample.n_alignments
    bg = PhyloBackground(0, baseTopology, ctx)
    PhyloBackground.EDGE_LEARNING = True
    try:
        for _ in range(global_rounds):
            bg.train(sample, None)
        globalTopology = bg.getNewickString(0)
        learned = []
        for i in range(n):
            bg.reinitEdgelengths(baseTopology)
            weights = np.full(n, (1.0 - weight) / n)
            weights[i] = weight
            bg.train(sample.slice(i, i + 1), weights)
            learned.append(combineBranchLengths(bg.getNewickString(0), globalTopology, weight))
    finally:
        PhyloBackground.EDGE_LEARNING = False
    sample.learnedTopology = learned
    sample.globalTopology = globalTopology
    return learned
