#!/usr/bin/env python
"""Prints what tests/refexec.py executes: for every reference method it translates, the Python source it generates from the
reference's file at run time (nothing of this is stored in the repository).  Put next to the Java (file:line in the headings)
to see that the translation is statement by statement: same statements, same order, same operands.

    python scripts/show_reference_translation.py [method-name-substring]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np                                    # noqa: E402
import refexec                                        # noqa: E402
from oracle import np_restatement as npr              # noqa: E402


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else ""
    parts = []
    src = refexec.translate()
    parts += [("algorithm/MeanFieldForBayesNet.java:52-69  init()", src["init"]),
              ("algorithm/MeanFieldForBayesNet.java:92-211  optimizeByNormalisation()", src["optimizeByNormalisation"]),
              ("algorithm/MeanFieldForBayesNet.java:232-365  calcFreeEnergy(int, int)", src["calcFreeEnergy2"])]
    es = refexec.translate_estep()
    parts += [("jstacs.jar SingleHiddenMotifMixture.java:499-566  getNewWeights", es["getNewWeights"]),
              ("jstacs.jar SingleHiddenMotifMixture.java:585-596  modify (case EM)", es["modify"]),
              ("jstacs.jar UniformPositionPrior.java:55-61  getLogPriorForPositions", es["getLogPriorForPositions"]),
              ("jstacs.jar Normalisation.java:241-253  logSumNormalisation (6 arguments)", es["lsn6"]),
              ("jstacs.jar Normalisation.java:352-376  logSumNormalisation (7 arguments)", es["lsn7"])]
    for name in ("FS81alpha", "FS81beta", "HKY"):
        parts.append((f"evolution/{name}.java  reinit(double...)", refexec.ReferenceEvolModel(name, np.full((1, 4), 0.25)).source))
    net = npr.Net.motif("(A:0.1,B:0.2)", 1, 0)
    net.build_tables()
    refexec.ReferenceObjective(net, [np.zeros((1, 2), np.uint8)], [1.0], 0)
    parts.append(("jstacs.jar NumericalDifferentiableFunction.java:64-84  evaluateGradientOfFunction", refexec.ReferenceObjective.gradient_source))
    parts.append(("optimizing/AbstractFreeEnergyOptimizer.java:95-114  getWeightedLogLikelihoodSum()", refexec.ReferenceObjective.sum_source))
    refexec.ReferenceSelector([0.2, 0.5, 0.3])
    parts.append(("io/SequenceSpecificDataSelector.java:82-96  prepareForTraining (its for statement)", refexec.ReferenceSelector.source))
    parts.append(("bayesNet/CPF.java:238-256  getQueryByIndex(int, boolean)", refexec.reference_lookup_table(2)[1]))
    bnet = npr.Net.background("(A:0.1,B:0.2)", 1)
    bnet.build_tables()
    refexec.ReferenceBackground(bnet).getMFfromCache(refexec._MDSeq(np.zeros((2, 2), np.uint8)))
    refexec.ReferenceMotif(net)
    parts.append(("algorithm/MeanFieldForBayesNet.java:78-89  initObservation()", refexec._SharedMeanField.init_observation_source))
    parts.append(("models/PhyloBackground.java:241-305  getLogProbFor(MultiDimensionalDiscreteSequence, int, int)", refexec.ReferenceBackground.source))
    parts.append(("models/PhyloBayesModel.java:222-267  getLogProbFor(Sequence, int, int)", refexec.ReferenceMotif.motif_source))
    refexec.ReferenceAlignmentBasedBackground([[0.25] * 4], 0)
    parts.append(("models/AlignmentBasedModel.java:187-228  getLnCondProb(sequence, posSeq, posMot)", refexec.ReferenceAlignmentBasedModel.source))
    parts.append(("models/AlignmentBasedBGModel.java:173-222  getLnCondProb(sequence, posSeq)", refexec.ReferenceAlignmentBasedBackground.source))
    refexec.ReferenceEdgeObjective(net, [np.zeros((1, 2), np.uint8)], [1.0], 0)
    parts.append(("optimizing/branchlengths/EdgeOptimizer.java:44-74  evaluateFunction", refexec.ReferenceEdgeObjective.edge_source))
    opt = refexec.ReferenceOptimizer(lambda x: 0.0, lambda x: [0.0]).source
    parts += [("jstacs.jar Optimizer.java:122-146  findBracket", opt["findBracket"]), ("jstacs.jar Optimizer.java:202-303  brentsMethod", opt["brentsMethod"]),
              ("jstacs.jar Optimizer.java:1000-1080  quasiNewtonBFGS", opt["quasiNewtonBFGS"]),
              ("jstacs.jar OneDimensionalSubFunction.java:62-74  evaluateFunction", opt["along"])]
    parts.append(("jstacs.jar StrandModel.java:561-583  getNewWeights + AbstractMixtureModel.java:1074-1078  modifyWeights (case EM)",
                  (refexec.ReferenceStrandEStep([0.0], [0.0], 0.5), refexec.ReferenceStrandEStep.source)[1]))
    refexec.ReferenceGlobalEdgeObjective(net, [np.zeros((1, 2), np.uint8)], [1.0])
    parts.append(("optimizing/branchlengths/GlobalEdgeOptimizer.java:33-58  evaluateFunction", refexec.ReferenceGlobalEdgeObjective.global_source))
    parts.append(("jstacs.jar termination/CombinedCondition.java:108-117  doNextIteration", refexec.reference_condition().source))
    parts.append(("jstacs.jar LimitedMedianStartDistance.java:64-84  getNewStartDistance / setLastDistance", refexec.reference_forecaster().source))
    for title, text in parts:
        if want.lower() in title.lower():
            print("=" * 100 + "\n" + title + "\n" + "=" * 100 + "\n" + text)


if __name__ == "__main__":
    main()
