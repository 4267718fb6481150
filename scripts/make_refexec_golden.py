#!/usr/bin/env python
"""Generates tests/golden/refexec_meanfield.npz: window scores and sweep counts produced by the REFERENCE'S OWN STATEMENTS --
MeanFieldForBayesNet.init / optimizeByNormalisation / calcFreeEnergy of /root/reference/main/java/src/algorithm/
MeanFieldForBayesNet.java, translated token by token and executed by tests/refexec.py (no JVM exists in this image) -- on
seeded inputs: motif windows of orders 0 and 1 on four trees (both strands, with and without gaps) and background ranges of
orders 0..2, and one ZOOPS E-step (SingleHiddenMotifMixture.getNewWeights / modify and Normalisation.logSumNormalisation from
libs/jstacs.jar, fed with those scores).  The .npz holds the inputs next to the outputs, so it can be checked anywhere: tests/test_reference_statements.py
holds the oracle against it (bit for bit) and tests/test_zz_gpu_reference_statements.py the CUDA path.

Runs only where /root/reference exists (the build container):    python scripts/make_refexec_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refexec                                        # noqa: E402
from oracle import np_restatement as npr              # noqa: E402
from phyfoo_b200 import synth                         # noqa: E402

TREES = {"star5": synth.star_newick(5, 0.2), "cat5": synth.caterpillar_newick(5, 0.2),
         "mixed6": "((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.4):0.25):0.1,F:0.3)", "rand12": synth.random_binary_newick(12, 3)}


def codes(rng, L, S, gap_rate):
    c = rng.integers(0, 4, size=(L, S)).astype(np.uint8)
    if gap_rate > 0:
        c[rng.random((L, S)) < gap_rate] = 4
        c[(c == 4).all(axis=1), 0] = 0                 # the reference drops all-gap columns when it reads a file
    return c


def motif_cases(rng):
    for name, newick in TREES.items():
        S = len(npr.parse_newick(newick)[3])
        for order, W in ((0, 5), (1, 4)):         # widths the GPU tests of both orders also use
            pwm = synth.random_pwm(W, order, 100 + 7 * order + S)
            net = npr.Net.motif(newick, W, order)
            for t, p in enumerate(pwm):
                net.set_pi(t, p)
            net.build_tables()
            ref = refexec.ReferenceMeanField(net)
            for gap_rate in (0.0, 0.15):
                aln = codes(rng, W + (3 if name == "rand12" else 5), S, gap_rate)
                n = aln.shape[0] - W + 1
                score, sweeps = np.zeros((2, n)), np.zeros((2, n), np.int32)
                for strand in (0, 1):
                    for j in range(n):
                        score[strand, j], sweeps[strand, j] = ref.score(aln[j:j + W], bool(strand))
                yield dict(kind="motif", tree=name, newick=newick, order=order, W=W, gap_rate=gap_rate,
                           pwm=np.concatenate([np.ravel(p) for p in pwm]), codes=aln, score=score, sweeps=sweeps)


def background_cases(rng):
    """PhyloBackground.getLogProbFor on ranges of exactly order + 1 columns (models/PhyloBackground.java:241-305): the net
    of order + 1 trees is observed, optimised, the first `order` terms calcFreeEnergy(i, i) are taken, then it is re-initialised,
    optimised again and calcFreeEnergy(order, order) is added; the value is minus the sum.  (The driver loop is restated in
    these few lines; every free-energy number comes from the reference's statements.)"""
    for name in ("cat5", "mixed6"):
        newick = TREES[name]
        S = len(npr.parse_newick(newick)[3])
        for order in (0, 1, 2):
            net = npr.Net.background(newick, order)
            r = np.random.default_rng(200 + order)
            for t in range(order + 1):
                p = r.dirichlet(np.ones(4), size=4 ** t)
                net.set_pi(t, p)
            net.build_tables()
            ref = refexec.ReferenceMeanField(net)
            for gap_rate in (0.0, 0.15):
                aln = codes(rng, order + 5, S, gap_rate)
                n = aln.shape[0] - order
                value, sweeps = np.zeros(n), np.zeros(n, np.int32)
                for j in range(n):
                    cols = aln[j:j + order + 1]
                    ref.initObservation(cols)
                    ref.init()
                    ref.free_energy_calls = 0
                    ref.optimizeByNormalisation()
                    s = sum(ref.calcFreeEnergy(i, i) for i in range(order))
                    ref.initObservation(cols)
                    ref.init()
                    ref.optimizeByNormalisation()
                    s += ref.calcFreeEnergy(order, order)
                    value[j], sweeps[j] = -s, ref.free_energy_calls - 2
                yield dict(kind="background", tree=name, newick=newick, order=order, W=order + 1, gap_rate=gap_rate,
                           pwm=np.concatenate([np.ravel(p) for p in net.pi]), codes=aln, score=value[None, :], sweeps=sweeps[None, :])


def estep_case(rng):
    """One ZOOPS E-step in which EVERY number comes from reference statements: the motif window scores and the background's
    per-column terms from MeanFieldForBayesNet (above), the posterior from SingleHiddenMotifMixture.getNewWeights / modify and
    Normalisation.logSumNormalisation of libs/jstacs.jar (refexec.ReferenceEStep).  Background of order 0: a range is the sum of
    its columns' terms in column order (models/PhyloBackground.java:262-294 with order = 0: one mean field per column)."""
    newick, W, zoops = TREES["cat5"], 5, 0.8
    lengths = [W, 14, 9]
    aln = codes(rng, sum(lengths), 5, 0.05)
    off = np.concatenate([[0], np.cumsum(lengths)])
    pwm = synth.random_pwm(W, 0, 311)
    net = npr.Net.motif(newick, W, 0)
    for t, p in enumerate(pwm):
        net.set_pi(t, p)
    net.build_tables()
    fg = refexec.ReferenceMeanField(net)
    bnet = npr.Net.background(newick, 0)
    bnet.build_tables()
    bg = refexec.ReferenceMeanField(bnet)
    scores = [fg.score(aln[off[i] + j:off[i] + j + W])[0] for i in range(len(lengths)) for j in range(lengths[i] - W + 1)]
    minus_fe = [bg.score(aln[c:c + 1])[0] for c in range(aln.shape[0])]

    def background(i, s, e):
        log_sum = 0.0
        for c in range(s, e + 1):                       # logSum += calcFreeEnergy(0, 0) column by column; the value is -logSum
            log_sum += -minus_fe[off[i] + c]
        return -log_sum if e >= s else 0.0

    weights = rng.random(len(lengths)) + 0.5
    es = refexec.ReferenceEStep(lengths, W, lambda w: scores[w], background, zoops, 0)
    ll, w, gamma = es.getNewWeights(weights)
    # the same E-step with a StrandModel as motif component (forward weight 0.6): a window's score is
    # Normalisation.getLogSum(logW0 + forward, logW1 + reverse complement) (jstacs AbstractMixtureModel.java:1158-1167)
    rev = [fg.score(aln[off[i] + j:off[i] + j + W], True)[0] for i in range(len(lengths)) for j in range(lengths[i] - W + 1)]
    slw = np.log([0.6, 0.4])
    both = [refexec.reference_log_sum([slw[0] + a, slw[1] + b]) for a, b in zip(scores, rev)]
    s_ll, s_w, s_gamma = refexec.ReferenceEStep(lengths, W, lambda w: both[w], background, zoops, 0).getNewWeights(weights)
    strand = dict(strand_forward_weight=0.6, strand_rev_scores=np.array(rev), strand_ll=s_ll, strand_w=np.array(s_w), strand_gamma=np.array(s_gamma))
    return dict(**strand, estep_newick=newick, estep_W=W, estep_zoops=zoops, estep_pwm=np.concatenate([np.ravel(p) for p in pwm]),
                estep_codes=aln, estep_col_offsets=off, estep_weights=weights, estep_scores=np.array(scores),
                estep_bg_columns=np.array(minus_fe), estep_ll=ll, estep_w=np.array(w), estep_gamma=np.array(gamma))


def objective_cases(rng):
    """The M-step objective f(lambda) = -sum_u w_u F(q_u fixed; theta(lambda)) and its forward-difference gradient (eps = 1e-4)
    for one motif position: MatrixLinearisation.lambda2pi, Normalisation.logSumNormalisation, MeanFieldForBayesNet.calcFreeEnergy
    and NumericalDifferentiableFunction.evaluateGradientOfFunction executed from the reference (refexec.ReferenceObjective)."""
    for i, (tree, order, W, pos) in enumerate((("cat5", 0, 4, 1), ("mixed6", 0, 4, 3), ("cat5", 1, 4, 0), ("cat5", 1, 4, 2), ("mixed6", 1, 4, 3))):
        newick = TREES[tree]
        S = len(npr.parse_newick(newick)[3])
        pwm = synth.random_pwm(W, order, 500 + i)
        net = npr.Net.motif(newick, W, order)
        for t, p in enumerate(pwm):
            net.set_pi(t, p)
        net.build_tables()
        windows = np.stack([codes(rng, W, S, 0.1 if i % 2 else 0.0) for _ in range(6)])
        weights = rng.random(6) + 0.1
        obj = refexec.ReferenceObjective(net, list(windows), weights, pos)
        lam = np.log(np.ravel(pwm[pos])) + 0.3 * rng.normal(size=np.ravel(pwm[pos]).size)
        f = obj.evaluateFunction(list(lam))
        g = obj.evaluateGradientOfFunction(list(lam))
        yield {f"obj{i}_newick": newick, f"obj{i}_order": order, f"obj{i}_W": W, f"obj{i}_pos": pos,
               f"obj{i}_pwm": np.concatenate([np.ravel(p) for p in pwm]), f"obj{i}_windows": windows, f"obj{i}_weights": weights,
               f"obj{i}_lambda": lam, f"obj{i}_f": f, f"obj{i}_grad": np.array(g)}


def selector_cases(rng):
    """SequenceSpecificDataSelector.prepareForTraining on the posterior of one alignment's windows (refexec.ReferenceSelector)."""
    for i, (n, t, q) in enumerate(((40, 0.8, 0.0), (40, 0.8, 0.01), (7, 0.5, 0.0), (1, 0.8, 0.0), (100, 0.95, 0.002), (25, 0.8, 0.0))):
        w = rng.dirichlet(np.ones(n) * 0.3) * (0.4 + 0.6 * rng.random())
        idx, ww = refexec.ReferenceSelector(list(w)).prepareForTraining(t, q)
        yield {f"sel{i}_weights": w, f"sel{i}_t": t, f"sel{i}_q": q, f"sel{i}_index": np.array(idx, np.int32), f"sel{i}_picked": np.array(ww)}


def background_driver_cases(rng):
    """PhyloBackground.getLogProbFor executed whole (refexec.ReferenceBackground: the reference's driver, initObservation, init,
    optimizeByNormalisation, calcFreeEnergy over ONE shared net): an ordered list of range calls on one alignment -- ranges
    shorter than order + 1 columns depend on what the calls before them left in the trailing trees -- and a ZOOPS E-step with
    such a background, in which getNewWeights issues the calls in the reference's order."""
    newick, W = TREES["cat5"], 5
    for order in (1, 2):
        net = npr.Net.background(newick, order)
        pis = [rng.dirichlet(np.ones(4), size=4 ** t) for t in range(order + 1)]
        for t, p in enumerate(pis):
            net.set_pi(t, p)
        net.build_tables()
        rb = refexec.ReferenceBackground(net)
        aln = codes(rng, 12, 5, 0.1)
        seq = refexec._MDSeq(aln)
        calls = [(0, 11), (3, 7), (2, 2), (5, 5 + order), (4, 4 + order - 1), (0, 0), (9, 11), (6, 5), (1, 1 + order), (11, 11)]
        values = [rb.getLogProbFor(seq, s, e) for s, e in calls]
        # the E-step
        lengths = [W, 11, 8]
        ealn = codes(rng, sum(lengths), 5, 0.06)
        off = np.concatenate([[0], np.cumsum(lengths)])
        pwm = synth.random_pwm(W, 0, 700 + order)
        mnet = npr.Net.motif(newick, W, 0)
        for t, p in enumerate(pwm):
            mnet.set_pi(t, p)
        mnet.build_tables()
        fg = refexec.ReferenceMeanField(mnet)
        scores = [fg.score(ealn[off[i] + j:off[i] + j + W])[0] for i in range(len(lengths)) for j in range(lengths[i] - W + 1)]
        rb = refexec.ReferenceBackground(net)
        seqs = [refexec._MDSeq(ealn[off[i]:off[i + 1]]) for i in range(len(lengths))]
        weights = rng.random(len(lengths)) + 0.5
        es = refexec.ReferenceEStep(lengths, W, lambda w: scores[w], lambda i, s, e: rb.getLogProbFor(seqs[i], s, e), 0.8, order)
        ll, w, gamma = es.getNewWeights(weights)
        k = f"bgd{order}_"
        yield {k + "newick": newick, k + "pi": np.concatenate([np.ravel(p) for p in pis]), k + "codes": aln, k + "calls": np.array(calls),
               k + "values": np.array(values, dtype=np.float64), k + "W": W, k + "pwm": np.concatenate([np.ravel(p) for p in pwm]),
               k + "estep_codes": ealn, k + "estep_col_offsets": off, k + "estep_weights": weights, k + "estep_ll": ll,
               k + "estep_w": np.array(w), k + "estep_gamma": np.array(gamma)}


def alignment_based_cases(rng):
    """The non-phylogenetic motif model (models/AlignmentBasedModel.java): window score = sum over species rows and motif
    positions of getLnCondProb (:176-228, :232-248), executed from the reference (refexec.ReferenceAlignmentBasedModel)."""
    for order in (0, 1, 2):
        W, S = 5, 4
        cond = [rng.dirichlet(np.ones(4), size=4 ** min(p, order)).ravel() for p in range(W)]
        aln = codes(rng, 14, S, 0.2)
        model = refexec.ReferenceAlignmentBasedModel(cond, order)
        score = np.zeros(aln.shape[0] - W + 1)
        for j in range(score.size):
            for o in range(S):                            # getLogProbFor :243-245, getLnCondProb(sequence, posSeq) :176-182
                total = 0
                for i in range(W):
                    total += model.getLnCondProb(aln[:, o], j + i, i)
                score[j] += total
        # the background model of the same family: per column the sum over species rows of AlignmentBasedBGModel.getLnCondProb
        # (models/AlignmentBasedBGModel.java:173-222, :239-243)
        bcond = [rng.dirichlet(np.ones(4), size=4 ** p).ravel() for p in range(order + 1)]
        bmodel = refexec.ReferenceAlignmentBasedBackground(bcond, order)
        columns = np.zeros(aln.shape[0])
        for l in range(aln.shape[0]):
            for o in range(S):
                columns[l] += bmodel.getLnCondProb(aln[:, o], l)
        k = f"ab{order}_"
        yield {k + "cond": np.concatenate(cond), k + "W": W, k + "codes": aln, k + "score": score,
               k + "bg_cond": np.concatenate(bcond), k + "bg_columns": columns}


def edge_case(rng):
    """EdgeOptimizer.evaluateFunction (optimizing/branchlengths/EdgeOptimizer.java:44-74) executed from the reference: the
    objective over fixed mean fields as a function of the branch-length parameters of one motif position."""
    newick, W, pos = TREES["mixed6"], 3, 1
    pwm = synth.random_pwm(W, 0, 901)
    net = npr.Net.motif(newick, W, 0)
    for t, p in enumerate(pwm):
        net.set_pi(t, p)
    net.build_tables()
    windows = np.stack([codes(rng, W, 6, 0.1) for _ in range(5)])
    weights = rng.random(5) + 0.1
    obj = refexec.ReferenceEdgeObjective(net, list(windows), weights, pos)
    xs = np.stack([rng.normal(size=net.N - 1) - 1.0 for _ in range(3)] + [np.array([-30.0] * (net.N - 1)), np.array([30.0] + [0.0] * (net.N - 2))])
    return dict(edge_newick=newick, edge_W=W, edge_pos=pos, edge_pwm=np.concatenate([np.ravel(p) for p in pwm]), edge_windows=windows,
                edge_weights=weights, edge_x=xs, edge_f=np.array([obj.evaluateEdges(x) for x in xs]))


def main():
    rng = np.random.default_rng(20261020)
    out, n = {}, 0
    for case in list(motif_cases(rng)) + list(background_cases(rng)):
        for k, v in case.items():
            out[f"c{n}_{k}"] = np.array(v)
        n += 1
    out["n_cases"] = np.array(n)
    out.update({k: np.array(v) for k, v in estep_case(rng).items()})
    n_obj = 0
    for case in objective_cases(rng):
        out.update({k: np.array(v) for k, v in case.items()})
        n_obj += 1
    out["n_objectives"] = np.array(n_obj)
    n_sel = 0
    for case in selector_cases(rng):
        out.update({k: np.array(v) for k, v in case.items()})
        n_sel += 1
    out["n_selectors"] = np.array(n_sel)
    for case in background_driver_cases(rng):
        out.update({k: np.array(v) for k, v in case.items()})
    for case in alignment_based_cases(rng):
        out.update({k: np.array(v) for k, v in case.items()})
    out.update({k: np.array(v) for k, v in edge_case(rng).items()})
    path = os.path.join(ROOT, "tests", "golden", "refexec_meanfield.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)} bytes, {n} cases, "
          f"{sum(out[f'c{i}_score'].size for i in range(n))} scores from the reference's statements")


if __name__ == "__main__":
    main()
