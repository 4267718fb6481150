"""Host-side mirror of the reference's model interface for the hot path, backed by libphyfoo.so.

Same names and argument meaning as the Java classes; the difference is granularity: where the
reference scores one window per call (`getLogProbFor(Sequence, start, end)`, ~10^7 calls per E-step),
these classes score the whole sample in one library call and serve per-window queries from the result.

  PhyloBayesModel            models/PhyloBayesModel.java        (motif; order 0 or 1)
  PhyloBackground            models/PhyloBackground.java        (flanking / background; order k)
  StrandModel                jstacs de/jstacs/models/mixture/StrandModel.java (forward + reverse complement)
  SingleHiddenMotifMixture   jstacs de/jstacs/models/mixture/motif/SingleHiddenMotifMixture.java (ZOOPS E-step)
"""
import contextlib
import math

import numpy as np

from . import _abi, core, optimize
from .data import PhyloSample, combineBranchLengths
from .newick import parse_newick


class _PhyloModel:
    """models/AbstractPhyloModel.java + PhyloPreparedAbstractModel.java: parameter access."""

    kind = None
    # static switches of the reference (models/PhyloBayesModel.java:43-69)
    CALC_LOGLIKELIHOOD = False
    EVOL_MODEL = _abi.EVOL_F81ALPHA   # evolution/EvolModel.java:7 MODEL_CLASS
    HKY_RATIO = 0.0                   # EVOL_HKY_RATIO only: transversion / transition ratio (the "HKY:<ratio>" of EvolModel.java:49-51)
    # mean-field sweep limits and stop tolerance (algorithm/MeanFieldForBayesNet.java:102-107,205); 0 = the reference's
    # constants (100, 2, 1e-5).  Read when the device model is created.
    MAX_SWEEPS = 0
    MIN_SWEEPS = 0
    TOL = 0.0

    def __init__(self, length, order, newick, ctx=None):
        self._ctx = ctx                     # resolved on first use: parameters can be set and files read without a device
        self.length, self.order, self.newick = int(length), int(order), newick
        self.tree = parse_newick(newick)
        self.n_trees = self.length if self.kind == _abi.MODEL_FG else self.order + 1
        # fillParameter(true): uniform start here; callers set parameters explicitly (the reference draws
        # them from an unseeded RNG, util/Util.java:35,108-121)
        self._pi = [np.full((self._ctx_rows(t), 4), 0.25) for t in range(self.n_trees)]
        self._branch = np.tile(self.tree.branch, (self.n_trees, 1))
        self._dev = None
        self._mode = None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = core.default_context()
        return self._ctx

    def _ctx_rows(self, t):
        if self.kind == _abi.MODEL_FG:
            return 4 if (self.order >= 1 and t > 0) else 1      # PhyloBayesModel.java:338-339
        return 4 ** min(self.order, t)                           # PhyloBackground.java:386-394

    def getLength(self):
        return self.length

    def getOrder(self):
        return self.order

    # ---- parameters (models/AbstractPhyloModel.java:131-134,160-163; util/MatrixLinearisation.java:89-110) ----
    def getCondProb(self, i):
        return self._pi[i].ravel().copy()

    def getCondProbs(self):
        return [self.getCondProb(i) for i in range(self.n_trees)]

    def getLnCondProb(self, i):
        return np.log(self.getCondProb(i))

    def getLnCondProbs(self):
        return [self.getLnCondProb(i) for i in range(self.n_trees)]

    def setCondProb(self, condProb, position):
        cp = np.asarray(condProb, np.float64).reshape(self._pi[position].shape)
        self._pi[position] = cp.copy()
        if self._dev is not None:
            self._dev.set_pi(position, cp)

    def setCondProbs(self, condProbs):
        for i, cp in enumerate(condProbs):
            self.setCondProb(cp, i)

    def setLnCondProb(self, lam, position):
        self.setCondProb(np.exp(np.asarray(lam, np.float64)), position)

    def setLnCondProbs(self, lambdas):
        for i, lam in enumerate(lambdas):
            self.setLnCondProb(lam, i)

    def reinitEdgelengths(self, *trees):
        """models/AbstractPhyloModel.java:77-91: one Newick string for all positions, or one per position."""
        if len(trees) not in (1, self.n_trees):
            raise ValueError("The number of given trees must be equal to length of this model")
        for t in range(self.n_trees):
            tr = parse_newick(trees[0] if len(trees) == 1 else trees[t])
            if tr.n_nodes != self.tree.n_nodes or (tr.parent != self.tree.parent).any():
                raise ValueError("tree topology differs from the model's")
            self._branch[t] = tr.branch
            if self._dev is not None:
                for k in range(1, tr.n_nodes):
                    self._dev.set_branch(t, k, tr.branch[k])

    def getNewickString(self, position=0):
        """getBNH().getVirtualTree(position).getNewickString() (bayesNet/VirtualTree.java:63-66,208-234): branch lengths
        rounded to three decimals."""
        from . import xmlio
        return xmlio.newick_string(self.tree, self._branch[position])

    def setBranchLength(self, position, node, length):
        self._branch[position, node] = length
        if self._dev is not None:
            self._dev.set_branch(position, node, length)

    # ---- device handle ------------------------------------------------------------------------
    def mode(self):
        return _abi.MODE_EXACT if type(self).CALC_LOGLIKELIHOOD or _PhyloModel.CALC_LOGLIKELIHOOD else _abi.MODE_MEANFIELD

    def device_model(self):
        if self._dev is None:
            self._dev = core.DeviceModel(self.ctx, self.kind, self.length, self.order, self.tree, self._branch,
                                         self._pi, evol=type(self).EVOL_MODEL, mode=self.mode(), hky_ratio=type(self).HKY_RATIO,
                                         max_sweeps=self.MAX_SWEEPS, min_sweeps=self.MIN_SWEEPS, tol=self.TOL)
            self._mode = self.mode()
        elif self._mode != self.mode():
            self._dev.set_mode(self.mode())
            self._mode = self.mode()
        return self._dev

    def _check(self, sample):
        if sample.n_species != self.tree.n_species:
            raise ValueError("The sample's number of species differs from the tree's number of leaves")

    # ---- per-alignment learned topologies (models/PhyloBackground.java:250-253,301-303; PhyloBayesModel.java:233-247,263-265) ----
    @contextlib.contextmanager
    def _under_topology(self, topology):
        """The model with the branch lengths of `topology` (an alignment's learnedTopology) for the duration of a call, as the
        reference's getLogProbFor does per sequence.  Afterwards the branch lengths are those of the Newick strings taken
        before, i.e. ROUNDED to three decimals (getNewickString) -- the reference's side effect, kept; a background model
        resets every position to the string of position 0.  topology = None: nothing happens."""
        if topology is None:
            yield
            return
        old = [self.getNewickString(t) for t in range(self.n_trees)]
        if self.kind == _abi.MODEL_FG:
            self.reinitEdgelengths(*[combineBranchLengths(topology, o, 1.0) for o in old])
        else:
            self.reinitEdgelengths(topology)
        try:
            yield
        finally:
            if self.kind == _abi.MODEL_FG:
                self.reinitEdgelengths(*old)
            else:
                self.reinitEdgelengths(old[0])

    def _per_topology(self, sample, fn):
        """fn(sample) for a sample without learned topologies; otherwise fn over every run of consecutive alignments with the
        same learnedTopology (each under that topology, in alignment order -- one batch call per run) and the list of the
        results.  Returns (results, split) with split = False when fn ran once on the whole sample."""
        runs = topology_runs(sample)
        if runs is None:
            return [fn(sample)], False
        out = []
        for lo, hi, topo in runs:
            with self._under_topology(topo):
                out.append(fn(sample.slice(lo, hi)))
        return out, True


def topology_runs(sample):
    """None when no alignment of the sample carries a learnedTopology, else [(lo, hi, topology or None)] over runs of
    consecutive alignments with the same one."""
    topo = getattr(sample, "learnedTopology", None)
    if not topo or all(t is None for t in topo):
        return None
    runs, i, n = [], 0, sample.n_alignments
    while i < n:
        j = i + 1
        while j < n and topo[j] == topo[i]:
            j += 1
        runs.append((i, j, topo[i]))
        i = j
    return runs


class PhyloBayesModel(_PhyloModel):
    """The phylogenetic motif model (phylogenetic PWM for order 0, phylogenetic Bayesian network for order 1)."""

    kind = _abi.MODEL_FG

    def __init__(self, motifLength, order, newickString, ctx=None):
        if order not in (0, 1):
            raise ValueError("motif order must be 0 or 1 (models/PhyloBayesModel.java:363)")
        super().__init__(motifLength, order, newickString, ctx)
        self.last_stats = None

    def getMotifLength(self):
        return self.length

    def toXML(self):
        """models/PhyloBayesModel.java:291-299 (branch lengths are written with three decimals, like the reference)."""
        from . import xmlio
        return xmlio.phylo_bayes_model_to_xml(self)

    @classmethod
    def fromXML(cls, xml, ctx=None):
        """PhyloBayesModel(StringBuffer) (models/PhyloBayesModel.java:300-309)."""
        from . import xmlio
        return xmlio.read_phylo_bayes_model(xml, ctx)

    def getLogProbFor(self, sample: PhyloSample, strand=0, want_sweeps=False):
        """Scores of every window of every alignment (alignment-major, offsets ascending) =
        PhyloBayesModel.getLogProbFor(seq, j, j+W-1) for all (seq, j)  (models/PhyloBayesModel.java:222-267)."""
        self._check(sample)

        def score(smp):
            out, sw, st = core.score_windows(self.ctx, smp.device(self.ctx), self.device_model(), strand, want_sweeps)
            self.last_stats = st
            return out, sw
        parts, split = self._per_topology(sample, score)          # alignments with a learnedTopology: under their own branch lengths
        out = np.concatenate([p[0] for p in parts]) if split else parts[0][0]
        sw = (np.concatenate([p[1] for p in parts]) if split else parts[0][1]) if want_sweeps else None
        return (out, sw) if want_sweeps else out


    # ---- M-step (models/PhyloBayesModel.java:122-219) -------------------------------------------------------------
    SKIP_TRAINING_STEPS = 1            # :57
    PROB_THRESH_FOR_WEIGHTS = 0.8      # models/PhyloPreparedAbstractModel.java:20
    MOTIF_QUALITY_THRESHOLD = 0.0      # :21
    OPTIMIZE_IN_TOTO = False           # :63
    # True (default; orders 0 and 1): one device pass collects the expected counts of the table entries, every evaluation
    # of the objective is then a dot product on the host inside libphyfoo.so and a whole M-step is one library call
    # (phyfoo_ss_optimize).  False: every evaluation is a pass over the selected windows on the device, the reference's own
    # evaluation order.  Same function, sums re-associated (1e-12 relative along a whole optimisation:
    # tests/test_gpu_mstep.py, tests/test_gpu_order1.py).
    MSTEP_SUFFICIENT_STATISTICS = True

    def train(self, sample: PhyloSample, weights, rng=None, rev_weights=None, sharded=False):
        """One M-step: select the top-posterior windows per alignment (io/SequenceSpecificDataSelector.java:37-99), fix
        their mean fields (one GPU pass), then optimise the parameters position by position in random order with BFGS on the
        forward-difference gradient (optimizing/FE_byParameter_ForBayesNodeJstacs.java:61-83).  weights: gamma per window, or
        None for the gamma of the most recent E-step still resident on the device.  rev_weights: when the caller is a
        StrandModel, the weights of the reverse-complement copies of the windows; the selector then thresholds the two strands
        of an alignment independently, as the reference does because it groups windows by their parent object
        (io/SequenceSpecificDataSelector.java:45-49).  sharded=True: `sample` is this rank's
        shard of the alignments (torch.distributed is initialised): every objective evaluation is all-reduced, and the random
        order of the positions comes from rank 0, so all ranks take the same optimiser steps.  Returns a dict of statistics."""
        self._check(sample)
        self.trainingStep = getattr(self, "trainingStep", 0)
        if self.trainingStep < self.SKIP_TRAINING_STEPS:
            self.trainingStep += 1
            return {"skipped": True}
        rng = rng or np.random.default_rng()
        ds = sample.device(self.ctx)
        sa, so, sw = core.select(self.ctx, ds, self.length, weights, self.PROB_THRESH_FOR_WEIGHTS, self.MOTIF_QUALITY_THRESHOLD)
        ss = None
        if rev_weights is not None:
            ra, ro, rw = core.select(self.ctx, ds, self.length, rev_weights, self.PROB_THRESH_FOR_WEIGHTS,
                                     self.MOTIF_QUALITY_THRESHOLD)
            ss = np.concatenate([np.zeros(sa.size, np.uint8), np.ones(ra.size, np.uint8)])
            sa, so, sw = np.concatenate([sa, ra]), np.concatenate([so, ro]), np.concatenate([sw, rw])
        return self.train_selected(sample, sa, so, sw, ss, rng, sharded)

    def train_selected(self, sample: PhyloSample, sa, so, sw, ss=None, rng=None, sharded=False):
        """The M-step proper on an explicit selection (alignment, offset, weight[, strand]) -- what train() does after the
        selector; a mixture of several motifs (MultiMotifMixture) calls it with the selection of each component."""
        rng = rng or np.random.default_rng()
        ds = sample.device(self.ctx)
        fe = core.FESet(self.ctx, ds, self.device_model(), sa, so, sw, ss)
        stats = {"skipped": False, "selected": int(sa.size), "evaluations": 0, "f_before": None, "f_after": None}
        local = fe
        from . import sharding
        order = rng.permutation(self.length).astype(np.int64)    # util/PWMUtil.java:131-144
        if sharded:
            # every rank holds a shard of the alignments: one all-reduce of [f, f_1..f_dim] per evaluation, after which all
            # ranks see the same objective and take the same optimiser steps
            fe = sharding.ShardedObjective(local, self.ctx.device)
            sharding.broadcast_host_(order)
        try:
            if self.MSTEP_SUFFICIENT_STATISTICS and self.order <= 1 and not self.OPTIMIZE_IN_TOTO:
                # one device pass for the expected counts (all-reduced when sharded), then the position-wise BFGS on host dot
                # products inside the library
                if sharded:
                    n_entries = (16 + 32 * (self.tree.n_nodes - 1)) if self.order == 1 else (4 + 16 * (self.tree.n_nodes - 1))
                    counts, entropy = local.suffstats(self.length, n_entries)
                    counts, entropy = sharding.allreduce_counts_(counts, entropy)
                    desc, keep = _abi.make_desc(self.kind, self.length, self.order, self.tree, self._branch, self._pi, type(self).EVOL_MODEL,
                                                _abi.MODE_MEANFIELD, type(self).HKY_RATIO)
                    obj = core.SSObjective(desc=desc, counts=counts, entropy=entropy)
                    del keep
                else:
                    obj = core.SSObjective(local)
                stats["f_before"], stats["f_after"], stats["evaluations"] = obj.optimize(order)
                for t in range(self.length):
                    self.setCondProb(obj.get_pi(t, self._pi[t].size).reshape(self._pi[t].shape), t)
                obj.free()
                self.trainingStep += 1
                return stats
            if self.OPTIMIZE_IN_TOTO:
                lam0 = np.concatenate([np.log(p.ravel()) for p in self._pi])
                stats["f_before"] = fe.eval(-1, lam0)
                x, f, it, ev = optimize.quasi_newton_bfgs(lambda z: -fe.eval(-1, z), lambda z: _neg(fe.grad(-1, z)), lam0,
                                                          optimize.TC_MOTIF_ALL_POSITIONS())
                stats["f_after"] = fe.eval(-1, x)
                stats["evaluations"] += ev
            else:
                for pos in order:
                    pos = int(pos)
                    lam0 = np.log(self._pi[pos].ravel())
                    f0 = fe.eval(pos, lam0)
                    if stats["f_before"] is None:
                        stats["f_before"] = f0
                    x, f, it, ev = optimize.quasi_newton_bfgs(lambda z: -fe.eval(pos, z), lambda z: _neg(fe.grad(pos, z)), lam0,
                                                              optimize.TC_MOTIF_SEPARATED_POSITIONS())
                    stats["f_after"] = fe.eval(pos, x)            # leaves the device model at x
                    stats["evaluations"] += ev
            for t in range(self.n_trees):
                self._pi[t] = self._dev.get_pi(t, self._pi[t].size).reshape(self._pi[t].shape)
        finally:
            local.free()
        self.trainingStep += 1
        return stats


    EDGE_LEARNING = 0                  # models/PhyloBayesModel.java:47: 0 off, 1 per position, 2 one set for all positions
    EDGE_LEARNING_ONE_LENGTH = False

    def trainEdges(self, sample: PhyloSample, weights, globally=None, equal=None):
        """Branch-length learning on the fixed-q objective (optimizing/branchlengths/EdgeOptimizer.java:44-125,
        GlobalEdgeOptimizer.java): logit-parametrised branch lengths, BFGS until the objective changes by less than 1e-3.
        One device pass collects the sufficient statistics; the optimisation itself runs on the host.  Orders 0 and 1."""
        if self.order == 1:
            return self._trainEdges_order1(sample, weights, globally, equal)
        from . import evolution
        globally = (self.EDGE_LEARNING == 2) if globally is None else globally
        equal = self.EDGE_LEARNING_ONE_LENGTH if equal is None else equal
        ds = sample.device(self.ctx)
        sa, so, sw = core.select(self.ctx, ds, self.length, weights, self.PROB_THRESH_FOR_WEIGHTS, self.MOTIF_QUALITY_THRESHOLD)
        fe = core.FESet(self.ctx, ds, self.device_model(), sa, so, sw)
        try:
            obj = optimize.SuffStatObjective(fe, self._pi, self._branch, type(self).EVOL_MODEL, type(self).HKY_RATIO)
        finally:
            fe.free()
        before = obj.value()
        evals = 0
        for t in ([0] if globally else range(self.length)):
            b = self._branch[t, 1:]
            x0 = evolution.logit_from_branch(np.array([b.mean()]) if equal else b)
            fun = lambda z: -obj.eval_edges(t, z, equal, globally)          # noqa: E731
            x, f, it, ev = optimize.quasi_newton_bfgs(fun, lambda z: obj.grad(fun, z), x0,
                                                      optimize.SMALL_DIFFERENCE_OF_FUNCTIONS())
            # (EdgeOptimizer.startOptimizing does not write the optimizer's result back, :104-121: the branch lengths stay at
            # the last point the optimizer evaluated -- the last perturbation of its last gradient, x + 1e-4 e_last)
            evals += ev
        for t in range(self.length):
            for k in range(1, self.tree.n_nodes):
                self.setBranchLength(t, k, float(obj.branch[t, k]))
        return {"f_before": before, "f_after": obj.value(), "evaluations": evals, "selected": int(sa.size)}


def _neg(fg):
    return -fg[0], -fg[1]


def _learn_edges(model, obj, globally, equal):
    """EdgeOptimizer.startOptimizing for every position (one run for all positions when `globally`) on the library's
    sufficient-statistics objective `obj` (phyfoo_ss_eval_branches): logit-parametrised branch lengths from the model's current
    ones (their mean when `equal`), BFGS on the forward-difference gradient until the objective changes by less than 1e-3
    (optimizing/branchlengths/EdgeOptimizer.java:44-125, GlobalEdgeOptimizer.java:33-103).  Writes the learned lengths -- those
    of the optimizer's last evaluation, as the reference leaves them -- into the model; returns (f_before, f_after, evaluations)."""
    from . import evolution
    N = model.tree.n_nodes
    before, evals, learned = obj.value(), 0, {}

    def eval_edges(t, z):
        alpha = evolution.branch_from_logit(z)
        b = np.zeros(N)
        b[1:] = alpha[0] if equal else alpha
        learned[t] = b
        return obj.eval_branches(-1 if globally else t, b)

    def fd(fun, x, eps=1e-4):                  # jstacs NumericalDifferentiableFunction.java:64-84 (no closing evaluation)
        x = np.array(x, np.float64)
        f0, g = fun(x), np.zeros(x.size)
        for i in range(x.size):
            keep = x[i]
            x[i] += eps
            g[i] = (fun(x) - f0) / eps
            x[i] = keep
        return f0, g

    for t in ([0] if globally else range(model.n_trees)):
        b = model._branch[t, 1:]
        x0 = evolution.logit_from_branch(np.array([b.mean()]) if equal else b)
        fun = lambda z: -eval_edges(t, z)                                   # noqa: E731
        x, f, it, ev = optimize.quasi_newton_bfgs(fun, lambda z: fd(fun, z), x0, optimize.SMALL_DIFFERENCE_OF_FUNCTIONS())
        # EdgeOptimizer.startOptimizing does not write the optimizer's result back (:104-121): the branch lengths stay at the
        # last point the optimizer evaluated -- the last perturbation of its last gradient, x + 1e-4 e_last in logit space
        evals += ev
    after = obj.value()
    for t in range(model.n_trees):
        b = learned[0] if globally else learned[t]
        for k in range(1, N):
            model.setBranchLength(t, k, float(b[k]))
    return before, after, evals


def _train_edges_order1(self, sample, weights, globally=None, equal=None):
    """trainEdges for the order-1 network: the same optimisation on the library's sufficient-statistics objective
    (phyfoo_ss_eval_branches over the expected counts of fe1_count_kernel)."""
    globally = (self.EDGE_LEARNING == 2) if globally is None else globally
    equal = self.EDGE_LEARNING_ONE_LENGTH if equal is None else equal
    ds = sample.device(self.ctx)
    sa, so, sw = core.select(self.ctx, ds, self.length, weights, self.PROB_THRESH_FOR_WEIGHTS, self.MOTIF_QUALITY_THRESHOLD)
    fe = core.FESet(self.ctx, ds, self.device_model(), sa, so, sw)
    try:
        obj = core.SSObjective(fe)
    finally:
        fe.free()
    try:
        before, after, evals = _learn_edges(self, obj, globally, equal)
    finally:
        obj.free()
    return {"f_before": before, "f_after": after, "evaluations": evals, "selected": int(sa.size)}


PhyloBayesModel._trainEdges_order1 = _train_edges_order1


class PhyloBackground(_PhyloModel):
    """The phylogenetic background / flanking model (order-k Markov chain over trees)."""

    kind = _abi.MODEL_BG

    def __init__(self, order, newickString, ctx=None):
        super().__init__(0, order, newickString, ctx)
        self.last_stats = None

    def toXML(self):
        """models/PhyloBackground.java:330-336."""
        from . import xmlio
        return xmlio.phylo_background_to_xml(self)

    @classmethod
    def fromXML(cls, xml, ctx=None):
        from . import xmlio
        return xmlio.read_phylo_background(xml, ctx)

    def getColumnLogProbs(self, sample: PhyloSample, want_sweeps=False):
        """Per-column terms whose sum over [start, end] is getLogProbFor(seq, start, end)
        (models/PhyloBackground.java:241-305; order 0: independent columns)."""
        self._check(sample)

        def score(smp):
            out, sw, st = core.bg_columns(self.ctx, smp.device(self.ctx), self.device_model(), want_sweeps)
            self.last_stats = st
            return out, sw
        parts, split = self._per_topology(sample, score)
        out = np.concatenate([p[0] for p in parts]) if split else parts[0][0]
        sw = (np.concatenate([p[1] for p in parts]) if split else parts[0][1]) if want_sweeps else None
        return (out, sw) if want_sweeps else out

    def getRangeTerms(self, sample: PhyloSample, want_sweeps=False):
        """Background of order k >= 1: (cond, first) per column with getLogProbFor(seq, s, e) = first[s] + sum(cond[s+k..e])
        for ranges of at least k + 1 columns (models/PhyloBackground.java:273-297); see phyfoo_bg_range_terms.  Order 1 runs
        the order-1 network kernel (phyfoo_bg_columns_order1), higher orders the generic net kernel."""
        if self.order < 1:
            raise ValueError("getRangeTerms is for backgrounds of order >= 1; order 0 uses getColumnLogProbs")
        self._check(sample)

        def terms(smp):
            ds = smp.device(self.ctx)
            cond, first = np.zeros(ds.n_cols), np.zeros(ds.n_cols)
            if self.order == 1 and not want_sweeps and not type(self).GENERIC_NET:
                rc = core.lib().phyfoo_bg_columns_order1(self.ctx.h, ds.h, self.device_model().h, core._p(cond, core.C.c_double),
                                                         core._p(first, core.C.c_double))
                sw = None
            else:
                sw = np.zeros(max(int(ds.windows(self.order + 1)), 1), np.int32) if want_sweeps else None
                rc = core.lib().phyfoo_bg_range_terms(self.ctx.h, ds.h, self.device_model().h, core._p(cond, core.C.c_double),
                                                      core._p(first, core.C.c_double), core._p(sw, core.C.c_int32))
                if sw is not None:
                    sw = sw[:int(ds.windows(self.order + 1))]
            core.raise_for(rc, self.ctx.h)
            return (cond, first, sw) if want_sweeps else (cond, first)
        parts, split = self._per_topology(sample, terms)
        if not split:
            return parts[0]
        return tuple(np.concatenate([p[i] for p in parts]) for i in range(len(parts[0])))

    GENERIC_NET = False                # True: order 1 through the generic net kernel as well (cross-check)

    def getLogProbForRange(self, sample: PhyloSample, aln, start, end):
        """getLogProbFor(seq, start, end) for one alignment (end inclusive)."""
        if end < start:
            return 0.0
        off = int(sample.col_offsets[aln])
        if self.order == 0:
            return float(self.getColumnLogProbs(sample)[off + start: off + end + 1].sum())
        k = self.order
        if end - start < k:
            raise ValueError("a range shorter than order + 1 columns has no defined value in the reference (its trailing trees "
                             "keep the observation of an earlier call)")
        cond, first = self.getRangeTerms(sample)
        return float(first[off + start] + cond[off + start + k: off + end + 1].sum())

    def getLogProbFor(self, sample: PhyloSample):
        """double[] getLogProbFor(Sample) (jstacs AbstractModel.java:217-240): one value per alignment."""
        off = sample.col_offsets
        if self.order >= 1:
            k = self.order
            cond, first = self.getRangeTerms(sample)
            return np.array([first[off[i]] + cond[off[i] + k:off[i + 1]].sum() for i in range(sample.n_alignments)])
        cols = self.getColumnLogProbs(sample)
        return np.array([cols[off[i]:off[i + 1]].sum() for i in range(sample.n_alignments)])


    TRAIN_MOTIF = True                 # models/PhyloBackground.java:41
    EDGE_LEARNING = False              # :43
    EDGE_LEARNING_ONE_LENGTH = False   # models/AbstractPhyloModel.java:25
    # True: the evaluations of the optimisers are host dot products over the expected counts of ONE device pass
    # (phyfoo_ss_*); False: a device pass over every column per evaluation of the pi objective (the reference's order of work)
    MSTEP_SUFFICIENT_STATISTICS = True

    def train(self, sample: PhyloSample, weights=None):
        """models/PhyloBackground.java:111-176 (order 0; order >= 1: _train_background_order_k below): objective over every
        column of every alignment with per-alignment
        weights (extendWeights :178-186 reads one weight per alignment of `sample`, so a longer vector is cut), q fixed after
        one mean-field pass; TRAIN_MOTIF: BFGS on pi until the objective changes by less than 1e-3; EDGE_LEARNING: then the
        branch lengths (EdgeOptimizer, per position, EDGE_LEARNING_ONE_LENGTH = one length for all branches) on the same
        fixed q."""
        self._check(sample)
        if weights is not None:
            weights = np.ascontiguousarray(np.asarray(weights, np.float64)[:sample.n_alignments])
        stats = {"f_before": None, "f_after": None, "evaluations": 0, "iterations": 0}
        if self.order >= 1:
            return self._train_order_k(sample, weights, stats)
        fe = core.FESet(self.ctx, sample.device(self.ctx), self.device_model(), aln_weights=weights)
        obj = None
        try:
            lam0 = np.log(self._pi[0].ravel())
            if type(self).MSTEP_SUFFICIENT_STATISTICS or type(self).EDGE_LEARNING:
                obj = core.SSObjective(fe)
            if type(self).TRAIN_MOTIF:
                ev_obj = obj if type(self).MSTEP_SUFFICIENT_STATISTICS else fe
                stats["f_before"] = ev_obj.eval(0, lam0)
                x, f, it, ev = optimize.quasi_newton_bfgs(lambda z: -ev_obj.eval(0, z), lambda z: _neg(ev_obj.grad(0, z)), lam0,
                                                          optimize.SMALL_DIFFERENCE_OF_FUNCTIONS())
                stats["f_after"] = ev_obj.eval(0, x)
                stats["evaluations"], stats["iterations"] = ev, it
                pi = ev_obj.get_pi(0, 4) if ev_obj is obj else self._dev.get_pi(0, 4)
                self.setCondProb(pi.reshape(1, 4), 0)
                if obj is not None and ev_obj is not obj:
                    obj.eval(0, np.log(pi))                      # the edge objective continues from the new pi
            if type(self).EDGE_LEARNING:
                before, after, evals = _learn_edges(self, obj, False, type(self).EDGE_LEARNING_ONE_LENGTH)
                stats["edges_f_before"], stats["edges_f_after"] = before, after
                stats["evaluations"] += evals
                if stats["f_before"] is None:
                    stats["f_before"] = before
                stats["f_after"] = after
        finally:
            if obj is not None:
                obj.free()
            fe.free()
        return stats


def _train_background_order_k(self, sample, weights, stats):
    """PhyloBackground.train for order k >= 1 (models/PhyloBackground.java:111-176): the mean fields of every (k + 1)-column
    window, fixed in one pass of the generic net kernel, and their expected counts (phyfoo_ss_create_bg); TRAIN_MOTIF: the
    stationary distributions position by position (position t: 4^min(t, k) context rows of four parameters,
    FE_byParameter_ForBayesNodeJstacs :61-83); EDGE_LEARNING: the branch lengths position by position.  All evaluations are
    host dot products inside the library.  `weights` keeps the reference's indexing: window number p carries the weight of the
    alignment that holds column number p."""
    obj = core.SSObjective(background=(self.ctx, sample.device(self.ctx), self.device_model(), weights))
    try:
        stats["f_before"] = obj.value()
        if type(self).TRAIN_MOTIF:
            for t in range(self.n_trees):
                lam0 = np.log(self._pi[t].ravel())
                x, f, it, ev = optimize.quasi_newton_bfgs(lambda z: -obj.eval(t, z), lambda z: _neg(obj.grad(t, z)), lam0,
                                                          optimize.SMALL_DIFFERENCE_OF_FUNCTIONS())
                obj.eval(t, x)                                   # initParameters(lambda), :78
                stats["evaluations"] += ev
                stats["iterations"] += it
                self.setCondProb(obj.get_pi(t, self._pi[t].size).reshape(self._pi[t].shape), t)
            stats["f_after"] = obj.value()
        if type(self).EDGE_LEARNING:
            before, after, evals = _learn_edges(self, obj, False, type(self).EDGE_LEARNING_ONE_LENGTH)
            stats["edges_f_before"], stats["edges_f_after"] = before, after
            stats["evaluations"] += evals
            stats["f_after"] = after
        if stats["f_after"] is None:
            stats["f_after"] = stats["f_before"]
    finally:
        obj.free()
    return stats


PhyloBackground._train_order_k = _train_background_order_k


class StrandModel:
    """jstacs StrandModel as PhyFoo builds it (models/ModelUtil.java:98-108): a mixture of a motif model and its reverse
    complement with estimated strand weights {forward, backward}, hyper-parameters {1, 1} and Dirichlet(1) initialisation."""

    def __init__(self, model: PhyloBayesModel, forwardProb=0.5, componentHyperParams=(1.0, 1.0), estimateComponentProbs=True,
                 alpha=1.0):
        self.model = model
        self.weights = np.array([forwardProb, 1.0 - forwardProb])
        self.hyper = np.asarray(componentHyperParams, np.float64)
        self.estimateComponentProbs = bool(estimateComponentProbs)
        self.alpha = float(alpha)
        self.sharded = False          # set by SingleHiddenMotifMixture.train(sharded=True)

    def getLength(self):
        return self.model.getLength()

    def logWeights(self):
        with np.errstate(divide="ignore"):
            return np.log(self.weights)

    def getLogProbFor(self, sample: PhyloSample):
        """logsumexp over the two strands per window (jstacs StrandModel.java:631-640)."""
        lw = self.logWeights()
        return np.logaddexp(lw[0] + self.model.getLogProbFor(sample, 0), lw[1] + self.model.getLogProbFor(sample, 1))

    def getNewWeights(self, sample: PhyloSample, dataWeights):
        """Strand E-step over every window (jstacs StrandModel.java:561-583): both strands scored on the GPU, the strand
        posterior times the window's data weight on the host (two multiplies per window).  Returns (forward weights, reverse
        weights, w[2] without the prior, L)."""
        lw = self.logWeights()
        a = lw[0] + self.model.getLogProbFor(sample, 0)
        b = lw[1] + self.model.getLogProbFor(sample, 1)
        tot = np.logaddexp(a, b)
        dw = np.ones_like(tot) if dataWeights is None else np.asarray(dataWeights, np.float64)
        with np.errstate(invalid="ignore"):
            pf = np.where(np.isfinite(tot), np.exp(a - tot), 0.5)
        f, r = pf * dw, (1.0 - pf) * dw
        return f, r, np.array([f.sum(), r.sum()]), float(np.sum(tot * dw))

    def _new_parameters(self, sample, f, r, w, rng):
        st = self.model.train(sample, f, rng=rng, rev_weights=r, sharded=self.sharded)
        if self.sharded:
            from . import sharding
            w = sharding.allreduce_host_(np.array(w, np.float64))
        if self.estimateComponentProbs:                        # AbstractMixtureModel.getNewComponentProbs :1542-1578
            v = np.asarray(w, np.float64) + self.hyper
            self.weights = v / v.sum()
        return st

    def doFirstIteration(self, sample: PhyloSample, dataWeights, rng=None):
        """jstacs StrandModel.java:530-557: each window's weight is split between the strands by a Dirichlet(alpha) draw, then
        one M-step.  This is what the enclosing mixture calls in its first M-step (AbstractMixtureModel.java:1024-1026)."""
        rng = rng or np.random.default_rng()
        dw = np.asarray(dataWeights, np.float64)
        split = rng.dirichlet(np.full(2, self.alpha), size=dw.size)
        f, r = dw * split[:, 0], dw * split[:, 1]
        return self._new_parameters(sample, f, r, np.array([f.sum(), r.sum()]), rng)

    def continueIterations(self, sample: PhyloSample, dataWeights, iterations=1, rng=None):
        """jstacs AbstractMixtureModel.java:946-977 as the enclosing mixture calls it from the second M-step on (:1028-1030):
        strand E-step, M-step, repeated `iterations` times, and a closing E-step whose likelihood is returned."""
        f, r, w, L = self.getNewWeights(sample, dataWeights)
        st = None
        for _ in range(iterations):
            st = self._new_parameters(sample, f, r, w, rng)
            f, r, w, L = self.getNewWeights(sample, dataWeights)
        return st, L


class SingleHiddenMotifMixture:
    """ZOOPS mixture of one motif model and a flanking model with a uniform position prior."""

    def __init__(self, motif, flanking: PhyloBackground, zoops=None, componentHyperParams=(1.0, 1.0)):
        self.strand = motif if isinstance(motif, StrandModel) else None
        self.motif = motif.model if self.strand else motif
        self.flanking = flanking
        if zoops is not None and 0 < zoops <= 1:
            # estimateComponentProbs = false (jstacs SingleHiddenMotifMixture.java:310-329, models/ModelUtil.java:110-121)
            self.weights = np.array([zoops, 1.0 - zoops])
            self.estimateComponentProbs = False
        else:
            self.weights = np.array([0.5, 0.5])
            self.estimateComponentProbs = True
        self.hyper = np.asarray(componentHyperParams, np.float64)
        self.last = None

    def getNumberOfMotifs(self):
        return 1   # jstacs SingleHiddenMotifMixture.java:716-718

    def toXML(self, **kw):
        """The reference's model file (jstacs AbstractMixtureModel.toXML :1628-1678; tools/Classification and
        models/ModelUtil.java:228-237 load it)."""
        from . import xmlio
        return xmlio.shm_to_xml(self, **kw)

    @classmethod
    def fromXML(cls, xml, ctx=None):
        from . import xmlio
        return xmlio.read_shm(xml, ctx)

    def getNewWeights(self, sample: PhyloSample, seqWeights=None, want_gamma=True, want_profile=False, out=None):
        """One E-step (jstacs SingleHiddenMotifMixture.java:499-566).  Returns a dict with
        gamma (seqweights[0] per window), w (sufficient statistics of the component weights), ll, argmax_off/strand."""
        with np.errstate(divide="ignore"):
            logw = np.log(self.weights)
        runs = topology_runs(sample)
        if runs is not None:
            res = self._estep_per_topology(sample, runs, logw, seqWeights, want_gamma, want_profile)
        else:
            res = core.zoops_estep(self.motif.ctx, sample.device(self.motif.ctx), self.motif.device_model(),
                                   self.flanking.device_model(), logw,
                                   None if self.strand is None else self.strand.logWeights(),
                                   seqWeights, want_gamma, want_profile, True, out)
        self.last = res
        return res

    def _estep_per_topology(self, sample, runs, logw, seqWeights, want_gamma, want_profile):
        """The E-step over a sample whose alignments carry learned topologies: one batch call per run of alignments with the
        same topology, motif and flanking model both under that run's branch lengths (the reference re-initialises them inside
        every getLogProbFor call); ll and w add up over the runs, the per-window and per-alignment results are concatenated.
        (The posteriors of such an E-step do not stay resident on the device for a following train(sample, None).)"""
        sl = None if self.strand is None else self.strand.logWeights()
        parts = []
        for lo, hi, topo in runs:
            with self.motif._under_topology(topo), self.flanking._under_topology(topo):
                sub = sample.slice(lo, hi)
                parts.append(core.zoops_estep(self.motif.ctx, sub.device(self.motif.ctx), self.motif.device_model(),
                                              self.flanking.device_model(), logw, sl,
                                              None if seqWeights is None else np.asarray(seqWeights, np.float64)[lo:hi],
                                              want_gamma, want_profile, True, None))
        cat = lambda k: None if parts[0][k] is None else np.concatenate([p[k] for p in parts])      # noqa: E731
        stats = dict(parts[-1]["stats"])
        for k in stats:
            if isinstance(stats[k], (int, float)):
                stats[k] = sum(p["stats"][k] for p in parts)
        return {"gamma": cat("gamma"), "profile": cat("profile"), "w": sum(p["w"] for p in parts), "ll": math.fsum(p["ll"] for p in parts),
                "argmax_off": cat("argmax_off"), "argmax_strand": cat("argmax_strand"), "stats": stats}

    def setComponentWeights(self, w):
        """AbstractMixtureModel.getNewComponentProbs (jstacs :1542-1578): (w + hyper) normalised."""
        if self.estimateComponentProbs:
            v = np.asarray(w, np.float64) + self.hyper
            self.weights = v / v.sum()

    def train(self, sample: PhyloSample, max_iterations=100, eps=1e-6, rng=None, verbose=False, sharded=False):
        """EM (jstacs AbstractMixtureModel.continueIterations :857-880 with trainOnlyMotifModel = true): E-step on the GPU,
        M-step of the motif model through PhyloBayesModel.train, component weights from the sufficient statistics.  Stops when
        the log-likelihood changes by less than eps (SmallDifferenceOfFunctionEvaluationsCondition) or after max_iterations.
        sharded=True: `sample` is this rank's shard (one process per GPU, torch.distributed initialised): [ll, w0, w1] of
        every E-step and every M-step objective value are all-reduced, so all ranks see the same numbers, leave the loop in
        the same iteration and set the same component weights.  Returns the list of log-likelihoods, one per E-step."""
        history, trained = [], False
        if self.strand is not None:
            self.strand.sharded = bool(sharded)
        for it in range(max_iterations):
            res = self.getNewWeights(sample, want_gamma=True)
            if sharded:
                from . import sharding
                v = sharding.allreduce_host_(np.array([res["ll"], res["w"][0], res["w"][1]]))
                res["ll"], res["w"] = float(v[0]), v[1:].copy()
            history.append(res["ll"])
            if verbose:
                print(f"{it}\t{res['ll']:.6f}\t{0.0 if it == 0 else res['ll'] - history[-2]:.3e}")
            if trained and abs(history[-1] - history[-2]) < eps:
                break
            if self.strand is None:
                st = self.motif.train(sample, None, rng=rng, sharded=sharded)   # gamma of this E-step is still resident
            elif it == 0:                                      # jstacs AbstractMixtureModel.java:1017-1033
                st = self.strand.doFirstIteration(sample, res["gamma"], rng)
            else:
                st, _ = self.strand.continueIterations(sample, res["gamma"], 1, rng)
            trained = not st["skipped"]                    # the motif ignores its first call (SKIP_TRAINING_STEPS = 1)
            self.setComponentWeights(res["w"])
        return history

    def getLogProbFor(self, sample: PhyloSample):
        """log-likelihood of the sample under the current parameters (sum over alignments)."""
        return self.getNewWeights(sample, want_gamma=False)["ll"]

    def getProfileOfScoresFor(self, sample: PhyloSample):
        """Unnormalised conditional profile a_j of every alignment (jstacs :643-688)."""
        return self.getNewWeights(sample, want_gamma=False, want_profile=True)["profile"]

    def getStrandProbabilitiesFor(self, sample: PhyloSample):
        res = self.getNewWeights(sample, want_gamma=False)
        return res["argmax_off"], res["argmax_strand"]


class MultiMotifMixture:
    """ZOOPS mixture of K motif models and a flanking model ("no motif"): BASELINE.json configs[3].  The reference's
    SingleHiddenMotifMixture has one motif component (getNumberOfMotifs() == 1); this is its generalisation with the same
    E-step arithmetic per component (include/phyfoo.h phyfoo_zoops_estep_multi, DESIGN.md "multi-motif mixture"): one site of
    at most one motif per alignment.  M-step: every motif is trained like the single one, with its own gamma_k as weights
    (PhyloBayesModel.train); component weights from the sufficient statistics w[K + 1] with hyper-parameters 1
    (jstacs AbstractMixtureModel.getNewComponentProbs :1542-1578)."""

    def __init__(self, motifs, flanking, weights=None, estimateComponentProbs=True):
        self.motifs, self.flanking = list(motifs), flanking
        K = len(self.motifs)
        self.weights = np.full(K + 1, 1.0 / (K + 1)) if weights is None else np.asarray(weights, np.float64)
        self.estimateComponentProbs = bool(estimateComponentProbs)
        self.hyper = np.ones(K + 1)
        self.ctx = self.motifs[0].ctx

    def getNumberOfMotifs(self):
        return len(self.motifs)

    def getNewWeights(self, sample: PhyloSample, seqWeights=None, want_gamma=True):
        with np.errstate(divide="ignore"):
            logw = np.log(self.weights)
        return core.zoops_estep_multi(self.ctx, sample.device(self.ctx), [m.device_model() for m in self.motifs],
                                      self.flanking.device_model(), logw, seqWeights, want_gamma)

    def train(self, sample: PhyloSample, max_iterations=20, eps=1e-6, rng=None, sharded=False):
        """EM; returns the log-likelihoods, one per E-step.  sharded: `sample` is this rank's shard (see
        SingleHiddenMotifMixture.train)."""
        rng = rng or np.random.default_rng(0)
        history, trained = [], False
        for it in range(max_iterations):
            res = self.getNewWeights(sample, want_gamma=False)
            ll, w = res["ll"], res["w"]
            if sharded:
                from . import sharding
                v = sharding.allreduce_host_(np.concatenate([[ll], w]))
                ll, w = float(v[0]), v[1:].copy()
            history.append(ll)
            if trained and abs(history[-1] - history[-2]) < eps:
                break
            ds = sample.device(self.ctx)
            for k, m in enumerate(self.motifs):
                m.trainingStep = getattr(m, "trainingStep", 0)
                if m.trainingStep < m.SKIP_TRAINING_STEPS:
                    m.trainingStep += 1
                    continue
                sa, so, sw = core.select_multi(self.ctx, ds, k, m.PROB_THRESH_FOR_WEIGHTS, m.MOTIF_QUALITY_THRESHOLD)
                m.train_selected(sample, sa, so, sw, rng=rng, sharded=sharded)
                trained = True
            if self.estimateComponentProbs:
                v = w + self.hyper
                self.weights = v / v.sum()
        return history


def _ab_sizes(n_tables, order):
    return [4 ** (min(p, order) + 1) for p in range(n_tables)]


def _ab_pack(tables, sizes):
    flat = np.concatenate([np.asarray(t, np.float64).ravel() for t in tables])
    if flat.size != sum(sizes):
        raise ValueError(f"expected tables of sizes {sizes}")
    return np.ascontiguousarray(flat)


def _ab_unpack(flat, sizes):
    out, o = [], 0
    for n in sizes:
        out.append(flat[o:o + n].copy())
        o += n
    return out


class AlignmentBasedModel:
    """models/AlignmentBasedModel.java: the non-phylogenetic motif model (`model.evolution=false`): per species row a Markov
    chain of the given order over the window's columns (order 0: one PWM applied to every row), a column whose context is
    unobserved scoring log(0.25), count-based M-step.  condProb[p] has 4^(min(p, order) + 1) entries indexed
    sum_k 4^k x(l - k) (the current symbol fastest, :127-147); orders 0..2."""

    PSEUDO_COUNT = 1e-8                  # :32
    SKIP_TRAINING_STEPS = 1              # :31
    PROB_THRESH_FOR_WEIGHTS = 0.8
    MOTIF_QUALITY_THRESHOLD = 0.0

    def __init__(self, length, order=0, ctx=None):
        if not 0 <= order <= 2:
            raise NotImplementedError("alignment-based models: orders 0..2")
        self.ctx = ctx or core.default_context()
        self.length, self.order = int(length), int(order)
        self.sizes = _ab_sizes(self.length, self.order)
        self.condProb = [np.full(n, 0.25) for n in self.sizes]
        self.trainingStep = 0

    def getLength(self):
        return self.length

    def getOrder(self):
        return self.order

    def setCondProbs(self, condProbs):
        self.condProb = _ab_unpack(_ab_pack(condProbs, self.sizes), self.sizes)

    def getCondProbs(self):
        return np.stack(self.condProb) if self.order == 0 else [c.copy() for c in self.condProb]

    def getLogProbFor(self, sample: PhyloSample, strand=0):
        ds = sample.device(self.ctx)
        out = np.zeros(ds.windows(self.length))
        cp = _ab_pack(self.condProb, self.sizes)
        if self.order == 0:
            rc = core.lib().phyfoo_ab_score_windows(self.ctx.h, ds.h, self.length, core._p(cp, core.C.c_double), int(strand),
                                                    core._p(out, core.C.c_double))
        else:
            rc = core.lib().phyfoo_ab_score_windows_order(self.ctx.h, ds.h, self.length, self.order, core._p(cp, core.C.c_double),
                                                          int(strand), core._p(out, core.C.c_double))
        core.raise_for(rc, self.ctx.h)
        return out

    def train(self, sample: PhyloSample, weights):
        if self.trainingStep < self.SKIP_TRAINING_STEPS:
            self.trainingStep += 1
            return {"skipped": True}
        ds = sample.device(self.ctx)
        sa, so, sw = core.select(self.ctx, ds, self.length, weights, self.PROB_THRESH_FOR_WEIGHTS, self.MOTIF_QUALITY_THRESHOLD)
        out = np.zeros(sum(self.sizes))
        if self.order == 0:
            rc = core.lib().phyfoo_ab_train(self.ctx.h, ds.h, self.length, sa.size, core._p(sa, core.C.c_int32),
                                            core._p(so, core.C.c_int32), None, core._p(sw, core.C.c_double),
                                            float(self.PSEUDO_COUNT), core._p(out, core.C.c_double))
        else:
            rc = core.lib().phyfoo_ab_train_order(self.ctx.h, ds.h, self.length, self.order, sa.size, core._p(sa, core.C.c_int32),
                                                  core._p(so, core.C.c_int32), None, core._p(sw, core.C.c_double),
                                                  float(self.PSEUDO_COUNT), core._p(out, core.C.c_double))
        core.raise_for(rc, self.ctx.h)
        self.condProb = _ab_unpack(out, self.sizes)
        self.trainingStep += 1
        return {"skipped": False, "selected": int(sa.size)}


class AlignmentBasedBGModel:
    """models/AlignmentBasedBGModel.java: the non-phylogenetic background, orders 0..2 (order + 1 tables; column l of an
    alignment uses table min(l, order) and looks back to the start of the alignment whatever range is asked for, so
    getLogProbFor(seq, s, e) is the sum of the per-column terms over [s, e])."""

    PSEUDO_COUNT = 1e-2                  # :31

    def __init__(self, order=0, ctx=None):
        if not 0 <= order <= 2:
            raise NotImplementedError("alignment-based models: orders 0..2")
        self.ctx = ctx or core.default_context()
        self.order = int(order)
        self.sizes = _ab_sizes(self.order + 1, self.order)
        self.condProb = [np.full(n, 0.25) for n in self.sizes]

    def setCondProbs(self, condProb):
        if self.order == 0:
            condProb = [np.asarray(condProb, np.float64).reshape(4)]
        self.condProb = _ab_unpack(_ab_pack(condProb, self.sizes), self.sizes)

    def getCondProbs(self):
        return self.condProb[0].copy() if self.order == 0 else [c.copy() for c in self.condProb]

    def getColumnLogProbs(self, sample: PhyloSample):
        ds = sample.device(self.ctx)
        out = np.zeros(ds.n_cols)
        cp = _ab_pack(self.condProb, self.sizes)
        if self.order == 0:
            rc = core.lib().phyfoo_ab_bg_columns(self.ctx.h, ds.h, core._p(cp, core.C.c_double), core._p(out, core.C.c_double))
        else:
            rc = core.lib().phyfoo_ab_bg_columns_order(self.ctx.h, ds.h, self.order, core._p(cp, core.C.c_double), core._p(out, core.C.c_double))
        core.raise_for(rc, self.ctx.h)
        return out

    def getLogProbFor(self, sample: PhyloSample):
        cols = self.getColumnLogProbs(sample)
        off = sample.col_offsets
        return np.array([cols[off[i]:off[i + 1]].sum() for i in range(sample.n_alignments)])

    def train(self, sample: PhyloSample, weights=None):
        """:40-137: counts over every column of every alignment, weighted per alignment."""
        ds = sample.device(self.ctx)
        out = np.zeros(sum(self.sizes))
        w = None if weights is None else np.ascontiguousarray(weights, np.float64)
        if w is not None and w.size != sample.n_alignments:
            raise ValueError("one weight per alignment")
        core.raise_for(core.lib().phyfoo_ab_bg_train_order(self.ctx.h, ds.h, self.order, core._p(w, core.C.c_double),
                                                           float(self.PSEUDO_COUNT), core._p(out, core.C.c_double)), self.ctx.h)
        self.condProb = _ab_unpack(out, self.sizes)
