"""Host-side consumer of the fixed-q objective: Jstacs' `Optimizer.optimize(QUASI_NEWTON_BFGS, ...)` restated statement by
statement (jstacs de/jstacs/algorithms/optimization/Optimizer.java: findBracket :122-148, brentsMethod :202-303,
quasiNewtonBFGS :1000-1080; OneDimensionalFunction.findMin; LimitedMedianStartDistance; the termination conditions of
de/jstacs/algorithms/optimization/termination/).  The optimizer stays on the host in the reference too (SURVEY.md section 8 row
a16); function values and forward-difference gradients come from the GPU (`core.FESet.eval` / `.grad`) or from the library's
sufficient-statistics objective (`core.SSObjective`); everything here is O(dim^2) per iteration.  The same statements in C++:
`csrc/suffstat_host.hpp::bfgs_minimise` (what `phyfoo_ss_optimize` runs); tests/test_optimize_host.py holds the two against each
other.

Parity status: restated from the Java source, not run against it (no JVM).  Given the same start vector and the same function
values the iterates follow the reference's; whole training runs of the reference are still not reproducible because it draws
start parameters and the position order from unseeded generators.
"""
import math

import numpy as np

LIM_FIB = (math.sqrt(5.0) - 1.0) / 2.0          # Optimizer.java:66-70
LIM_FIB_PLUS1 = LIM_FIB + 1.0
LIM_FIB_PLUS2 = LIM_FIB + 2.0
_GUARD = 100000                                 # the Java loops have no bound; a function unbounded below would never return


class CombinedCondition:
    """jstacs CombinedCondition(threshold, ...): go on while at least `threshold` of the sub-conditions say so
    (termination/CombinedCondition.java).  Sub-conditions: SmallDifferenceOfFunctionEvaluationsCondition (eps <= |f_last - f|),
    IterationCondition (iteration < max), SmallGradientConditon (sum |g_i| >= eps), SmallStepCondition (|d|^2 alpha^2 >= eps --
    the SQUARED length of the step)."""

    def __init__(self, threshold, small_difference=None, iterations=None, small_gradient=None, small_step=None):
        self.threshold = threshold
        self.small_difference, self.iterations = small_difference, iterations
        self.small_gradient, self.small_step = small_gradient, small_step

    def do_next_iteration(self, it, f_last, f_cur, grad, direction, alpha):
        votes = 0
        if self.small_difference is not None:
            votes += bool(self.small_difference <= abs(f_last - f_cur))
        if self.iterations is not None:
            votes += bool(it < self.iterations)
        if self.small_gradient is not None:
            votes += bool(float(np.abs(grad).sum()) >= self.small_gradient)
        if self.small_step is not None:
            with np.errstate(invalid="ignore"):
                votes += bool(float(np.dot(direction, direction)) * (alpha * alpha) >= self.small_step)
        return votes >= self.threshold


# optimizing/PreparedConditions.java:20-35
def TC_MOTIF_SEPARATED_POSITIONS():
    return CombinedCondition(3, 1e-4, 4, 1e-4, 1e-4)


def TC_MOTIF_ALL_POSITIONS():
    return CombinedCondition(2, 1e-5, None, 1e-4, 1e-4)


def SMALL_DIFFERENCE_OF_FUNCTIONS():
    return CombinedCondition(1, 1e-3)


class LimitedMedianStartDistance:
    """LimitedMedianStartDistance(slots, value): 0.667 x the median of the last `slots` line-search results, the memory
    initially filled with `value`."""

    def __init__(self, slots=10, value=0.1):
        self.memory = [float(value)] * int(slots)
        self.index = 0

    def get_new_start_distance(self):
        srt = sorted(self.memory)
        n = len(srt)
        median = srt[n // 2] if n % 2 == 1 else 0.5 * (srt[n // 2 - 1] + srt[n // 2])
        return 0.667 * median

    def set_last_distance(self, last):
        self.memory[self.index] = float(last)
        self.index = (self.index + 1) % len(self.memory)


def find_bracket(f, lower, f_lower, start_distance):
    """Optimizer.findBracket (:122-148): [x_lower, f, x_middle, f, x_upper, f]."""
    erg = [lower, f_lower, lower + start_distance, 0.0, 0.0, 0.0]
    erg[3] = f(erg[2])
    if erg[1] < erg[3]:
        for _ in range(_GUARD):
            erg[4], erg[5] = erg[2], erg[3]
            erg[2] = erg[0] + (1 - LIM_FIB) * (erg[2] - erg[0])
            erg[3] = f(erg[2])
            if not erg[1] < erg[3]:
                return erg
    else:
        for _ in range(_GUARD):
            erg[4] = LIM_FIB_PLUS2 * erg[2] - LIM_FIB_PLUS1 * erg[0]
            erg[5] = f(erg[4])
            if erg[3] >= erg[5]:
                erg[0], erg[1], erg[2], erg[3] = erg[2], erg[3], erg[4], erg[5]
            if not erg[3] == erg[5]:
                return erg
    raise ArithmeticError("findBracket does not terminate on this function (the reference would loop forever)")


def brents_method(f, a, x, fx, b, tol):
    """Optimizer.brentsMethod (:202-303): (x_min, f(x_min)) in [a, b] from an inner point x."""
    c = 1 - LIM_FIB
    d = 0.0
    eps = math.sqrt(1.2e-16)
    e = 0.0
    v = w = x
    fv = fw = fx
    tol3 = tol / 3.0
    xm = .5 * (a + b)
    tol1 = eps * abs(x) + tol3
    t2 = 2.0 * tol1
    for _ in range(_GUARD):
        if not abs(x - xm) > (t2 - .5 * (b - a)):
            return x, fx
        p = q = r = 0.0
        if abs(e) > tol1:                       # fit the parabola
            r = (x - w) * (fx - fv)
            q = (x - v) * (fx - fw)
            p = (x - v) * q - (x - w) * r
            q = 2.0 * (q - r)
            if q > 0.0:
                p = -p
            else:
                q = -q
            r = e
            e = d
        if abs(p) < abs(.5 * q * r) and p > q * (a - x) and p < q * (b - x):
            d = p / q                           # a parabolic interpolation step
            u = x + d
            if (u - a) < t2 or (b - u) < t2:    # f must not be evaluated too close to a or b
                d = tol1
                if x >= xm:
                    d = -d
        else:                                   # a golden-section step
            e = b - x if x < xm else a - x
            d = c * e
        if abs(d) >= tol1:                      # f must not be evaluated too close to x
            u = x + d
        else:
            u = x + tol1 if d > 0.0 else x - tol1
        fu = f(u)
        if fx <= fu:
            if u < x:
                a = u
            else:
                b = u
        if fu <= fx:
            if u < x:
                b = x
            else:
                a = x
            v, fv, w, fw, x, fx = w, fw, x, fx, u, fu
        elif fu <= fw or w == x:
            v, fv, w, fw = w, fw, u, fu
        elif not (fu > fv and v != x and v != w):
            v, fv = u, fu
        xm = .5 * (a + b)
        tol1 = eps * abs(x) + tol3
        t2 = 2.0 * tol1
    raise ArithmeticError("brentsMethod does not terminate on this function")


def quasi_newton_bfgs(fun, grad, x0, condition, lin_eps=1e-3, start_distance=0.1):
    """Optimizer.quasiNewtonBFGS (:1000-1080) minimising fun from x0.  grad(x) -> (f(x), g(x)) (only g is used: like the Java,
    the function value of an iterate is the one its line search found).  start_distance: the initial value of the
    LimitedMedianStartDistance(10, .) the reference passes.  Returns (x, f, iterations, function evaluations)."""
    x = np.array(x0, np.float64)
    n = x.size
    evals = [0]

    def f_at(z):
        evals[0] += 1
        return float(fun(z))

    def g_at(z):
        evals[0] += n + 1                       # NumericalDifferentiableFunction.java:64-84: f(x) and n perturbed points
        return np.array(grad(z)[1], np.float64)

    sdf = LimitedMedianStartDistance(10, start_distance)
    help_ = [math.inf, f_at(x)]
    g_new = g_at(x)
    d = -g_new
    matrix = np.eye(n)
    i = 0
    nxt = condition.do_next_iteration(i, math.inf, help_[1], g_new, d, help_[0])
    while nxt:
        current = help_[1]
        sd = sdf.get_new_start_distance()
        if not sd > 0:                          # Optimizer.checkStartDistance
            raise ValueError("The start distance has to be positive.")
        x_cur, d_cur = x.copy(), d.copy()
        along = lambda a: f_at(x_cur + a * d_cur)                                    # noqa: E731  OneDimensionalSubFunction
        br = find_bracket(along, 0.0, current, sd)
        help_ = list(brents_method(along, br[0], br[2], br[3], br[4], lin_eps))       # OneDimensionalFunction.findMin
        i += 1
        sdf.set_last_distance(help_[0])
        x = x + d * help_[0]
        g_old = g_new
        g_new = g_at(x)
        nxt = condition.do_next_iteration(i, current, help_[1], g_new, d, help_[0])
        if nxt:
            s = d * help_[0]
            v = g_new - g_old
            sv = vv = vmatrixv = 0.0
            matrixv = np.zeros(n)
            for c1 in range(n):
                sv += s[c1] * v[c1]
                vv += v[c1] * v[c1]
                acc = 0.0
                for c2 in range(n):
                    acc += matrix[c1, c2] * v[c2]
                matrixv[c1] = acc
                vmatrixv += v[c1] * matrixv[c1]
            with np.errstate(divide="ignore", invalid="ignore"):
                u = s / sv - matrixv / vmatrixv
                for c1 in range(n):
                    for c2 in range(n):
                        matrix[c1, c2] += s[c1] * s[c2] / sv - matrixv[c1] * matrixv[c2] / vmatrixv + vmatrixv * u[c1] * u[c2]
                    acc = 0.0
                    for c2 in range(n):
                        acc += -matrix[c1, c2] * g_new[c2]
                    d[c1] = acc
    return x, help_[1], i, evals[0]


class SuffStatObjective:
    """The fixed-q objective in sufficient-statistics form (SURVEY.md section 7 hard part 3): after ONE device pass
    (core.FESet.suffstats) every evaluation is a dot product on the host,
        f(theta) = sum_t sum_e C[t][e] * log table_theta,t[e] - ENT,
    for the stationary distributions (FE_byParameter_ForBayesNodeJstacs / MotifParamOptimizer) and for the branch
    lengths (optimizing/branchlengths/EdgeOptimizer.java, GlobalEdgeOptimizer.java) alike.  Order-0 models."""

    def __init__(self, fe, pi, branch, evol, hky_ratio=0.0):
        from . import evolution
        self.hky_ratio = float(hky_ratio)
        self._ev = evolution
        self.pi = [np.asarray(p, np.float64).ravel().copy() for p in pi]
        self.branch = np.array(branch, np.float64)                   # [W][N]
        self.evol = evol
        self.W, self.N = self.branch.shape
        self.E = 4 + 16 * (self.N - 1)
        self._counts, self.entropy = fe.suffstats(self.W, self.E)
        self._index()
        self._G = np.array([self.position_term(t, self.pi[t], self.branch[t]) for t in range(self.W)])

    @property
    def counts(self):
        return self._counts

    @counts.setter
    def counts(self, c):
        self._counts = c
        self._index()

    def _index(self):
        # entries with a zero count never enter the sum (their log table may be -inf)
        self._nz = [np.nonzero(self._counts[t] != 0.0)[0] for t in range(self.W)]
        self._cnz = [self._counts[t][self._nz[t]] for t in range(self.W)]

    def position_term(self, t, pi, branch):
        with np.errstate(divide="ignore"):
            lt = np.log(self._ev.tables(pi, branch, self.evol, self.hky_ratio)[self._nz[t]])
        return float(np.dot(self._cnz[t], lt))

    def value(self):
        return float(self._G.sum() - self.entropy)

    def set_position(self, t, pi=None, branch=None):
        if pi is not None:
            self.pi[t] = np.asarray(pi, np.float64).ravel().copy()
        if branch is not None:
            self.branch[t] = branch
        self._G[t] = self.position_term(t, self.pi[t], self.branch[t])

    # ---- evaluateFunction / evaluateGradientOfFunction in the reference's parametrisations ----
    def eval_lambda(self, t, lam):
        """FE_byParameter_ForBayesNodeJstacs.evaluateFunction: pi_t = softmax(lambda); the model is left at lambda."""
        lam = np.asarray(lam, np.float64)
        e = np.exp(lam - lam.max())
        s = e.sum()
        self.set_position(t, pi=e / s if s != 1.0 else e)
        return self.value()

    def eval_edges(self, t, x, equal=False, globally=False):
        """EdgeOptimizer.evaluateFunction: branch lengths of position t (or of all positions) from logits x."""
        alpha = self._ev.branch_from_logit(x)
        b = np.zeros(self.N)
        b[1:] = alpha[0] if equal else alpha
        for pos in (range(self.W) if globally else [t]):
            self.set_position(pos, branch=b)
        return self.value()

    def grad(self, fun, x, eps=1e-4):
        """jstacs NumericalDifferentiableFunction.java:64-84: forward differences, x[i] perturbed in place and restored."""
        x = np.array(x, np.float64)
        f0 = fun(x)
        g = np.zeros_like(x)
        for i in range(x.size):
            keep = x[i]
            x[i] += eps
            g[i] = (fun(x) - f0) / eps
            x[i] = keep
        return f0, g                             # (no closing evaluation: the object stays at the last perturbed point, like the Java)
