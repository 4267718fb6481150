"""Host-side consumer of the fixed-q objective: a small quasi-Newton (BFGS) minimiser standing in for Jstacs'
`Optimizer.optimize(QUASI_NEWTON_BFGS, ...)` (jstacs de/jstacs/algorithms/optimization/Optimizer.java:1000-1080), which stays
on the host in the reference too.  Function values and forward-difference gradients come from
the GPU (`core.FESet.eval` / `.grad`); everything here is O(dim^2) numpy.

Not parity-checked: the reference's optimisation trajectory is not reproducible run to run (random position order and
start parameters); what is mirrored is the termination logic of optimizing/PreparedConditions.java.
"""
import numpy as np


class CombinedCondition:
    """jstacs CombinedCondition(threshold, ...): keep iterating while at least `threshold` of the sub-conditions say so."""

    def __init__(self, threshold, small_difference=None, iterations=None, small_gradient=None, small_step=None):
        self.threshold = threshold
        self.small_difference, self.iterations = small_difference, iterations
        self.small_gradient, self.small_step = small_gradient, small_step

    def do_next_iteration(self, it, f_last, f_cur, grad, step):
        votes = []
        if self.small_difference is not None:
            votes.append(abs(f_last - f_cur) >= self.small_difference)
        if self.iterations is not None:
            votes.append(it < self.iterations)
        if self.small_gradient is not None:
            votes.append(np.abs(grad).sum() >= self.small_gradient)
        if self.small_step is not None:
            votes.append(step >= self.small_step)
        return sum(votes) >= self.threshold


# optimizing/PreparedConditions.java:20-35
def TC_MOTIF_SEPARATED_POSITIONS():
    return CombinedCondition(3, 1e-4, 4, 1e-4, 1e-4)


def TC_MOTIF_ALL_POSITIONS():
    return CombinedCondition(2, 1e-5, None, 1e-4, 1e-4)


def SMALL_DIFFERENCE_OF_FUNCTIONS():
    return CombinedCondition(1, 1e-3)


def _line_search(fun, x, f0, d, start, lin_eps):
    """Bracket by golden-ratio expansion from `start`, then golden-section to lin_eps (Optimizer.java:122-160,202-330 use the same
    bracket + Brent scheme; every evaluation is one GPU pass)."""
    gr = 1.618033988749895
    a, fa = 0.0, f0
    b = start
    fb = fun(x + b * d)
    if fb < fa:
        c, fc = b * (1 + gr), fun(x + b * (1 + gr) * d)
        while fc < fb and c < 1e6:
            a, fa, b, fb = b, fb, c, fc
            c = b * (1 + gr)
            fc = fun(x + c * d)
    else:
        c, fc = b, fb
        b = c / (1 + gr)
        fb = fun(x + b * d)
        while fb >= fa and b > 1e-12:
            c, fc = b, fb
            b = c / (1 + gr)
            fb = fun(x + b * d)
        if fb >= fa:
            return 0.0, f0
    # a < b < c with fb the smallest
    while (c - a) > lin_eps * max(1.0, abs(b)):
        if (c - b) > (b - a):
            t = b + (c - b) / (1 + gr)
            ft = fun(x + t * d)
            if ft < fb:
                a, fa, b, fb = b, fb, t, ft
            else:
                c, fc = t, ft
        else:
            t = b - (b - a) / (1 + gr)
            ft = fun(x + t * d)
            if ft < fb:
                c, fc, b, fb = b, fb, t, ft
            else:
                a, fa = t, ft
    return b, fb


def quasi_newton_bfgs(fun, grad, x0, condition, lin_eps=1e-3, start_distance=0.1):
    """Minimises fun from x0.  grad(x) -> (f(x), g(x)).  Returns (x, f, iterations, evaluations)."""
    x = np.array(x0, np.float64)
    n = x.size
    H = np.eye(n)
    f, g = grad(x)
    evals = n + 1
    it, steps = 0, []
    while True:
        d = -H @ g
        if g @ d >= 0:
            H = np.eye(n)
            d = -g
        norm = np.linalg.norm(d)
        if norm == 0:
            break
        start = (0.667 * np.median(steps[-10:]) if steps else start_distance) / norm     # LimitedMedianStartDistance(10, .1)
        counter = [0]

        def f1(z):
            counter[0] += 1
            return fun(z)

        alpha, f_new = _line_search(f1, x, f, d, max(start, 1e-8), lin_eps)
        evals += counter[0]
        s = alpha * d
        x_new = x + s
        f_new2, g_new = grad(x_new)
        evals += n + 1
        y = g_new - g
        sy = s @ y
        if sy > 1e-14:
            rho = 1.0 / sy
            V = np.eye(n) - rho * np.outer(s, y)
            H = V @ H @ V.T + rho * np.outer(s, s)
        step = float(np.linalg.norm(s))
        steps.append(step)
        it += 1
        go_on = condition.do_next_iteration(it, f, f_new2, g_new, step)
        x, f, g = x_new, f_new2, g_new
        if not go_on or alpha == 0.0:
            break
    return x, f, it, evals


class SuffStatObjective:
    """The fixed-q objective in sufficient-statistics form (SURVEY.md section 7 hard part 3): after ONE device pass
    (core.FESet.suffstats) every evaluation is a dot product on the host,
        f(theta) = sum_t sum_e C[t][e] * log table_theta,t[e] - ENT,
    for the stationary distributions (FE_byParameter_ForBayesNodeJstacs / MotifParamOptimizer) and for the branch
    lengths (optimizing/branchlengths/EdgeOptimizer.java, GlobalEdgeOptimizer.java) alike.  Order-0 models."""

    def __init__(self, fe, pi, branch, evol):
        from . import evolution
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
            lt = np.log(self._ev.tables(pi, branch, self.evol)[self._nz[t]])
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
        fun(x)
        return f0, g
