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
