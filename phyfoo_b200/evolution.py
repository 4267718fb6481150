"""Per-branch substitution tables on the host (numpy mirror of evolution/FS81alpha.java:76-92 and FS81beta.java:81-103, the
same formulas csrc/model_host.hpp uses), in the entry layout of the device tables: 4 root entries (pi), then 16 per
non-root node k at 4 + 16 (k - 1), row-major [parent symbol a][own symbol b].

Used by the sufficient-statistics form of the M-step objective (SuffStatObjective): with the expected counts C from ONE
device pass, f(theta) = sum_t sum_e C[t][e] log table_theta,t[e] - ENT is a host-side dot product for any pi or branch
lengths."""
import math

import numpy as np

from . import _abi


def hky_transition(pi, alpha, beta):
    """evolution/HKY.java:77-106 as written: transitions a<->g, c<->t at rate alpha, transversions at rate beta."""
    a, b = alpha * np.asarray(pi, np.float64), beta * np.asarray(pi, np.float64)
    return np.array([[1 - (a[2] + b[3] + b[1]), b[1], a[2], b[3]],
                     [b[0], 1 - (a[3] + b[0] + b[2]), b[2], a[3]],
                     [a[0], b[1], 1 - (a[0] + b[3] + b[1]), b[3]],
                     [b[0], a[1], b[2], 1 - (a[1] + b[0] + b[2])]])


def transition(pi, length, evol=_abi.EVOL_F81ALPHA, hky_ratio=0.0):
    """4x4 P(b | a) for one branch."""
    pi = np.asarray(pi, np.float64)
    if evol == _abi.EVOL_HKY_RATIO:
        return hky_transition(pi, length, length * hky_ratio)
    if evol == _abi.EVOL_F81ALPHA:
        keep = 1.0 - length
        mix = length
    elif evol == _abi.EVOL_F81BETA:
        beta = 1.0 / (1.0 - pi[0] * pi[0] - pi[1] * pi[1] - pi[2] * pi[2] - pi[3] * pi[3])       # the reference's order (FS81beta.java:86)
        ta = math.exp(-length * beta)
        if ta < 1e-4:
            ta = 0.0
        elif ta > 1 - 1e-4:
            ta = 1.0
        keep, mix = ta, 1.0 - ta
    else:
        raise ValueError("HKY as implemented has zero transversion probabilities "
                         "(HKY.java:32) and no finite objective")
    T = np.tile(pi * mix, (4, 1))
    T[np.arange(4), np.arange(4)] = keep + pi * mix
    return T


def tables(pi, branch, evol=_abi.EVOL_F81ALPHA, hky_ratio=0.0):
    """Probability table of one position: entries [pi(4), node 1 (16), ..., node N-1 (16)]; branch[N] (entry 0 unused).
    All nodes at once (this sits in the inner loop of the sufficient-statistics objective); entry for entry the values of
    transition()."""
    pi = np.asarray(pi, np.float64)
    b = np.asarray(branch, np.float64)[1:]
    if evol == _abi.EVOL_HKY_RATIO:
        return np.concatenate([pi] + [hky_transition(pi, x, x * hky_ratio).ravel() for x in b])
    if evol == _abi.EVOL_F81ALPHA:
        keep, mix = 1.0 - b, b
    elif evol == _abi.EVOL_F81BETA:
        beta = 1.0 / (1.0 - pi[0] * pi[0] - pi[1] * pi[1] - pi[2] * pi[2] - pi[3] * pi[3])       # the reference's order (FS81beta.java:86)
        ta = np.array([math.exp(-x * beta) for x in b])                # libm's exp like csrc/model_host.hpp (numpy's differs by an ulp)
        ta = np.where(ta < 1e-4, 0.0, np.where(ta > 1 - 1e-4, 1.0, ta))
        keep, mix = ta, 1.0 - ta
    else:
        raise ValueError("HKY as implemented has zero transversion probabilities "
                         "(HKY.java:32) and no finite objective")
    off = pi[None, :] * mix[:, None]                                  # [N-1][4]: P(b | a != b)
    T = np.repeat(off[:, None, :], 4, axis=1)                         # [N-1][4][4]
    d = np.arange(4)
    T[:, d, d] = keep[:, None] + off
    out = np.empty(4 + 16 * b.size)
    out[:4] = pi
    out[4:] = T.ravel()
    return out


def branch_from_logit(x, minimum=0.0005):
    """optimizing/branchlengths/EdgeOptimizer.java:44-51: alpha = (e^x + 0.0005) / (1 + e^x), NaN -> 1."""
    with np.errstate(over="ignore", invalid="ignore"):
        e = np.exp(np.asarray(x, np.float64))
        a = (e + minimum) / (1.0 + e)
    return np.where(np.isnan(a), 1.0, a)


def logit_from_branch(alpha):
    """EdgeOptimizer.java:103-109: x = log(alpha / (1 - alpha)), alpha == 1 is nudged to 1 - 1e-6."""
    a = np.asarray(alpha, np.float64).copy()
    a[a == 1.0] -= 1e-6
    return np.log(a / (1.0 - a))
