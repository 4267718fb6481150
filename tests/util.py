"""Shared helpers for the tests: build the SAME model in the oracle and in the product."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402  (test infrastructure)


def oracle_fg(newick, pwm, order=0, mode=orc.MEANFIELD, evol=orc.F81ALPHA, hky_ratio=0.0):
    m = orc.Model(orc.FG, len(pwm), order, newick, evol, hky_ratio)
    for t, p in enumerate(pwm):
        m.set_pi(t, np.asarray(p, np.float64).ravel())
    m.set_options(mode=mode)
    return m


def oracle_bg(newick, order=0, pi=None, mode=orc.MEANFIELD, evol=orc.F81ALPHA, hky_ratio=0.0):
    m = orc.Model(orc.BG, 0, order, newick, evol, hky_ratio)
    if pi is not None:
        for t, p in enumerate(pi):
            m.set_pi(t, np.asarray(p, np.float64).ravel())
    m.set_options(mode=mode)
    return m


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def random_codes(rng, L, S, gap_rate=0.0):
    c = rng.integers(0, 4, size=(L, S)).astype(np.uint8)
    if gap_rate > 0:
        g = rng.random((L, S)) < gap_rate
        g[g.all(axis=1), 0] = False
        c[g] = 4
    return c
