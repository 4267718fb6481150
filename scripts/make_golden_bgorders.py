#!/usr/bin/env python
"""Generates tests/golden/c1_tree1_first8_bgorders.npz: the first alignments of the bundled data (taken from the committed
tests/golden/c1_tree1_first8.npz, so this script needs nothing outside the repo) under PhyloBackground models of orders 1 and 2
(models/PhyloBackground.java:241-305, 381-398): whole-alignment log-probabilities and the ZOOPS E-step with such a background,
as the ORACLE computes them.  Like the other golden vectors they pin the oracle against drift and give the GPU tests a fixed
target; they are not reference outputs (no JVM).

    python scripts/make_golden_bgorders.py
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as orc                    # noqa: E402

N_ALN, N_ESTEP, W = 8, 3, 10


def bg_pis(order):
    rng = np.random.default_rng(0x5EED0100 + order)
    return [rng.dirichlet(np.full(4, 8.0), size=4 ** min(order, t)) for t in range(order + 1)]


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "c1_tree1_first8.npz"))
    codes, off, pwm, newick = g["codes_gapped"], g["col_offsets"], g["pwm"], str(g["newick_cat"])
    out = {"oracle_git": np.array(subprocess.run(["git", "-C", ROOT, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip())}
    fg = orc.Model(orc.FG, W, 0, newick)
    for t in range(W):
        fg.set_pi(t, pwm[t])
    for order in (1, 2):
        pis = bg_pis(order)
        for t, p in enumerate(pis):
            out[f"pi_o{order}_t{t}"] = p
        bg = orc.Model(orc.BG, 0, order, newick)
        for t, p in enumerate(pis):
            bg.set_pi(t, p.ravel())
        out[f"logprob_o{order}"] = np.array([bg.bg_range(codes[off[i]:off[i + 1]], 0, int(off[i + 1] - off[i]) - 1) for i in range(N_ALN)])
        out[f"range_o{order}"] = np.array([bg.bg_range(codes[off[i]:off[i + 1]], 7, 7 + order + 11) for i in range(N_ALN)])
        sub_off = off[: N_ESTEP + 1]
        res = orc.estep(fg, bg, sub_off, codes[: sub_off[-1]], np.log([0.9, 0.1]))
        out[f"estep_gamma_o{order}"] = res["gamma"]
        out[f"estep_llw_o{order}"] = np.array([res["ll"], res["w"][0], res["w"][1]])
        out[f"estep_argmax_off_o{order}"] = res["argmax_off"]
    path = os.path.join(ROOT, "tests", "golden", "c1_tree1_first8_bgorders.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
