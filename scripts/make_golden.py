#!/usr/bin/env python
"""Generates tests/golden/c1_tree1_first8.npz: BASELINE config 1 in miniature.

Inputs: the first 8 alignments of the reference's bundled data/synthetic_data/trees/data_tree1_1.0.fg (read from
/root/reference, which exists only in the build container -- hence the committed fixture) and the generating
PWM / tree of data/synthetic_data/trees/description_tree1.txt.  Outputs: what the ORACLE (oracle/phyfoo_oracle.cpp)
computes for them.  The reference itself cannot run here (no JVM), so these vectors pin the oracle against drift and
give the GPU tests a fixed target; they are not reference outputs.

    python scripts/make_golden.py
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as orc                    # noqa: E402
from phyfoo_b200.data import PhyloSample, read_fasta  # noqa: E402

FASTA = "/root/reference/data/synthetic_data/trees/data_tree1_1.0.fg"
CAT = "((((SPECIES_0:0.2,SPECIES_1:0.2):0.2,SPECIES_2:0.2):0.2,SPECIES_3:0.2):0.2,SPECIES_4:0.2)"
STAR = "(SPECIES_0:0.2,SPECIES_1:0.2,SPECIES_2:0.2,SPECIES_3:0.2,SPECIES_4:0.2)"
# description_tree1.txt "Motif alignments" (rows sum to 1 within rounding; renormalised below)
PWM = np.array([
    [0.1993, 0.3155, 0.4603, 0.0249], [0.2817, 0.0617, 0.1859, 0.4707], [0.1483, 0.061, 0.4091, 0.3817],
    [0.1673, 0.5331, 0.1219, 0.1777], [0.5726, 0.0482, 0.1283, 0.251], [0.6365, 0.0872, 0.0963, 0.18],
    [0.8143, 0.0465, 0.1381, 0.001], [0.6613, 0.2003, 0.0827, 0.0557], [0.639, 0.0866, 0.2522, 0.0223],
    [0.4891, 0.0454, 0.1373, 0.3281]])
N_ALN, W = 8, 10


def main():
    pwm = PWM / PWM.sum(axis=1, keepdims=True)
    smp = PhyloSample.from_fasta(FASTA, CAT, limit=N_ALN)
    records = read_fasta(FASTA)[: N_ALN * 5]
    out = {
        "fasta_headers": np.array([h for h, _ in records]), "fasta_seqs": np.array([s for _, s in records]),
        "codes": smp.codes, "col_offsets": smp.col_offsets, "pwm": pwm, "newick_cat": np.array(CAT), "newick_star": np.array(STAR),
        "motif_pos": np.array([int(a["MOTIF_POS"]) for a in smp.annotations]),
        "oracle_git": np.array(subprocess.run(["git", "-C", ROOT, "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()),
    }
    rng = np.random.default_rng(0x5EED0001)
    gapped = smp.codes.copy()
    g = rng.random(gapped.shape) < 0.05
    g[g.all(axis=1), 0] = False
    gapped[g] = 4
    out["codes_gapped"] = gapped
    bases, gaps = orc.pack_columns(gapped)
    out["packed_bases_gapped"], out["packed_gaps_gapped"] = bases, gaps
    for name, newick in (("cat", CAT), ("star", STAR)):
        for variant, codes in (("", smp.codes), ("_gapped", gapped)):
            fg = orc.Model(orc.FG, W, 0, newick)
            for t in range(W):
                fg.set_pi(t, pwm[t])
            bg = orc.Model(orc.BG, 0, 0, newick)
            sc, sw = [], []
            for i in range(N_ALN):
                a, b = smp.col_offsets[i], smp.col_offsets[i + 1]
                s, k = fg.score_windows(codes[a:b])
                sc.append(s); sw.append(k)
            key = f"{name}{variant}"
            out[f"mf_scores_{key}"], out[f"mf_sweeps_{key}"] = np.concatenate(sc), np.concatenate(sw)
            fg.set_options(mode=orc.EXACT_PRUNING)
            out[f"exact_scores_{key}"] = np.concatenate(
                [fg.score_windows(codes[smp.col_offsets[i]:smp.col_offsets[i + 1]])[0] for i in range(N_ALN)])
            fg.set_options(mode=orc.MEANFIELD)
            out[f"bg_columns_{key}"] = bg.bg_columns(codes)
            for strand in (False, True):
                res = orc.estep(fg, bg, smp.col_offsets, codes, np.log([0.9, 0.1]), np.log([0.5, 0.5]) if strand else None)
                sk = f"{key}{'_strand' if strand else ''}"
                out[f"estep_gamma_{sk}"] = res["gamma"]
                out[f"estep_llw_{sk}"] = np.array([res["ll"], res["w"][0], res["w"][1]])
                out[f"estep_argmax_off_{sk}"] = res["argmax_off"]
                out[f"estep_argmax_strand_{sk}"] = res["argmax_strand"]
            # M-step: selection from the single-strand gamma, fixed-q objective and forward-difference gradient at position 3
            gamma = out[f"estep_gamma_{key}"]
            woff = np.concatenate([[0], np.cumsum(smp.lengths() - W + 1)])
            sa, so, swt = [], [], []
            for i in range(N_ALN):
                idx, ww = orc.select_group(gamma[woff[i]:woff[i + 1]], 0.8, 0.0)
                sa += [i] * len(idx); so += list(idx); swt += list(ww)
            cols = np.stack([codes[smp.col_offsets[a] + o: smp.col_offsets[a] + o + W] for a, o in zip(sa, so)])
            fe = orc.FESet(fg, cols, np.array(swt))
            lam = np.log(pwm[3]) + 0.1 * np.array([1.0, -1.0, 0.5, 0.0])
            f0, grad = fe.grad(3, lam, 1e-4)
            fe.eval(3, np.log(pwm[3]))
            out[f"sel_aln_{key}"], out[f"sel_off_{key}"], out[f"sel_w_{key}"] = np.array(sa, np.int32), np.array(so, np.int32), np.array(swt)
            out[f"fe_lambda_{key}"], out[f"fe_f0_{key}"], out[f"fe_grad_{key}"] = lam, np.array(f0), grad
    path = os.path.join(ROOT, "tests", "golden", "c1_tree1_first8.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)} bytes, {len(out)} arrays; planted motif found by arg-max in "
          f"{int((out['estep_argmax_off_cat'] == out['motif_pos']).sum())}/{N_ALN} alignments")


if __name__ == "__main__":
    main()
