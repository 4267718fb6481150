#!/usr/bin/env python
"""Generates tests/golden/generator_pin.npz: column-pattern histograms of the reference's bundled synthetic data
(/root/reference/data/synthetic_data, written by the reference's own sampler, bayesNet/BayesNetHandler.java:114-142 through
PhyloBackground / PhyloBayesModel.emitSample) together with the generating parameters their description*.txt files state.
The data are the only outputs of reference code that exist in this image (no JVM), so the histograms are what the oracle's
exact column likelihoods can be held against (tests/test_oracle_pins.py::test_*generator*): Newick parsing, species order, the
F81alpha meaning of a branch length and the pruning likelihood, to the statistical resolution of ~260 000 columns per file.

A pattern index is sum_s code[s] * 4**s over the five species in Newick order (codes a, c, g, t = 0..3).

    python scripts/make_generator_pin_fixture.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phyfoo_b200.data import PhyloSample            # noqa: E402

DATA = "/root/reference/data/synthetic_data/"
SP = [f"SPECIES_{i}" for i in range(5)]
TREE1 = "((((SPECIES_0:0.2,SPECIES_1:0.2):0.2,SPECIES_2:0.2):0.2,SPECIES_3:0.2):0.2,SPECIES_4:0.2)"   # trees/description_tree1.txt:6
TREE2 = "(((SPECIES_0:0.2,SPECIES_1:0.2):0.2,SPECIES_2:0.2):0.2,(SPECIES_3:0.2,SPECIES_4:0.2):0.2)"   # trees/description_tree2.txt
TREE3 = "((SPECIES_0:0.2,SPECIES_1:0.2):0.2,(SPECIES_2:0.2,SPECIES_3:0.2,SPECIES_4:0.2):0.2)"         # trees/description_tree3.txt
STAR = "(SPECIES_0:0.2,SPECIES_1:0.2,SPECIES_2:0.2,SPECIES_3:0.2,SPECIES_4:0.2)"                      # heterogeneity/description.txt:5
STAR04 = STAR.replace("0.2", "0.4")                                                                   # HKY/description.txt:50

# ratio_positive_negative_testing/description.txt:25-39
PWM_RATIO = np.array([
    [0.0761, 0.0761, 0.7716, 0.0761], [0.1803, 0.1803, 0.1803, 0.4592], [0.497, 0.1677, 0.1677, 0.1677],
    [0.1891, 0.1891, 0.4327, 0.1891], [0.0098, 0.0098, 0.9706, 0.0098], [0.1188, 0.6435, 0.1188, 0.1188],
    [0.1585, 0.1585, 0.5245, 0.1585], [0.4517, 0.1828, 0.1828, 0.1828], [0.1236, 0.1236, 0.6293, 0.1236],
    [0.1543, 0.1543, 0.1543, 0.537], [0.1637, 0.509, 0.1637, 0.1637], [0.0419, 0.8744, 0.0419, 0.0419],
    [0.1913, 0.4261, 0.1913, 0.1913], [0.1582, 0.1582, 0.5254, 0.1582], [0.1986, 0.1986, 0.1986, 0.4042]])
# trees/description_tree1.txt:25-34
PWM_TREE1 = np.array([
    [0.1993, 0.3155, 0.4603, 0.0249], [0.2817, 0.0617, 0.1859, 0.4707], [0.1483, 0.061, 0.4091, 0.3817],
    [0.1673, 0.5331, 0.1219, 0.1777], [0.5726, 0.0482, 0.1283, 0.251], [0.6365, 0.0872, 0.0963, 0.18],
    [0.8143, 0.0465, 0.1381, 0.001], [0.6613, 0.2003, 0.0827, 0.0557], [0.639, 0.0866, 0.2522, 0.0223],
    [0.4891, 0.0454, 0.1373, 0.3281]])
POW = 4 ** np.arange(5)


def load(rel, limit=None):
    smp = PhyloSample.from_fasta(DATA + rel, STAR, limit=limit)      # the Newick only fixes the species order here
    assert smp.species == SP and (smp.codes < 4).all()               # the bundled data hold no gaps
    return smp


def column_hist(smp):
    return np.bincount((smp.codes.astype(np.int64) * POW).sum(axis=1), minlength=1024)


def site_and_flank_hist(smp, W):
    sites, flanks = np.zeros((W, 1024), np.int64), np.zeros(1024, np.int64)
    for i in range(smp.n_alignments):
        idx = (smp.getElementAt(i).astype(np.int64) * POW).sum(axis=1)
        p = int(smp.annotations[i]["MOTIF_POS"])
        for k in range(W):
            sites[k, idx[p + k]] += 1
        flanks += np.bincount(np.concatenate([idx[:p], idx[p + W:]]), minlength=1024)
    return sites, flanks


def main():
    out = {"tree1": np.array(TREE1), "tree2": np.array(TREE2), "tree3": np.array(TREE3), "star": np.array(STAR),
           "star04": np.array(STAR04), "pwm_ratio": PWM_RATIO, "pwm_tree1": PWM_TREE1}
    for key, rel in (("bg_tree1", "trees/data_tree1_1.0.bg"), ("bg_tree2", "trees/data_tree2_1.0.bg"),
                     ("bg_tree3", "trees/data_tree3_1.0.bg"), ("bg_star", "heterogeneity/data_1.0.bg")):
        out[key] = column_hist(load(rel))
    # 6000 alignments each in the files; the first 750 give the same resolution as the other files
    out["bg_hky2"] = column_hist(load("HKY/data_HKY2_1.0.bg", limit=750))
    out["bg_hky05"] = column_hist(load("HKY/data_HKY0.5_1.0.bg", limit=750))
    # HKY on the three trees (HKY+trees/description_tree*.txt: branches 0.2 with HKY(2), 0.4 with HKY(0.5))
    for key, rel in (("bg_tree1_hky2", "HKY+trees/data_tree1_HKY2_1.0.bg"), ("bg_tree1_hky05", "HKY+trees/data_tree1_HKY0.5_1.0.bg"),
                     ("bg_tree3_hky2", "HKY+trees/data_tree3_HKY2_1.0.bg"), ("bg_tree2_hky05", "HKY+trees/data_tree2_HKY0.5_1.0.bg")):
        out[key] = column_hist(load(rel))
    out["sites_ratio"], out["flanks_ratio"] = site_and_flank_hist(load("ratio_positive_negative_testing/data_1.0.fg"), 15)
    out["sites_tree1"], out["flanks_tree1"] = site_and_flank_hist(load("trees/data_tree1_1.0.fg"), 10)
    path = os.path.join(ROOT, "tests", "golden", "generator_pin.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)} bytes")
    for k, v in out.items():
        if v.dtype.kind == "i":
            print(f"  {k}: shape {v.shape}, {int(v.sum())} columns")
    # the first 400 positives of the star-tree file as they are (inputs of the posterior-calibration test: under the generating
    # model the ZOOPS posterior of the E-step is the exact posterior of the planted position)
    smp = load("ratio_positive_negative_testing/data_1.0.fg", limit=400)
    path = os.path.join(ROOT, "tests", "golden", "ratio_star_first400.npz")
    np.savez_compressed(path, codes=smp.codes, col_offsets=smp.col_offsets,
                        motif_pos=np.array([int(a["MOTIF_POS"]) for a in smp.annotations]))
    print(f"wrote {path}: {os.path.getsize(path)} bytes, {smp.n_alignments} alignments")


if __name__ == "__main__":
    main()
