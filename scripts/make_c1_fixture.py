#!/usr/bin/env python
"""Generates tests/golden/c1_tree1_first200.npz: the inputs of BASELINE config 1 -- the first 200 alignments (the reference's
default training subset, io/SampleUtil.java:284) of its bundled data/synthetic_data/trees/data_tree1_1.0.fg as a column
tensor, with the planted MOTIF_POS annotations, the generating PWM and the two trees of description_tree1.txt.  Inputs only
(the FASTA exists only in the build container, /root/reference; the GPU box gets this fixture); bench.py --config c1 scores them.

    python scripts/make_c1_fixture.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phyfoo_b200.data import PhyloSample            # noqa: E402
from scripts.make_golden import CAT, FASTA, PWM, STAR  # noqa: E402


def main():
    smp = PhyloSample.from_fasta(FASTA, CAT, limit=200)
    out = {"codes": smp.codes, "col_offsets": smp.col_offsets, "pwm": PWM / PWM.sum(axis=1, keepdims=True),
           "motif_pos": np.array([int(a["MOTIF_POS"]) for a in smp.annotations]),
           "newick_cat": np.array(CAT), "newick_star": np.array(STAR)}
    path = os.path.join(ROOT, "tests", "golden", "c1_tree1_first200.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)} bytes, {smp.n_alignments} alignments, lengths {smp.lengths().min()}..{smp.lengths().max()}")
    # the negatives of the same data set (no planted site): the first 200 alignments of data_tree1_1.0.bg, for the classification
    # front-end (classification/ClassificationUtil.java: thresholds from the negatives, sensitivity on the positives)
    neg = PhyloSample.from_fasta(FASTA.replace(".fg", ".bg"), CAT, limit=200)
    path = os.path.join(ROOT, "tests", "golden", "c1_tree1_bg_first200.npz")
    np.savez_compressed(path, codes=neg.codes, col_offsets=neg.col_offsets)
    print(f"wrote {path}: {os.path.getsize(path)} bytes, {neg.n_alignments} alignments")


if __name__ == "__main__":
    main()
