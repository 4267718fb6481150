#!/usr/bin/env python
"""Generates tests/golden/heterogeneity_bg_first60.npz: the first 60 alignments of the reference's bundled
data/synthetic_data/heterogeneity/data_1.0.bg (star tree, every branch 0.2) and data_rBG_1.0.bg (star tree, every branch of
every alignment drawn from beta(3, 7), mean 0.3) as column tensors -- inputs for the tree-learning tests
(training/TrainingUtil.java:72-80, io/PhyloSample.java:243-270).  The FASTA files exist only in the build container.

    python scripts/make_heterogeneity_fixture.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phyfoo_b200.data import PhyloSample            # noqa: E402

DIR = "/root/reference/data/synthetic_data/heterogeneity"
STAR = "(SPECIES_0:0.2,SPECIES_1:0.2,SPECIES_2:0.2,SPECIES_3:0.2,SPECIES_4:0.2)"


def main():
    out = {"newick_star": np.array(STAR)}
    for key, name in (("homogeneous", "data_1.0.bg"), ("heterogeneous", "data_rBG_1.0.bg")):
        smp = PhyloSample.from_fasta(os.path.join(DIR, name), STAR, limit=60)
        out[f"codes_{key}"], out[f"col_offsets_{key}"] = smp.codes, smp.col_offsets
    path = os.path.join(ROOT, "tests", "golden", "heterogeneity_bg_first60.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
