"""Scores the windows of a C3-shaped sample once with the order-1 motif (for ncu captures of the scoring kernel alone)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phyfoo_b200 import core, synth
from phyfoo_b200.models import PhyloBayesModel

n_aln = int(sys.argv[1]) if len(sys.argv) > 1 else 500
kernel = int(sys.argv[2]) if len(sys.argv) > 2 else 2
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
qglobal = int(sys.argv[4]) if len(sys.argv) > 4 else 0
cfg, newick, pwm, sample, planted = synth.make_config("c3", n_aln=n_aln)
ctx = core.default_context()
if len(sys.argv) > 5:
    ctx.set_option("l2_persist", int(sys.argv[5]))
ctx.set_option("order1_kernel", kernel)
ctx.set_option("order1_qglobal", qglobal)
fg = PhyloBayesModel(cfg["W"], 1, newick, ctx)
fg.setCondProbs(pwm)
for _ in range(reps):
    out = fg.getLogProbFor(sample)
    print(ctx.last_kernel_times(), fg.last_stats)
print(float(out.sum()))
