"""Scores the windows of a c3o0 / c5-shaped sample once with the order-0 motif (for ncu captures of the scoring kernel alone)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyfoo_b200 import core, synth
from phyfoo_b200.models import PhyloBayesModel

name = sys.argv[1] if len(sys.argv) > 1 else "c3o0"
n_aln = int(sys.argv[2]) if len(sys.argv) > 2 else 600
kernel = int(sys.argv[3]) if len(sys.argv) > 3 else -1
cfg, newick, pwm, sample, planted = synth.make_config(name, n_aln=n_aln)
ctx = core.default_context()
if len(sys.argv) > 5:
    ctx.set_option("jit_max_threads", int(sys.argv[5]))
if len(sys.argv) > 4:
    ctx.set_option("l2_persist", int(sys.argv[4]))
ctx.set_option("order0_kernel", kernel)
fg = PhyloBayesModel(cfg["W"], 0, newick, ctx)
fg.setCondProbs(pwm)
for _ in range(2):
    out = fg.getLogProbFor(sample)
    print(ctx.last_kernel_times(), fg.last_stats)
print(float(out.sum()))
