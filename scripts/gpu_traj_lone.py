"""How long does the trajectory kernel take when every warp has a scheduler to itself?  (C2's tree and patterns, motif width 1..12:
1024 x W trees.)  The time at W = 1 is the latency of the longest chain; the growth with W is the issue-bound part."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyfoo_b200 import core, synth
from phyfoo_b200.models import PhyloBayesModel

ctx = core.default_context()
ctx.set_option("pattern_memo", 2)
cfg, newick, pwm, sample, planted = synth.make_config("c2", n_aln=200)
for W in (1, 2, 4, 8, 12):
    fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(pwm[:W])
    for wave in (1, 0):
        ctx.set_option("pattern_memo_wave", wave)
        ts = []
        for _ in range(5):
            fg.getLogProbFor(sample)
            ts.append([ms for n, ms in ctx.last_kernel_times() if n == "memo_traj_kernel"][0])
        print(f"W={W:2d} wave={wave}  traj {np.median(ts)*1e3:.1f} us")
