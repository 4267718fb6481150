import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phyfoo_b200 import core, synth
from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
ctx = core.Context(0)
cfg, newick, pwm, smp, planted = synth.make_config("c2")
fg = PhyloBayesModel(cfg["W"], 0, newick, ctx); fg.setCondProbs(pwm)
bg = PhyloBackground(0, newick, ctx)
shm = SingleHiddenMotifMixture(fg, bg, zoops=0.9)
for i in range(6):
    t0 = time.perf_counter()
    res = shm.getNewWeights(smp, want_gamma=False)
    dt = time.perf_counter() - t0
    print(i, "wall ms", round(dt * 1e3, 3), "gpu_ms", round(res["stats"]["gpu_ms"], 3), [(n, round(ms, 4)) for n, ms in ctx.last_kernel_times()][:4])
print("fp64 peak", ctx.fp64_peak())
for i in range(3):
    res = shm.getNewWeights(smp, want_gamma=False)
    print("after peak", i, round(res["stats"]["gpu_ms"], 3), [(n, round(ms, 4)) for n, ms in ctx.last_kernel_times()][:2])
