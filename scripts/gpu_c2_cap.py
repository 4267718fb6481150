"""C2 E-step time against the pattern table's sweep cap (pattern_memo_sweep_cap): fewer stored sweeps shorten the trajectory
kernel, more windows go through the fallback list to the direct kernel."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyfoo_b200 import core, synth
from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture

cfg, newick, pwm, sample, planted = synth.make_config("c2")
ctx = core.default_context()
fg = PhyloBayesModel(cfg["W"], 0, newick, ctx); fg.setCondProbs(pwm)
bg = PhyloBackground(0, newick, ctx)
ds = sample.device(ctx)
logw = np.log([0.9, 0.1])
ref = None
for cap in (64, 48, 32, 24, 20, 16, 12):
    ctx.set_option("pattern_memo_sweep_cap", cap)
    for _ in range(3):
        res = core.zoops_estep(ctx, ds, fg.device_model(), bg.device_model(), logw, None, None, False, False, False)
    t = []
    for _ in range(10):
        core.zoops_estep(ctx, ds, fg.device_model(), bg.device_model(), logw, None, None, False, False, False)
        t.append(sum(ms for _, ms in ctx.last_kernel_times()))
    res = core.zoops_estep(ctx, ds, fg.device_model(), bg.device_model(), logw, None, None, False, False, False)
    if ref is None:
        ref = res["ll"]
    print(cap, "kernel ms (median of 10)", round(float(np.median(t)), 4), "ll equal", res["ll"] == ref, [(n, round(ms, 4)) for n, ms in ctx.last_kernel_times()])
