"""Wall-clock breakdown of one M-step (selection, fixing q, expected counts, optimiser): time_mstep.py [config]."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phyfoo_b200 import core, synth
from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture

cfg, newick, pwm, sample, planted = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "c2")
ctx = core.default_context()
W = cfg["W"]
fg = PhyloBayesModel(W, cfg["order"], newick, ctx); fg.setCondProbs(pwm)
bg = PhyloBackground(0, newick, ctx)
shm = SingleHiddenMotifMixture(fg, bg, zoops=0.9)
ds = sample.device(ctx)
logw = np.log(shm.weights)
for rep in range(3):
    t = [time.perf_counter()]
    core.zoops_estep(ctx, ds, fg.device_model(), bg.device_model(), logw, None, None, True, False, False); t.append(time.perf_counter())
    sa, so, sw = core.select(ctx, ds, W, None, 0.8, 0.0); t.append(time.perf_counter()); k_sel = ctx.last_kernel_times()
    fe = core.FESet(ctx, ds, fg.device_model(), sa, so, sw); t.append(time.perf_counter()); k_fix = ctx.last_kernel_times()
    ss = core.SSObjective(fe); t.append(time.perf_counter())
    fb, fa, ev = ss.optimize(np.arange(W, dtype=np.int32)); t.append(time.perf_counter())
    pis = [ss.get_pi(p, fg._pi[p].size) for p in range(W)]; t.append(time.perf_counter())
    ss.free(); t.append(time.perf_counter()); fe.free(); t.append(time.perf_counter()); ctx.synchronize(); t.append(time.perf_counter())
    names = ["estep+gamma", "select", "fix q (FESet)", "expected counts (ss_create)", "optimise (ss_optimize)", "get_pi", "ss.free", "fe.free", "sync"]
    print(rep, {n: round((b - a) * 1e3, 3) for n, a, b in zip(names, t, t[1:])}, "evals", ev, "selected", sa.size, "select kernels", k_sel, "fix-q kernels", k_fix)
