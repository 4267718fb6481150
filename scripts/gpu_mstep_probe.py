"""Where the M-step time goes on C2: selection, fixing q (prepare), one gradient, per-call host overhead."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyfoo_b200 import core, synth
from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture

ctx = core.default_context()
cfg, newick, pwm, sample, planted = synth.make_config("c2")
W = cfg["W"]
fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(pwm)
bg = PhyloBackground(0, newick, ctx)
shm = SingleHiddenMotifMixture(fg, bg, zoops=0.9)
ds = sample.device(ctx)
shm.getNewWeights(sample, want_gamma=True)
for rep in range(3):
    t0 = time.perf_counter()
    sa, so, sw = core.select(ctx, ds, W, None, 0.8, 0.0)
    t1 = time.perf_counter()
    k_sel = ctx.last_kernel_times()
    fe = core.FESet(ctx, ds, fg.device_model(), sa, so, sw)
    ctx.synchronize()
    t2 = time.perf_counter()
    k_prep = ctx.last_kernel_times()
    lam = np.log(np.asarray(fg.getCondProb(0)).ravel())
    fe.grad(0, lam)
    t3 = time.perf_counter()
    for _ in range(100):
        fe.grad(0, lam)
    t4 = time.perf_counter()
    for _ in range(100):
        fe.eval(0, lam)
    t5 = time.perf_counter()
    k_grad = ctx.last_kernel_times()
    c, e = fe.suffstats(W, 4 + 16 * (fg.tree.n_nodes - 1))
    t6 = time.perf_counter()
    fe.free()
    print(f"select {1e3*(t1-t0):.3f} ms  prepare {1e3*(t2-t1):.3f} ms  first grad {1e3*(t3-t2):.3f}  grad {10*(t4-t3):.4f} ms  eval {10*(t5-t4):.4f} ms  suffstats {1e3*(t6-t5):.3f} ms  selected {sa.size}")
print("select kernels", k_sel)
print("prepare kernels", k_prep)
print("eval kernels", k_grad)
