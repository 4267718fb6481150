"""Where the end-to-end E-step time goes (C2): dataset create (H2D + pack), E-step call (kernels + D2H), free."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyfoo_b200 import core, synth
from phyfoo_b200.data import PhyloSample
from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture

ctx = core.default_context()
cfg, newick, pwm, sample, planted = synth.make_config("c2")
W = cfg["W"]
fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(pwm)
bg = PhyloBackground(0, newick, ctx)
shm = SingleHiddenMotifMixture(fg, bg, zoops=1.0)
pinned = ctx.pinned_empty(sample.codes.shape, np.uint8); pinned[:] = sample.codes
n_windows = sample.device(ctx).windows(W)
out = {"gamma": ctx.pinned_empty(n_windows), "argmax_off": ctx.pinned_empty(sample.n_alignments, np.int32),
       "argmax_strand": ctx.pinned_empty(sample.n_alignments, np.uint8)}
T = {"create": [], "estep": [], "estep_nogamma": [], "free": [], "estep_kernels": []}
for it in range(12):
    t0 = time.perf_counter()
    smp = PhyloSample(pinned, sample.col_offsets)
    ds = smp.device(ctx)
    t1 = time.perf_counter()
    res = shm.getNewWeights(smp, want_gamma=True, out=out)
    t2 = time.perf_counter()
    k = sum(ms for _, ms in ctx.last_kernel_times())
    first_list = ctx.last_kernel_times()
    res2 = shm.getNewWeights(smp, want_gamma=False)
    t3 = time.perf_counter()
    ds.free()
    t4 = time.perf_counter()
    if it >= 2:
        T["create"].append(t1 - t0); T["estep"].append(t2 - t1); T["estep_nogamma"].append(t3 - t2); T["free"].append(t4 - t3)
        T["estep_kernels"].append(k * 1e-3)
for k, v in T.items():
    print(f"{k:15s} median {np.median(v) * 1e3:.3f} ms   min {np.min(v) * 1e3:.3f} ms")
print("first E-step on a new data set:", [(n, round(ms, 4)) for n, ms in first_list])
