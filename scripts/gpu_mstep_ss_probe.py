"""Timing of the sufficient-statistics M-step: the device pass vs the host optimisation."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyfoo_b200 import core, synth, optimize
from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture

ctx = core.default_context()
cfg, newick, pwm, sample, planted = synth.make_config("c2")
W = cfg["W"]
fg = PhyloBayesModel(W, 0, newick, ctx); fg.setCondProbs(pwm)
bg = PhyloBackground(0, newick, ctx)
shm = SingleHiddenMotifMixture(fg, bg, zoops=0.9)
ds = sample.device(ctx)
for rep in range(3):
    fg.setCondProbs(pwm)
    shm.getNewWeights(sample, want_gamma=True)
    t0 = time.perf_counter()
    sa, so, sw = core.select(ctx, ds, W, None, 0.8, 0.0)
    fe = core.FESet(ctx, ds, fg.device_model(), sa, so, sw)
    t1 = time.perf_counter()
    obj = optimize.SuffStatObjective(fe, fg._pi, fg._branch, type(fg).EVOL_MODEL)
    t2 = time.perf_counter()
    n = 0
    for pos in range(W):
        fun = lambda z: -obj.eval_lambda(pos, z)
        x, f, it, ev = optimize.quasi_newton_bfgs(fun, lambda z: obj.grad(fun, z), np.log(fg._pi[pos].ravel()), optimize.TC_MOTIF_SEPARATED_POSITIONS())
        n += ev
    t3 = time.perf_counter()
    lam = np.log(fg._pi[0].ravel())
    for _ in range(200):
        obj.eval_lambda(0, lam)
    t4 = time.perf_counter()
    fe.free()
    print(f"select+prepare {1e3*(t1-t0):.2f} ms  suffstats pass {1e3*(t2-t1):.2f} ms  BFGS {1e3*(t3-t2):.2f} ms for {n} evaluations  eval_lambda {5*(t4-t3)*1e3:.1f} us")
