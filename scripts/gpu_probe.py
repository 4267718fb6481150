"""Ad-hoc GPU probe: per-kernel times of the C2 E-step and the sweep histogram (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phyfoo_b200 import core, synth
from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
ctx = core.Context(0)
cfg, newick, pwm, smp, planted = synth.make_config(name, n_aln=int(sys.argv[2]) if len(sys.argv) > 2 else None)
fg = PhyloBayesModel(cfg["W"], cfg["order"], newick, ctx); fg.setCondProbs(pwm)
bg = PhyloBackground(0, newick, ctx)
shm = SingleHiddenMotifMixture(fg, bg, zoops=0.9)
for cap in (100, 24, 56):
    ctx.set_option("pattern_memo_sweep_cap", cap)
    for _ in range(3):
        res = shm.getNewWeights(smp)
    print("cap", cap, "gpu_ms", res["stats"]["gpu_ms"], [(n, round(ms, 4)) for n, ms in ctx.last_kernel_times()])
sc, sw = fg.getLogProbFor(smp, want_sweeps=True)
print("K histogram:", np.bincount(sw)[:60].tolist(), "max", sw.max())
bsc, bsw = bg.getColumnLogProbs(smp, want_sweeps=True)
print("bg K histogram:", np.bincount(bsw)[:60].tolist(), "max", bsw.max())
