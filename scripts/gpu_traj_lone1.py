import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from phyfoo_b200 import core, synth
from phyfoo_b200.models import PhyloBayesModel
ctx = core.default_context()
ctx.set_option("pattern_memo", 2)
cfg, newick, pwm, sample, planted = synth.make_config("c2", n_aln=200)
fg = PhyloBayesModel(1, 0, newick, ctx); fg.setCondProbs(pwm[:1])
for _ in range(3):
    fg.getLogProbFor(sample)
print([ (n, ms) for n, ms in ctx.last_kernel_times()])
