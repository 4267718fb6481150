"""Non-phylogenetic alignment models of order 0 (models/AlignmentBasedModel.java, AlignmentBasedBGModel.java): window scores
and background columns bit-exact against the oracle (pure sums of table values in the reference's order), the count M-step
within fp64 re-association."""
import numpy as np
import pytest

from util import orc, random_codes

pytestmark = pytest.mark.gpu


def test_alignment_based_models_against_oracle():
    from phyfoo_b200 import core
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import AlignmentBasedBGModel, AlignmentBasedModel
    ctx = core.default_context()
    rng = np.random.default_rng(3)
    for S in (1, 5, 12, 30):
        W = 9
        lens = [W, 40, 3, 77]
        smp = PhyloSample(random_codes(rng, sum(lens), S, 0.1), np.concatenate([[0], np.cumsum(lens)]))
        cp = rng.dirichlet(np.ones(4), size=W)
        m = AlignmentBasedModel(W, 0, ctx); m.setCondProbs(cp)
        for strand in (0, 1):
            got = m.getLogProbFor(smp, strand=strand)
            ref = np.concatenate([orc.ab_score_windows(cp, smp.getElementAt(i), bool(strand)) for i in range(smp.n_alignments)])
            assert np.array_equal(got, ref)
        bg = AlignmentBasedBGModel(0, ctx); bg.setCondProbs([0.1, 0.2, 0.3, 0.4])
        cols = bg.getColumnLogProbs(smp)
        for i in range(smp.n_alignments):
            a, b = smp.col_offsets[i], smp.col_offsets[i + 1]
            r = orc.ab_bg_range([0.1, 0.2, 0.3, 0.4], smp.getElementAt(i), 0, int(b - a) - 1)
            assert abs(cols[a:b].sum() - r) <= 1e-12 * abs(r)
            assert cols[a] == orc.ab_bg_range([0.1, 0.2, 0.3, 0.4], smp.getElementAt(i), 0, 0)
        # count M-step on the top-posterior windows
        gamma = rng.random(smp.device(ctx).windows(W)) ** 8
        assert m.train(smp, gamma)["skipped"]
        st = m.train(smp, gamma)
        sa, so, sw = core.select(ctx, smp.device(ctx), W, gamma)
        sel_cols = np.stack([smp.getElementAt(a)[o:o + W] for a, o in zip(sa, so)])
        ref = orc.ab_train(sel_cols, sw, 1e-8)
        assert st["selected"] == sa.size and np.max(np.abs(m.getCondProbs() - ref)) < 1e-12
    with pytest.raises(NotImplementedError):
        AlignmentBasedModel(5, 1, ctx)
