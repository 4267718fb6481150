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
        AlignmentBasedModel(5, 3, ctx)


@pytest.mark.parametrize("order", [1, 2])
def test_alignment_based_models_of_higher_order_against_oracle(order):
    """Orders 1 and 2 (AlignmentBasedModel.java:97-228, AlignmentBasedBGModel.java:40-222) against the oracle's
    statement-by-statement restatement of the Java, 15 % gaps so that every expansion case occurs (including the stale list
    entries of the background likelihood): scores at 1e-12 relative (a logarithm of a short sum), trained tables at 1e-12."""
    from phyfoo_b200 import core
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import AlignmentBasedBGModel, AlignmentBasedModel
    ctx = core.default_context()
    rng = np.random.default_rng(40 + order)
    for S in (1, 4, 17):
        W = 6
        lens = [W, 31, 2, 45]
        smp = PhyloSample(random_codes(rng, sum(lens), S, 0.15), np.concatenate([[0], np.cumsum(lens)]))
        cond = [rng.dirichlet(np.ones(4), size=4 ** min(p, order)).ravel() for p in range(W)]
        m = AlignmentBasedModel(W, order, ctx)
        m.setCondProbs(cond)
        for strand in (0, 1):
            got = m.getLogProbFor(smp, strand=strand)
            ref = np.concatenate([orc.ab_score_windows_order(cond, order, smp.getElementAt(i), bool(strand)) for i in range(smp.n_alignments)])
            assert got.shape == ref.shape and np.max(np.abs(got - ref) / np.abs(ref)) < 1e-12
        bgc = [rng.dirichlet(np.ones(4), size=4 ** p).ravel() for p in range(order + 1)]
        bg = AlignmentBasedBGModel(order, ctx)
        bg.setCondProbs(bgc)
        cols = bg.getColumnLogProbs(smp)
        for i in range(smp.n_alignments):
            a, b = int(smp.col_offsets[i]), int(smp.col_offsets[i + 1])
            codes = smp.getElementAt(i)
            for s, e in ((0, b - a - 1), (0, 0), (min(3, b - a - 1), b - a - 1), (1, min(4, b - a - 1))):
                if e < s:
                    continue
                r = orc.ab_bg_range_order(bgc, order, codes, s, e)              # the context reaches back before `s`
                assert abs(cols[a + s:a + e + 1].sum() - r) <= 1e-12 * abs(r)
        # count M-step of the motif on the top-posterior windows
        gamma = rng.random(smp.device(ctx).windows(W)) ** 8
        assert m.train(smp, gamma)["skipped"]
        st = m.train(smp, gamma)
        sa, so, sw = core.select(ctx, smp.device(ctx), W, gamma)
        sel_cols = np.stack([smp.getElementAt(a)[o:o + W] for a, o in zip(sa, so)])
        ref = orc.ab_train_order(sel_cols, sw, order, 1e-8)
        assert st["selected"] == sa.size
        for got_t, ref_t in zip(m.getCondProbs(), ref):
            assert np.max(np.abs(got_t - ref_t)) < 1e-12
        # background counts over whole alignments, weighted per alignment
        w = rng.random(smp.n_alignments) + 0.5
        bg.train(smp, w)
        ref = orc.ab_bg_train_order([smp.getElementAt(i) for i in range(smp.n_alignments)], w, order, 1e-2)
        for got_t, ref_t in zip(bg.getCondProbs(), ref):
            assert np.max(np.abs(got_t - ref_t)) < 1e-12
        bg.train(smp)
        ref = orc.ab_bg_train_order([smp.getElementAt(i) for i in range(smp.n_alignments)], None, order, 1e-2)
        for got_t, ref_t in zip(bg.getCondProbs(), ref):
            assert np.max(np.abs(got_t - ref_t)) < 1e-12
