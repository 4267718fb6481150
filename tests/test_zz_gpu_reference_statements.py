"""libphyfoo.so against outputs of the reference's own statements (tests/golden/refexec_meanfield.npz: produced by executing
the reference's methods -- MeanFieldForBayesNet, PhyloBayesModel / PhyloBackground.getLogProbFor, the Jstacs E-step and
normalisations, the M-step objective with its forward-difference gradient, the selector -- translated token by token by
tests/refexec.py; see tests/test_reference_statements.py and scripts/make_refexec_golden.py).  Bars of north_star: scores,
posteriors, objective values within 1e-9 relative, mean-field sweep counts, selected windows and arg-max calls exact.
(Named to run after every other GPU test.)"""
import numpy as np
import pytest

from test_reference_statements import background_pi, cases, motif_pwm

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    from phyfoo_b200 import core
    return core.default_context()


def test_motif_window_scores_equal_the_reference_statements(ctx):
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBayesModel
    n = 0
    for c in cases("motif"):
        W, order = int(c["W"]), int(c["order"])
        fg = PhyloBayesModel(W, order, str(c["newick"]), ctx)
        fg.setCondProbs(motif_pwm(c))
        smp = PhyloSample(np.ascontiguousarray(c["codes"], np.uint8), [0, c["codes"].shape[0]])
        for strand in (0, 1):
            got, sw = fg.getLogProbFor(smp, strand=strand, want_sweeps=True)
            ref, rsw = c["score"][strand], c["sweeps"][strand]
            assert got.shape == ref.shape
            assert np.array_equal(sw, rsw), (str(c["tree"]), order, strand, sw, rsw)
            assert float(np.max(np.abs(got - ref) / np.abs(ref))) < TOL, (str(c["tree"]), order, strand)
            n += got.size
    assert n == 176


def test_background_ranges_equal_the_reference_statements(ctx):
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBackground
    n = 0
    for c in cases("background"):
        order = int(c["order"])
        bg = PhyloBackground(order, str(c["newick"]), ctx)
        bg.setCondProbs(background_pi(c))
        smp = PhyloSample(np.ascontiguousarray(c["codes"], np.uint8), [0, c["codes"].shape[0]])
        for j in range(c["score"].shape[1]):
            got, ref = bg.getLogProbForRange(smp, 0, j, j + order), c["score"][0, j]
            assert abs(got - ref) <= TOL * abs(ref), (str(c["tree"]), order, j)
            n += 1
    assert n == 60


def test_zoops_estep_equals_the_jstacs_statements(ctx):
    """SingleHiddenMotifMixture.getNewWeights on the device against the posterior the Jstacs statements produced."""
    from test_reference_statements import G, estep_inputs
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    newick, W, pwm, zoops = estep_inputs()
    fg = PhyloBayesModel(W, 0, newick, ctx)
    fg.setCondProbs(pwm)
    shm = SingleHiddenMotifMixture(fg, PhyloBackground(0, newick, ctx), zoops=zoops)
    smp = PhyloSample(np.ascontiguousarray(G["estep_codes"], np.uint8), G["estep_col_offsets"])
    res = shm.getNewWeights(smp, seqWeights=G["estep_weights"])
    assert abs(res["ll"] - float(G["estep_ll"])) <= TOL * abs(float(G["estep_ll"]))
    assert np.max(np.abs(res["w"] - G["estep_w"])) <= TOL * np.max(G["estep_w"])
    assert np.max(np.abs(res["gamma"] - G["estep_gamma"])) < TOL
    lens = np.diff(G["estep_col_offsets"]) - W + 1
    woff = np.concatenate([[0], np.cumsum(lens)])
    assert [int(np.argmax(G["estep_gamma"][woff[i]:woff[i + 1]])) for i in range(lens.size)] == res["argmax_off"].tolist()


def test_objective_and_gradient_equal_the_reference_statements(ctx):
    """phyfoo_fe_eval / phyfoo_fe_grad over windows with fixed mean fields against lambda2pi + calcFreeEnergy +
    evaluateGradientOfFunction executed from the reference (every window its own alignment of exactly W columns)."""
    from test_reference_statements import objectives
    from phyfoo_b200 import core
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBayesModel
    for c, pwm in objectives():
        W, order, pos = int(c["W"]), int(c["order"]), int(c["pos"])
        n = c["windows"].shape[0]
        fg = PhyloBayesModel(W, order, str(c["newick"]), ctx)
        fg.setCondProbs(pwm)
        smp = PhyloSample(np.ascontiguousarray(c["windows"].reshape(n * W, -1), np.uint8), np.arange(n + 1) * W)
        fe = core.FESet(ctx, smp.device(ctx), fg.device_model(), np.arange(n), np.zeros(n, np.int32), c["weights"])
        assert len(fe) == n
        f = fe.eval(pos, c["lambda"])
        assert abs(f - float(c["f"])) <= TOL * abs(float(c["f"]))
        f0, g = fe.grad(pos, c["lambda"], 1e-4)
        assert abs(f0 - float(c["f"])) <= TOL * abs(float(c["f"]))
        assert np.all(np.abs(g - c["grad"]) <= TOL * np.abs(c["grad"]) + 64 * np.finfo(float).eps * abs(f0) / 1e-4)
        fe.free()


def test_selector_equals_the_reference_statements(ctx):
    """phyfoo_select on the windows of one alignment against SequenceSpecificDataSelector.prepareForTraining's walk."""
    from test_reference_statements import selectors
    from phyfoo_b200 import core
    from phyfoo_b200.data import PhyloSample
    W = 4
    for c in selectors():
        n = c["weights"].size
        codes = np.random.default_rng(n).integers(0, 4, size=(n + W - 1, 5)).astype(np.uint8)
        smp = PhyloSample(codes, [0, n + W - 1])
        sa, so, sw = core.select(ctx, smp.device(ctx), W, c["weights"], float(c["t"]), float(c["q"]))
        assert np.array_equal(so, c["index"]) and np.array_equal(sw, c["picked"]) and not sa.any()


def test_strand_estep_equals_the_jstacs_statements(ctx):
    from test_reference_statements import G, estep_inputs
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture, StrandModel
    newick, W, pwm, zoops = estep_inputs()
    fg = PhyloBayesModel(W, 0, newick, ctx)
    fg.setCondProbs(pwm)
    shm = SingleHiddenMotifMixture(StrandModel(fg, float(G["strand_forward_weight"])), PhyloBackground(0, newick, ctx), zoops=zoops)
    smp = PhyloSample(np.ascontiguousarray(G["estep_codes"], np.uint8), G["estep_col_offsets"])
    res = shm.getNewWeights(smp, seqWeights=G["estep_weights"])
    assert abs(res["ll"] - float(G["strand_ll"])) <= TOL * abs(float(G["strand_ll"]))
    assert np.max(np.abs(res["w"] - G["strand_w"])) <= TOL * np.max(G["strand_w"])
    assert np.max(np.abs(res["gamma"] - G["strand_gamma"])) < TOL


@pytest.mark.parametrize("order", [1, 2])
def test_estep_with_a_background_of_higher_order_equals_the_reference_driver(ctx, order):
    """The E-step whose background ranges (flanks shorter than order + 1 columns included) come from PhyloBackground.getLogProbFor
    executed from the reference in getNewWeights' call order."""
    from test_reference_statements import G, background_driver
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    k, newick, pis, W, pwm = background_driver(order)
    fg = PhyloBayesModel(W, 0, newick, ctx)
    fg.setCondProbs(pwm)
    bg = PhyloBackground(order, newick, ctx)
    bg.setCondProbs(pis)
    smp = PhyloSample(np.ascontiguousarray(G[k + "estep_codes"], np.uint8), G[k + "estep_col_offsets"])
    res = SingleHiddenMotifMixture(fg, bg, zoops=0.8).getNewWeights(smp, seqWeights=G[k + "estep_weights"])
    assert abs(res["ll"] - float(G[k + "estep_ll"])) <= TOL * abs(float(G[k + "estep_ll"]))
    assert np.max(np.abs(res["w"] - G[k + "estep_w"])) <= TOL * np.max(G[k + "estep_w"])
    assert np.max(np.abs(res["gamma"] - G[k + "estep_gamma"])) < TOL
    # full-length ranges through the range-terms call
    for (s, e), ref in zip(G[k + "calls"], G[k + "values"]):
        if e - s >= order:
            one = PhyloSample(np.ascontiguousarray(G[k + "codes"], np.uint8), [0, G[k + "codes"].shape[0]])
            assert abs(bg.getLogProbForRange(one, 0, int(s), int(e)) - ref) <= TOL * abs(ref)


@pytest.mark.parametrize("order", [0, 1, 2])
def test_alignment_based_model_equals_the_reference_statements(ctx, order):
    from test_reference_statements import alignment_based
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import AlignmentBasedModel
    cond, codes, score = alignment_based(order)
    m = AlignmentBasedModel(len(cond), order, ctx)
    m.setCondProbs(cond)
    got = m.getLogProbFor(PhyloSample(np.ascontiguousarray(codes, np.uint8), [0, codes.shape[0]]), strand=0)
    assert got.shape == score.shape and np.max(np.abs(got - score) / np.abs(score)) < 1e-12


@pytest.mark.parametrize("order", [0, 1, 2])
def test_alignment_based_background_equals_the_reference_statements(ctx, order):
    from test_reference_statements import alignment_based_background
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import AlignmentBasedBGModel
    cond, codes, columns = alignment_based_background(order)
    bg = AlignmentBasedBGModel(order, ctx)
    bg.setCondProbs(cond[0] if order == 0 else cond)
    got = bg.getColumnLogProbs(PhyloSample(np.ascontiguousarray(codes, np.uint8), [0, codes.shape[0]]))
    assert got.shape == columns.shape and np.max(np.abs(got - columns) / np.abs(columns)) < 1e-12
