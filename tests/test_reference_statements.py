"""The mean-field sweep and the free energy pinned by the reference's own statements.

tests/refexec.py reads algorithm/MeanFieldForBayesNet.java from /root/reference, translates init(), optimizeByNormalisation()
and calcFreeEnergy(int, int) statement by statement and executes them (there is no JVM here; nothing of the file is copied
into the repository).  scripts/make_refexec_golden.py ran them on seeded inputs and committed inputs and outputs as
tests/golden/refexec_meanfield.npz.  Here:
  * anywhere: the C++ oracle and the independent numpy restatement against those golden vectors -- scores to the last bits,
    mean-field sweep counts exactly (SURVEY.md section 8 rows a5, a6, a7, and a10 for ranges of order + 1 columns);
  * where /root/reference exists: the translation is redone and re-executed, and must reproduce the committed vectors.
The CUDA path is held against the same vectors in tests/test_zz_gpu_reference_statements.py.
"""
import os

import numpy as np
import pytest

from util import oracle_bg, oracle_fg

from oracle import np_restatement as npr

import refexec

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "refexec_meanfield.npz"))
HAVE_REFERENCE = os.path.exists(refexec.REFERENCE_FILE)
needs_reference = pytest.mark.skipif(not HAVE_REFERENCE, reason="/root/reference is not on this machine")


def cases(kind):
    for i in range(int(G["n_cases"])):
        if str(G[f"c{i}_kind"]) == kind:
            yield {k: G[f"c{i}_{k}"] for k in ("newick", "order", "W", "pwm", "codes", "score", "sweeps", "tree", "gap_rate")}


def motif_pwm(c):
    """per position [contexts][4]: one context row at position 0 and for order 0, four behind an order-1 edge"""
    W, order, flat = int(c["W"]), int(c["order"]), np.asarray(c["pwm"])
    rows = [1 if (order == 0 or t == 0) else 4 for t in range(W)]
    off = np.concatenate([[0], np.cumsum(rows)]) * 4
    return [flat[off[t]:off[t + 1]].reshape(rows[t], 4) for t in range(W)]


def background_pi(c):
    order, flat = int(c["order"]), np.asarray(c["pwm"])
    rows = [4 ** t for t in range(order + 1)]
    off = np.concatenate([[0], np.cumsum(rows)]) * 4
    return [flat[off[t]:off[t + 1]].reshape(rows[t], 4) for t in range(order + 1)]


def test_golden_vectors_cover_the_path():
    motif, bg = list(cases("motif")), list(cases("background"))
    assert len(motif) == 16 and len(bg) == 12
    assert {int(c["order"]) for c in motif} == {0, 1} and {int(c["order"]) for c in bg} == {0, 1, 2}
    assert any((c["codes"] == 4).any() for c in motif) and any(not (c["codes"] == 4).any() for c in motif)
    sweeps = np.concatenate([c["sweeps"].ravel() for c in motif])
    assert sweeps.min() >= 2 and sweeps.max() > 20 and len(set(sweeps.tolist())) > 10      # the stop rule is exercised


def test_oracle_reproduces_the_reference_statements_on_motif_windows():
    worst, identical, total = 0.0, 0, 0
    for c in cases("motif"):
        W, pwm = int(c["W"]), motif_pwm(c)
        om = oracle_fg(str(c["newick"]), pwm, order=int(c["order"]))
        for strand in (0, 1):
            got, sw = om.score_windows(c["codes"], bool(strand))
            assert np.array_equal(sw, c["sweeps"][strand]), (str(c["tree"]), int(c["order"]), strand)
            worst = max(worst, float(np.max(np.abs(got - c["score"][strand]) / np.abs(c["score"][strand]))))
            identical += int((got == c["score"][strand]).sum())
            total += got.size
    assert worst < 1e-13
    assert identical >= 0.9 * total            # same statements in the same order: nearly every score is the same double


def test_oracle_reproduces_the_reference_statements_on_background_ranges():
    for c in cases("background"):
        order = int(c["order"])
        ob = oracle_bg(str(c["newick"]), order, background_pi(c))
        for j in range(c["score"].shape[1]):
            ob.reset_stats()
            got = ob.bg_range(c["codes"], j, j + order)
            ref = c["score"][0, j]
            assert abs(got - ref) <= 1e-13 * abs(ref)
            assert ob.sweeps_total() == c["sweeps"][0, j]            # both optimisations of the range's net


def test_numpy_restatement_reproduces_the_reference_statements():
    for c in cases("motif"):
        W, order = int(c["W"]), int(c["order"])
        net = npr.Net.motif(str(c["newick"]), W, order)
        for t, p in enumerate(motif_pwm(c)):
            net.set_pi(t, p)
        net.build_tables()
        for strand in (0, 1):
            for j in range(c["score"].shape[1]):
                v, k = npr.score_window(net, c["codes"][j:j + W], bool(strand))
                assert k == c["sweeps"][strand, j]
                assert abs(v - c["score"][strand, j]) <= 1e-12 * abs(v)


@needs_reference
def test_translation_is_of_the_three_methods_and_nothing_is_rewritten():
    src = refexec.translate()
    assert set(src) == {"init", "optimizeByNormalisation", "calcFreeEnergy2", "calcFreeEnergy0"}
    opt, fe = src["optimizeByNormalisation"], src["calcFreeEnergy2"]
    # landmarks of the reference's text (MeanFieldForBayesNet.java:100-107,192-201,204-209,351-357)
    assert "maxSteps = 100" in opt and "minsteps = 2" in opt and "< 1e-5" in opt
    assert "while k < maxSteps and not converged or k < minsteps:" in opt
    assert opt.count("_jlog") == 2 and opt.count("_jexp") == 1 and "/ sum" in opt
    assert fe.count("_jlog") == 4 and fe.count("_NEG_INF") == 2 and "return part1 - part2" in fe
    for s in src.values():
        compile(s, "translated", "exec")


@needs_reference
def test_reexecuting_the_reference_statements_reproduces_the_committed_vectors():
    for c in cases("motif"):
        if str(c["tree"]) == "rand12" and int(c["order"]) == 1:
            continue                                                   # (seconds of pure Python; covered by the generator run)
        W, order = int(c["W"]), int(c["order"])
        net = npr.Net.motif(str(c["newick"]), W, order)
        for t, p in enumerate(motif_pwm(c)):
            net.set_pi(t, p)
        net.build_tables()
        ref = refexec.ReferenceMeanField(net)
        for strand in (0, 1):
            for j in range(0, c["score"].shape[1], 2):
                v, k = ref.score(c["codes"][j:j + W], bool(strand))
                assert k == c["sweeps"][strand, j] and v == c["score"][strand, j]


# ---- the ZOOPS E-step -------------------------------------------------------------------
def estep_inputs():
    W = int(G["estep_W"])
    pwm = [G["estep_pwm"][4 * t:4 * t + 4].reshape(1, 4) for t in range(W)]
    return str(G["estep_newick"]), W, pwm, float(G["estep_zoops"])


def test_oracle_estep_reproduces_the_jstacs_statements():
    """SingleHiddenMotifMixture.getNewWeights / modify and Normalisation.logSumNormalisation executed from libs/jstacs.jar on
    scores that the mean-field statements produced: the oracle's E-step gives the same log-likelihood, component statistics
    and posterior, and so do its window scores and background columns."""
    from util import orc
    newick, W, pwm, zoops = estep_inputs()
    fg, bg = oracle_fg(newick, pwm), oracle_bg(newick, 0)
    off, codes = G["estep_col_offsets"], G["estep_codes"]
    scores = np.concatenate([fg.score_windows(codes[off[i]:off[i + 1]])[0] for i in range(off.size - 1)])
    assert np.max(np.abs(scores - G["estep_scores"]) / np.abs(G["estep_scores"])) < 1e-13
    assert np.max(np.abs(bg.bg_columns(codes) - G["estep_bg_columns"]) / np.abs(G["estep_bg_columns"])) < 1e-13
    res = orc.estep(fg, bg, off, codes, np.log([zoops, 1 - zoops]), None, G["estep_weights"])
    assert abs(res["ll"] - float(G["estep_ll"])) <= 1e-13 * abs(float(G["estep_ll"]))
    assert np.max(np.abs(res["w"] - G["estep_w"])) <= 1e-13 * np.max(G["estep_w"])
    assert np.max(np.abs(res["gamma"] - G["estep_gamma"])) < 1e-13
    # arg-max site call = first maximum of the posterior (util/Util.java:528-538)
    lens = np.diff(off) - W + 1
    woff = np.concatenate([[0], np.cumsum(lens)])
    assert [int(np.argmax(G["estep_gamma"][woff[i]:woff[i + 1]])) for i in range(lens.size)] == res["argmax_off"].tolist()
    # sum of an alignment's posterior = its probability of holding a site, times its weight
    assert abs(G["estep_gamma"].sum() - G["estep_w"][0]) < 1e-12 and abs(G["estep_w"].sum() - G["estep_weights"].sum()) < 1e-12


@needs_reference
def test_jstacs_estep_statements_translate_and_reproduce_the_committed_vectors():
    src = refexec.translate_estep()
    gw = src["getNewWeights"]
    # landmarks of jstacs SingleHiddenMotifMixture.java:533-564 and Normalisation.java:352-376
    assert "b1 = - self.bgMaxMarkovOrder" in gw and "b2 = motifLength + self.bgMaxMarkovOrder - 1" in gw
    assert gw.count("getLogProbFor") == 5 and "getLogPriorForPositions ( l , j )" in gw
    assert "ll += currentWeight * ( logPSeq + self.modify ( help , seqweights [ 0 ] , start , w0Index ) )" in gw
    assert "if sum != 1:" in src["lsn7"] and "return offset + _jlog ( sum )" in src["lsn7"]
    assert "return - _jlog ( seqLength - self.motifLength + 1 )" in src["getLogPriorForPositions"]
    newick, W, pwm, zoops = estep_inputs()
    off = G["estep_col_offsets"]
    col = G["estep_bg_columns"]

    def background(i, s, e):
        log_sum = 0.0
        for c in range(s, e + 1):
            log_sum += -col[off[i] + c]
        return -log_sum if e >= s else 0.0

    es = refexec.ReferenceEStep(np.diff(off), W, lambda w: G["estep_scores"][w], background, zoops, 0)
    ll, w, gamma = es.getNewWeights(G["estep_weights"])
    assert ll == float(G["estep_ll"]) and np.array_equal(w, G["estep_w"]) and np.array_equal(gamma, G["estep_gamma"])


# ---- substitution tables (row a3) -------------------------------------------------------------------------------------------
@needs_reference
def test_substitution_tables_equal_the_reference_reinit_statements():
    """reinit() of FS81alpha / FS81beta / HKY executed from the reference's files against the tables the restatements build:
    oracle/np_restatement.f81alpha (what the golden vectors above ran on) and phyfoo_b200.evolution.transition (the numpy mirror
    of csrc/model_host.hpp::transition) -- bit for bit (FS81beta: to an ulp of its one exponential), including FS81beta's clamps and HKY with the field at 0 (as shipped)
    and at the intended ratio."""
    from phyfoo_b200 import _abi, evolution
    rng = np.random.default_rng(12)
    for _ in range(5):
        pi = rng.dirichlet(np.ones(4), size=4)
        for dist in (0.0005, 0.05, 0.2, 0.37, 0.9, 1.0):
            ref = np.array(refexec.ReferenceEvolModel("FS81alpha", pi).reinit(dist))
            assert np.array_equal(ref, npr.f81alpha(pi, dist))
            for c in range(4):
                assert np.array_equal(ref[4 * c:4 * c + 4], evolution.transition(pi[c], dist, _abi.EVOL_F81ALPHA))
        for dist in (1e-6, 0.01, 0.2, 1.5, 40.0):                 # 1e-6 and 40 hit the two clamps (FS81beta.java:88-92)
            ref = np.array(refexec.ReferenceEvolModel("FS81beta", pi).reinit(dist))
            for c in range(4):
                mine = evolution.transition(pi[c], dist, _abi.EVOL_F81BETA)
                assert np.max(np.abs(ref[4 * c:4 * c + 4] - mine)) <= 4e-16      # one exp: Math.exp / libm / numpy agree to an ulp
        for dist, field in ((0.2, 0.0), (0.2, 2.0), (0.4, 0.5)):
            ref = np.array(refexec.ReferenceEvolModel("HKY", pi[:1], hky_field=field).reinit(dist))
            assert np.array_equal(ref, evolution.hky_transition(pi[0], dist, dist * field))
            if field == 0.0:
                assert (ref == 0).sum() == 8                      # every transversion: probability zero as shipped


# ---- the M-step objective and its forward-difference gradient (rows a14, a15) -------------------------------------------------
def objectives():
    for i in range(int(G["n_objectives"])):
        c = {k: G[f"obj{i}_{k}"] for k in ("newick", "order", "W", "pos", "pwm", "windows", "weights", "lambda", "f", "grad")}
        yield c, motif_pwm(c)


def test_oracle_objective_and_gradient_reproduce_the_reference_statements():
    from util import orc
    n = 0
    for c, pwm in objectives():
        m = oracle_fg(str(c["newick"]), pwm, order=int(c["order"]))
        fs = orc.FESet(m, np.ascontiguousarray(c["windows"], np.uint8), c["weights"])
        f = fs.eval(int(c["pos"]), c["lambda"])
        assert abs(f - float(c["f"])) <= 1e-13 * abs(float(c["f"]))
        f0, g = fs.grad(int(c["pos"]), c["lambda"], 1e-4)
        assert abs(f0 - float(c["f"])) <= 1e-13 * abs(float(c["f"]))
        # forward differences of two sums of magnitude |f|: a last-bit difference in f shows as ulp(f) / eps in g
        assert np.all(np.abs(g - c["grad"]) <= 4 * np.finfo(float).eps * abs(f0) / 1e-4)
        assert c["grad"].size == (16 if int(c["order"]) == 1 and int(c["pos"]) > 0 else 4)
        n += 1
    assert n == 5


@needs_reference
def test_reexecuting_the_objective_statements_reproduces_the_committed_vectors():
    for c, pwm in objectives():
        net = npr.Net.motif(str(c["newick"]), int(c["W"]), int(c["order"]))
        for t, p in enumerate(pwm):
            net.set_pi(t, p)
        net.build_tables()
        obj = refexec.ReferenceObjective(net, list(c["windows"]), c["weights"], int(c["pos"]))
        assert obj.evaluateFunction(list(c["lambda"])) == float(c["f"])
        assert np.array_equal(obj.evaluateGradientOfFunction(list(c["lambda"])), c["grad"])
    src = refexec.ReferenceObjective.gradient_source
    assert "gradient [ i ] = ( self.evaluateFunction ( x ) - current ) / self.eps" in src and "x [ i ] += self.eps" in src


# ---- the selector of the M-step (row a13) ---------------------------------------------------------------------------------------
def selectors():
    for i in range(int(G["n_selectors"])):
        yield {k: G[f"sel{i}_{k}"] for k in ("weights", "t", "q", "index", "picked")}


def test_oracle_selector_reproduces_the_reference_statements():
    from util import orc
    for c in selectors():
        idx, w = orc.select_group(c["weights"], float(c["t"]), float(c["q"]))
        assert np.array_equal(idx, c["index"]) and np.array_equal(w, c["picked"])
        assert 1 <= idx.size <= c["weights"].size and np.all(np.diff(w) <= 0)           # from the largest posterior down


@needs_reference
def test_reexecuting_the_selector_statements_reproduces_the_committed_vectors():
    for c in selectors():
        idx, w = refexec.ReferenceSelector(list(c["weights"])).prepareForTraining(float(c["t"]), float(c["q"]))
        assert np.array_equal(idx, c["index"]) and np.array_equal(w, c["picked"])
    assert "if sum > t:" in refexec.ReferenceSelector.source and "if self.weightProfile [ i ] [ k ] >= q:" in refexec.ReferenceSelector.source


# ---- strand mixture inside the E-step (row a11) ---------------------------------------------------------------------------------
def test_oracle_strand_estep_reproduces_the_jstacs_statements():
    """A StrandModel as motif component: per window Normalisation.getLogSum(logW0 + forward, logW1 + reverse complement), then
    the same getNewWeights -- all from the reference's statements; the oracle's strand E-step gives the same numbers."""
    from util import orc
    newick, W, pwm, zoops = estep_inputs()
    fg, bg = oracle_fg(newick, pwm), oracle_bg(newick, 0)
    off, codes = G["estep_col_offsets"], G["estep_codes"]
    rev = np.concatenate([fg.score_windows(codes[off[i]:off[i + 1]], True)[0] for i in range(off.size - 1)])
    assert np.max(np.abs(rev - G["strand_rev_scores"]) / np.abs(rev)) < 1e-13
    fw = float(G["strand_forward_weight"])
    res = orc.estep(fg, bg, off, codes, np.log([zoops, 1 - zoops]), np.log([fw, 1 - fw]), G["estep_weights"])
    assert abs(res["ll"] - float(G["strand_ll"])) <= 1e-13 * abs(float(G["strand_ll"]))
    assert np.max(np.abs(res["w"] - G["strand_w"])) <= 1e-13 * np.max(G["strand_w"])
    assert np.max(np.abs(res["gamma"] - G["strand_gamma"])) < 1e-13
    assert abs(float(G["strand_ll"]) - float(G["estep_ll"])) > 1e-3          # (not the single-strand E-step)


@needs_reference
def test_cpf_lookup_rows_equal_the_reference_statements():
    """Row l of a CPF table stands for parent values query[p] = (l / 4^p) % 4 (bayesNet/CPF.java:238-256 executed): parent 0 --
    the tree parent -- is the least significant digit, which is the row layout of the restatements and of the kernels' tables."""
    for parents in range(4):
        table, _ = refexec.reference_lookup_table(parents)
        assert table == [[(l // 4 ** p) % 4 for p in range(parents)] for l in range(4 ** parents)]


@needs_reference
def test_site_call_is_the_first_maximum():
    """Util.whichMax executed (row a17): strict `>`, so of equal maxima the first wins -- numpy's argmax rule, which the
    oracle and the posterior kernel follow (arg-max calls are compared bit for bit everywhere)."""
    rng = np.random.default_rng(17)
    for _ in range(50):
        v = rng.integers(0, 4, size=rng.integers(1, 12)).astype(float)          # many ties
        assert refexec.reference_which_max(v) == int(np.argmax(v))
    assert refexec.reference_which_max(G["estep_gamma"][1:11]) == int(np.argmax(G["estep_gamma"][1:11]))


# ---- PhyloBackground.getLogProbFor executed whole, over the shared net (row a10) ------------------------------------------------
def background_driver(order):
    k = f"bgd{order}_"
    flat = G[k + "pi"]
    rows = [4 ** t for t in range(order + 1)]
    o = np.concatenate([[0], np.cumsum(rows)]) * 4
    pis = [flat[o[t]:o[t + 1]].reshape(rows[t], 4) for t in range(order + 1)]
    W = int(G[k + "W"])
    pwm = [G[k + "pwm"][4 * t:4 * t + 4].reshape(1, 4) for t in range(W)]
    return k, str(G[k + "newick"]), pis, W, pwm


@pytest.mark.parametrize("order", [1, 2])
def test_oracle_background_reproduces_the_reference_driver(order):
    """Range values in the reference's call order on one model object -- including ranges shorter than order + 1 columns, whose
    trailing trees keep the observation an earlier call left (MeanFieldForBayesNet.java:84) -- and the ZOOPS E-step with that
    background, against PhyloBackground.getLogProbFor + getNewWeights executed from the reference."""
    from util import orc
    k, newick, pis, W, pwm = background_driver(order)
    ob = oracle_bg(newick, order, pis)
    for (s, e), ref in zip(G[k + "calls"], G[k + "values"]):
        got = ob.bg_range(G[k + "codes"], int(s), int(e))
        assert abs(got - ref) <= 1e-13 * abs(ref) if ref else got == 0
    short = [(int(s), int(e)) for s, e in G[k + "calls"] if 0 <= e - s < order]
    assert len(short) >= 3
    res = orc.estep(oracle_fg(newick, pwm), oracle_bg(newick, order, pis), G[k + "estep_col_offsets"], G[k + "estep_codes"],
                    np.log([0.8, 0.2]), None, G[k + "estep_weights"])
    assert abs(res["ll"] - float(G[k + "estep_ll"])) <= 1e-13 * abs(float(G[k + "estep_ll"]))
    assert np.max(np.abs(res["w"] - G[k + "estep_w"])) <= 1e-13 * np.max(G[k + "estep_w"])
    assert np.max(np.abs(res["gamma"] - G[k + "estep_gamma"])) < 1e-13


@needs_reference
@pytest.mark.parametrize("order", [1, 2])
def test_reexecuting_the_background_driver_reproduces_the_committed_vectors(order):
    k, newick, pis, W, pwm = background_driver(order)
    net = npr.Net.background(newick, order)
    for t, p in enumerate(pis):
        net.set_pi(t, p)
    net.build_tables()
    rb = refexec.ReferenceBackground(net)
    seq = refexec._MDSeq(G[k + "codes"])
    assert [rb.getLogProbFor(seq, int(s), int(e)) for s, e in G[k + "calls"]] == G[k + "values"].tolist()
    src = refexec.ReferenceBackground.source
    assert "while i < self.order and startpos <= endpos:" in src and "logSum += tmpPointer . get ( i ) . calcFreeEnergy ( self.order , self.order )" in src
    assert "min ( self.bnh . motifLength , self.seq . getLength ( ) )" in refexec._SharedMeanField.init_observation_source


@needs_reference
def test_motif_get_log_prob_for_executed_whole_reproduces_the_committed_vectors():
    """PhyloBayesModel.getLogProbFor (row a9) executed from the reference -- sub-sequence, mean field from the cache that never
    hits, initObservation, optimizeByNormalisation, -calcFreeEnergy() -- gives the committed scores and sweep counts."""
    n = 0
    for c in cases("motif"):
        if str(c["tree"]) == "rand12":
            continue
        W, order = int(c["W"]), int(c["order"])
        net = npr.Net.motif(str(c["newick"]), W, order)
        for t, p in enumerate(motif_pwm(c)):
            net.set_pi(t, p)
        net.build_tables()
        model, seq = refexec.ReferenceMotif(net), refexec._MDSeq(c["codes"])
        for j in range(0, c["score"].shape[1], 2):
            v, k = model.getLogProbFor(seq, j, j + W - 1)
            assert v == c["score"][0, j] and k == c["sweeps"][0, j]
            n += 1
    assert n == 36
    assert "logLikelihood = - tmpPointer . calcFreeEnergy ( )" in refexec.ReferenceMotif.motif_source


# ---- the non-phylogenetic motif model (row f-3) ---------------------------------------------------------------------------------
def alignment_based(order):
    k = f"ab{order}_"
    W, flat = int(G[k + "W"]), G[k + "cond"]
    sizes = [4 * 4 ** min(p, order) for p in range(W)]
    off = np.concatenate([[0], np.cumsum(sizes)])
    return [flat[off[p]:off[p + 1]] for p in range(W)], G[k + "codes"], G[k + "score"]


@pytest.mark.parametrize("order", [0, 1, 2])
def test_oracle_alignment_based_model_reproduces_the_reference_statements(order):
    from util import orc
    cond, codes, score = alignment_based(order)
    got = orc.ab_score_windows_order(cond, order, codes)
    assert got.shape == score.shape and np.max(np.abs(got - score) / np.abs(score)) < 1e-13
    assert (codes == 4).any()


@needs_reference
def test_alignment_based_statements_against_every_gap_pattern():
    """AlignmentBasedModel.getLnCondProb executed from the reference against the oracle's restatement for every pattern of
    symbols and gaps in the context (orders 0 to 2): the same doubles."""
    import itertools
    from util import orc
    rng = np.random.default_rng(2)
    for order in (0, 1, 2):
        W = order + 2
        cond = [rng.dirichlet(np.ones(4), size=4 ** min(p, order)).ravel() for p in range(W)]
        model = refexec.ReferenceAlignmentBasedModel(cond, order)
        for row in itertools.product(range(5), repeat=W):
            for pos in range(W):
                assert model.getLnCondProb(list(row), pos, pos) == orc.ab_motif_ln_cond_prob(cond, order, list(row), pos, pos)
    assert "self.ACCESS_TABLE [ accessHead ] = self.ACCESS_TABLE [ i ] + a * self.pows [ k ]" in refexec.ReferenceAlignmentBasedModel.source


def alignment_based_background(order):
    k = f"ab{order}_"
    flat = G[k + "bg_cond"]
    sizes = [4 * 4 ** p for p in range(order + 1)]
    off = np.concatenate([[0], np.cumsum(sizes)])
    return [flat[off[p]:off[p + 1]] for p in range(order + 1)], G[k + "codes"], G[k + "bg_columns"]


@pytest.mark.parametrize("order", [0, 1, 2])
def test_oracle_alignment_based_background_reproduces_the_reference_statements(order):
    from util import orc
    cond, codes, columns = alignment_based_background(order)
    for s, e in ((0, codes.shape[0] - 1), (3, 9), (5, 5)):
        got = orc.ab_bg_range_order(cond, order, codes, s, e)
        ref = columns[s:e + 1].sum()
        assert abs(got - ref) <= 1e-12 * abs(ref)


@needs_reference
def test_alignment_based_background_statements_against_every_gap_pattern():
    """AlignmentBasedBGModel.getLnCondProb executed from the reference (with the list entries it keeps from before the last gap
    expansion, :185-191) against the oracle's restatement: the same doubles for every pattern (orders 0 to 2)."""
    import itertools
    from util import orc
    rng = np.random.default_rng(3)
    for order in (0, 1, 2):
        cond = [rng.dirichlet(np.ones(4), size=4 ** p).ravel() for p in range(order + 1)]
        model = refexec.ReferenceAlignmentBasedBackground(cond, order)
        for row in itertools.product(range(5), repeat=order + 2):
            for pos in range(order + 2):
                assert model.getLnCondProb(list(row), pos) == orc.ab_bg_ln_cond_prob(cond, order, list(row), pos)


# ---- branch-length learning (row f-2) -------------------------------------------------------------------------------------------
def test_oracle_branch_length_objective_reproduces_the_reference_statements():
    """EdgeOptimizer.evaluateFunction executed from the reference (parameters -> branch lengths (e^x + 0.0005) / (1 + e^x), the
    tables of the position, the weighted free-energy sum over fixed mean fields) against the oracle under the branch lengths
    the host mirror's map gives (phyfoo_b200.evolution.branch_from_logit, what trainEdges optimises over)."""
    from util import orc
    from phyfoo_b200 import evolution
    W, pos = int(G["edge_W"]), int(G["edge_pos"])
    pwm = [G["edge_pwm"][4 * t:4 * t + 4].reshape(1, 4) for t in range(W)]
    m = oracle_fg(str(G["edge_newick"]), pwm)
    fs = orc.FESet(m, np.ascontiguousarray(G["edge_windows"], np.uint8), G["edge_weights"])
    for x, f in zip(G["edge_x"], G["edge_f"]):
        alpha = evolution.branch_from_logit(x)
        assert alpha.min() >= 0.0005 * (1 - 1e-12) and alpha.max() <= 1
        for k in range(1, x.size + 1):
            m.set_branch(pos, k, float(alpha[k - 1]))
        assert abs(fs.sum() - f) <= 1e-13 * abs(f)


@needs_reference
def test_reexecuting_the_branch_length_objective_reproduces_the_committed_vectors():
    W, pos = int(G["edge_W"]), int(G["edge_pos"])
    net = npr.Net.motif(str(G["edge_newick"]), W, 0)
    for t in range(W):
        net.set_pi(t, G["edge_pwm"][4 * t:4 * t + 4].reshape(1, 4))
    net.build_tables()
    obj = refexec.ReferenceEdgeObjective(net, list(G["edge_windows"]), G["edge_weights"], pos)
    assert [obj.evaluateEdges(x) for x in G["edge_x"]] == G["edge_f"].tolist()
    assert "( _jexp ( edgesExp [ i ] ) + self.MINIMUM_BRANCH_LENGTH ) / ( 1 + _jexp ( edgesExp [ i ] ) )" in refexec.ReferenceEdgeObjective.edge_source


# ---- the optimiser (row a16) ------------------------------------------------------------------------------------------------------
@needs_reference
def test_bfgs_and_line_search_executed_from_the_jar_take_the_restatements_path():
    """Optimizer.quasiNewtonBFGS with findBracket and brentsMethod, executed from libs/jstacs.jar on the M-step objective of a
    motif position (the library's sufficient-statistics form, forward-difference gradient): the same number of iterations and
    of function evaluations as phyfoo_b200.optimize.quasi_newton_bfgs and the same final point to the last bit -- and that
    restatement is the one tests/test_optimize_host.py holds evaluation for evaluation against the C++ in libphyfoo.so."""
    from phyfoo_b200 import _abi, core, optimize, synth
    from phyfoo_b200.newick import parse_newick
    tree = parse_newick("((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.2):0.25):0.1)")
    W = 3
    rng = np.random.default_rng(21)
    pwm = synth.random_pwm(W, 0, 9)
    counts = rng.uniform(0.0, 2.0, (W, 4 + 16 * (tree.n_nodes - 1)))
    desc, keep = _abi.make_desc(_abi.MODEL_FG, W, 0, tree, np.tile(tree.branch, (W, 1)), pwm, _abi.EVOL_F81ALPHA, _abi.MODE_MEANFIELD)
    ss = core.SSObjective(desc=desc, counts=counts, entropy=0.75)
    iterations = []
    for pos in range(W):
        fun = lambda lam: -ss.eval(pos, np.asarray(lam, float))                                    # noqa: E731

        def grad(lam, eps=1e-4):
            lam = np.array(lam, float)
            f0, g = fun(lam), np.zeros(lam.size)
            for i in range(lam.size):
                h = lam[i]
                lam[i] += eps
                g[i] = (fun(lam) - f0) / eps
                lam[i] = h
            return f0, g
        x0 = np.log(np.ravel(pwm[pos]))
        x, f, it, evals = optimize.quasi_newton_bfgs(fun, grad, x0.copy(), optimize.TC_MOTIF_SEPARATED_POSITIONS(), 1e-3, 0.1)
        ref = refexec.ReferenceOptimizer(fun, lambda lam: grad(lam)[1])
        # termination condition and start-distance forecaster: PhyFoo's (PreparedConditions.java:22-26, FE_byParameter...:74-80),
        # their doNextIteration / getNewStartDistance / setLastDistance executed from the jar as well
        xr, itr, evr = ref.minimise(list(x0), refexec.reference_condition("TC_MOTIF_SEPARATED_POSITIONS"), refexec.reference_forecaster(10, 0.1), 1e-3)
        assert (it, evals) == (itr, evr) and np.array_equal(x, xr)
        iterations.append(it)
    assert max(iterations) >= 3
    src = refexec.ReferenceOptimizer.source
    assert "erg [ 4 ] = self.limFibPlus2 * erg [ 2 ] - self.limFibPlus1 * erg [ 0 ]" in src["findBracket"]
    assert "u [ counter1 ] = s [ counter1 ] / sv - matrixv [ counter1 ] / vmatrixv" in src["quasiNewtonBFGS"]


@needs_reference
def test_termination_conditions_and_forecaster_executed_from_the_jar_equal_the_restatements():
    from phyfoo_b200 import optimize
    rng = np.random.default_rng(1)
    for name, mine in (("TC_MOTIF_SEPARATED_POSITIONS", optimize.TC_MOTIF_SEPARATED_POSITIONS()), ("TC_MOTIF_ALL_POSITIONS", optimize.TC_MOTIF_ALL_POSITIONS())):
        ref, seen = refexec.reference_condition(name), set()
        for _ in range(400):
            it = int(rng.integers(0, 7))
            f_last, f_cur = rng.normal(size=2) * 10.0 ** rng.integers(-6, 1)
            g = rng.normal(size=4) * 10.0 ** rng.integers(-6, 0)
            d = rng.normal(size=4) * 10.0 ** rng.integers(-4, 1)
            alpha = float(rng.random() * 10.0 ** rng.integers(-3, 1))
            r = ref.do_next_iteration(it, f_last, f_cur, list(g), list(d), alpha)
            assert r == mine.do_next_iteration(it, f_last, f_cur, g, d, alpha)
            seen.add(bool(r))
        assert seen == {True, False}
    ref, mine = refexec.reference_forecaster(10, 0.1), optimize.LimitedMedianStartDistance(10, 0.1)
    for _ in range(25):
        assert ref.get_new_start_distance() == mine.get_new_start_distance()
        v = float(rng.random())
        ref.set_last_distance(v)
        mine.set_last_distance(v)


@needs_reference
def test_strand_model_m_step_weights_executed_from_the_jar_equal_the_host_mirror():
    """StrandModel.getNewWeights (jstacs StrandModel.java:561-583, with modifyWeights and logSumNormalisation) executed from the
    jar against phyfoo_b200.models.StrandModel.getNewWeights fed with the same window scores: per-window strand weights, the
    strand statistics (without the prior) and the log-likelihood."""
    from phyfoo_b200.models import StrandModel
    rng = np.random.default_rng(5)
    forward, reverse, weights = rng.normal(size=40) * 3 - 20, rng.normal(size=40) * 3 - 20, rng.random(40)

    class Scores:                                             # the motif model: forward and reverse-complement window scores
        def getLogProbFor(self, sample, strand):
            return reverse if strand else forward
    for dw in (weights, None):
        L, w, sw = refexec.ReferenceStrandEStep(forward, reverse, 0.6).getNewWeights(dw)
        f, r, w_mine, L_mine = StrandModel(Scores(), 0.6).getNewWeights(None, dw)
        assert abs(L - L_mine) <= 1e-13 * abs(L)
        assert np.max(np.abs(np.array(sw[0::2]) - f)) < 1e-14 and np.max(np.abs(np.array(sw[1::2]) - r)) < 1e-14
        assert np.max(np.abs(np.array(w) - w_mine)) <= 1e-13 * max(w)
    assert "w [ 1 ] += seqweights [ 0 ] [ counter1 ]" in refexec.ReferenceStrandEStep.source


# ---- classification front-end (row f-1) -----------------------------------------------------------------------------------------
@needs_reference
def test_classification_thresholds_and_sensitivity_executed_from_the_reference_equal_the_host_mirror():
    """ClassificationUtil.determineThresholdsForSpecifitiesPerPos / determineSensitivityPerSeq (:244-278) executed from the
    reference against phyfoo_b200.classification on ragged score matrices (the device variants are held against these host
    functions in tests/test_gpu_scan.py)."""
    from phyfoo_b200 import classification as cl
    ref = refexec.reference_classification()
    rng = np.random.default_rng(4)
    for _ in range(5):
        neg = [rng.normal(size=rng.integers(1, 30)) for _ in range(20)]
        pos = [rng.normal(size=rng.integers(1, 30)) + 1 for _ in range(20)]
        spec = [0.0, 0.5, 0.9, 0.99, 0.999, 1.0]
        thresholds = ref["thresholds"](neg, spec)
        assert np.array_equal(thresholds, cl.determineThresholdsForSpecifitiesPerPos(neg, spec))
        for t in thresholds:
            assert ref["sensitivity"](t, pos) == cl.determineSensitivityPerSeq(t, pos)


@needs_reference
def test_global_branch_length_objective_executed_from_the_reference_equals_the_oracle():
    """GlobalEdgeOptimizer.evaluateFunction (one set of branch lengths for the trees of every position; what EDGE_LEARNING = 2
    optimises, models/PhyloBayesModel.java:204) executed from the reference against the oracle under those branch lengths."""
    from util import orc, random_codes
    from phyfoo_b200 import evolution, synth
    rng = np.random.default_rng(12)
    newick, W = "((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.4):0.25):0.1,F:0.3)", 3
    pwm = synth.random_pwm(W, 0, 33)
    net = npr.Net.motif(newick, W, 0)
    for t, p in enumerate(pwm):
        net.set_pi(t, p)
    net.build_tables()
    cols = [random_codes(rng, W, 6, 0.1) for _ in range(5)]
    weights = rng.random(5) + 0.1
    ref = refexec.ReferenceGlobalEdgeObjective(net, cols, weights)
    m = oracle_fg(newick, pwm)
    fs = orc.FESet(m, np.stack(cols), weights)
    for _ in range(3):
        x = rng.normal(size=net.N - 1) - 1.0
        alpha = evolution.branch_from_logit(x)
        for t in range(W):
            for k in range(1, net.N):
                m.set_branch(t, k, float(alpha[k - 1]))
        f = ref.evaluateEdges(x)
        assert abs(fs.sum() - f) <= 1e-13 * abs(f)
    assert "bn . setDistanceToParent ( edgesPi [ k ] )" in refexec.ReferenceGlobalEdgeObjective.global_source
