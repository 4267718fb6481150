"""Committed golden vectors (tests/golden/c1_tree1_first8.npz, made by scripts/make_golden.py from the first 8
alignments of the reference's bundled trees/data_tree1_1.0.fg): CPU tests keep the oracle and the host encoder pinned
to them; the GPU tests compare the CUDA path (through the C ABI) with the same vectors."""
import os

import numpy as np
import pytest

from util import ROOT, orc, rel_err

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "c1_tree1_first8.npz"))
W, N_ALN = 10, 8
KEYS = ["cat", "star", "cat_gapped", "star_gapped"]


def _newick(key):
    return str(GOLD["newick_cat"] if key.startswith("cat") else GOLD["newick_star"])


def _codes(key):
    return GOLD["codes_gapped"] if key.endswith("gapped") else GOLD["codes"]


def test_host_encoder_reproduces_the_fixture_codes():
    from phyfoo_b200.data import encode, parse_annotation
    from phyfoo_b200.newick import species_order
    order = species_order(_newick("cat"))
    assert order == [f"SPECIES_{i}" for i in range(5)]
    heads, seqs = GOLD["fasta_headers"], GOLD["fasta_seqs"]
    off = GOLD["col_offsets"]
    for i in range(N_ALN):
        rows = {}
        for h, s in zip(heads[5 * i:5 * i + 5], seqs[5 * i:5 * i + 5]):
            ann = parse_annotation(str(h))
            assert ann["gene"] == f"id_{i}" and int(ann["MOTIF_POS"]) == GOLD["motif_pos"][i]
            rows[ann["species"]] = str(s)
        mat = np.stack([encode(rows[name]) for name in order], axis=1)
        assert np.array_equal(mat, GOLD["codes"][off[i]:off[i + 1]])
        assert np.array_equal(orc.encode_alignment([rows[name] for name in order]), mat)
        # the planted site is the upper-case run of width 10 starting at MOTIF_POS
        s0 = rows["SPECIES_0"]
        p = int(GOLD["motif_pos"][i])
        assert s0[p:p + W].isupper() and not s0[:p].isupper() and (p + W == len(s0) or s0[p + W].islower())


@pytest.mark.parametrize("key", KEYS)
def test_oracle_reproduces_golden(key):
    newick, codes, off, pwm = _newick(key), _codes(key), GOLD["col_offsets"], GOLD["pwm"]
    fg = orc.Model(orc.FG, W, 0, newick)
    for t in range(W):
        fg.set_pi(t, pwm[t])
    bg = orc.Model(orc.BG, 0, 0, newick)
    sc = np.concatenate([fg.score_windows(codes[off[i]:off[i + 1]])[0] for i in range(N_ALN)])
    sw = np.concatenate([fg.score_windows(codes[off[i]:off[i + 1]])[1] for i in range(N_ALN)])
    assert rel_err(sc, GOLD[f"mf_scores_{key}"]) < 1e-13 and np.array_equal(sw, GOLD[f"mf_sweeps_{key}"])
    assert rel_err(bg.bg_columns(codes), GOLD[f"bg_columns_{key}"]) < 1e-13
    res = orc.estep(fg, bg, off, codes, np.log([0.9, 0.1]))
    assert np.array_equal(res["argmax_off"], GOLD[f"estep_argmax_off_{key}"])
    assert rel_err([res["ll"], *res["w"]], GOLD[f"estep_llw_{key}"]) < 1e-13
    if key == "cat_gapped":
        b, g = orc.pack_columns(codes)
        assert np.array_equal(b, GOLD["packed_bases_gapped"]) and np.array_equal(g, GOLD["packed_gaps_gapped"])


def test_golden_is_sane():
    # mean field never exceeds the exact log-likelihood; on the star tree they coincide
    for key in KEYS:
        assert (GOLD[f"mf_scores_{key}"] <= GOLD[f"exact_scores_{key}"] + 1e-9).all()
    assert rel_err(GOLD["mf_scores_star"], GOLD["exact_scores_star"]) < 1e-9
    assert (GOLD["mf_sweeps_star"] == 2).all()
    assert (GOLD["estep_argmax_off_cat"] == GOLD["motif_pos"]).sum() >= 6


@pytest.mark.gpu
@pytest.mark.parametrize("key", KEYS)
def test_gpu_matches_golden(key):
    from phyfoo_b200 import core
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture, StrandModel
    ctx = core.default_context()
    newick, codes, pwm = _newick(key), _codes(key), GOLD["pwm"]
    smp = PhyloSample(codes, GOLD["col_offsets"])
    fg = PhyloBayesModel(W, 0, newick, ctx)
    fg.setCondProbs([p.reshape(1, 4) for p in pwm])
    bg = PhyloBackground(0, newick, ctx)
    sc, sw = fg.getLogProbFor(smp, want_sweeps=True)
    assert rel_err(sc, GOLD[f"mf_scores_{key}"]) < 1e-9
    assert np.array_equal(sw, GOLD[f"mf_sweeps_{key}"])
    PhyloBayesModel.CALC_LOGLIKELIHOOD = True
    try:
        assert rel_err(fg.getLogProbFor(smp), GOLD[f"exact_scores_{key}"]) < 1e-9
    finally:
        PhyloBayesModel.CALC_LOGLIKELIHOOD = False
    assert rel_err(bg.getColumnLogProbs(smp), GOLD[f"bg_columns_{key}"]) < 1e-9
    if key == "cat_gapped":
        b, g = smp.device(ctx).packed()
        assert np.array_equal(b, GOLD["packed_bases_gapped"]) and np.array_equal(g, GOLD["packed_gaps_gapped"])
    for strand in (False, True):
        sk = f"{key}{'_strand' if strand else ''}"
        shm = SingleHiddenMotifMixture(StrandModel(fg, 0.5) if strand else fg, bg, zoops=0.9)
        res = shm.getNewWeights(smp)
        assert np.array_equal(res["argmax_off"], GOLD[f"estep_argmax_off_{sk}"])
        assert np.array_equal(res["argmax_strand"], GOLD[f"estep_argmax_strand_{sk}"])
        assert np.max(np.abs(res["gamma"] - GOLD[f"estep_gamma_{sk}"])) < 1e-9
        assert rel_err([res["ll"], *res["w"]], GOLD[f"estep_llw_{sk}"]) < 1e-9
    # M-step: selection bit-exact, objective and forward-difference gradient
    sa, so, swt = core.select(ctx, smp.device(ctx), W, GOLD[f"estep_gamma_{key}"])
    assert np.array_equal(sa, GOLD[f"sel_aln_{key}"]) and np.array_equal(so, GOLD[f"sel_off_{key}"])
    assert np.array_equal(swt, GOLD[f"sel_w_{key}"])
    fe = core.FESet(ctx, smp.device(ctx), fg.device_model(), sa, so, swt)
    f0, g = fe.grad(3, GOLD[f"fe_lambda_{key}"], 1e-4)
    r0, rg = float(GOLD[f"fe_f0_{key}"]), GOLD[f"fe_grad_{key}"]
    assert abs(f0 - r0) <= 1e-9 * abs(r0)
    assert np.all(np.abs(g - rg) <= 1e-9 * np.abs(rg) + 64 * np.finfo(float).eps * abs(r0) / 1e-4)


def test_c1_fixture_is_the_first_200_bundled_alignments_and_the_oracle_finds_the_planted_sites():
    """BASELINE configs[0] (bench.py --config c1): tests/golden/c1_tree1_first200.npz (scripts/make_c1_fixture.py) extends
    the 8-alignment fixture; under the generating PWM and tree the oracle's arg-max site is the annotated MOTIF_POS in most
    alignments (the number bench.py reports for the CUDA path as `planted_sites`)."""
    from phyfoo_b200 import synth
    cfg, newick, pwm, smp, planted = synth.make_config("c1")
    assert smp.n_alignments == 200 and cfg["W"] == 10 and len(pwm) == 10 and newick == _newick("cat")
    off8 = GOLD["col_offsets"]
    assert np.array_equal(smp.col_offsets[:N_ALN + 1], off8) and np.array_equal(smp.codes[:off8[-1]], GOLD["codes"])
    assert np.array_equal(planted[:N_ALN], GOLD["motif_pos"])
    lens = smp.lengths()
    assert lens.min() >= 300 and lens.max() <= 400 and np.all(planted + 10 <= lens)
    fg = orc.Model(orc.FG, 10, 0, newick)
    for t, p in enumerate(pwm):
        fg.set_pi(t, np.asarray(p).ravel())
    sub = smp.slice(0, 40)
    ref = orc.estep(fg, orc.Model(orc.BG, 0, 0, newick), sub.col_offsets, sub.codes, np.log([0.9, 0.1]))
    assert np.mean(ref["argmax_off"] == planted[:40]) >= 0.6


def test_c4_shards_are_distinct_blocks_with_one_planted_site_of_one_motif():
    from phyfoo_b200 import synth
    a = synth.make_config_multi("c4", n_aln=64, shard=(0, 2))
    b = synth.make_config_multi("c4", n_aln=64, shard=(1, 2))
    assert a[3].n_alignments == b[3].n_alignments == 32 and not np.array_equal(a[3].codes, b[3].codes)
    assert len(a[2]) == 3 and all(np.allclose(x, y) for x, y in zip(a[2], b[2]))          # the same three PWMs on every rank
    planted, kind = a[4], a[5]
    assert np.array_equal(planted >= 0, kind >= 0) and set(np.unique(kind)) <= {-1, 0, 1, 2}


# ---- backgrounds of orders 1 and 2 (tests/golden/c1_tree1_first8_bgorders.npz, scripts/make_golden_bgorders.py) ----------
GOLD_BG = np.load(os.path.join(ROOT, "tests", "golden", "c1_tree1_first8_bgorders.npz"))
N_ESTEP = 3


def _bg_pis(order):
    return [GOLD_BG[f"pi_o{order}_t{t}"] for t in range(order + 1)]


@pytest.mark.parametrize("order", [1, 2])
def test_oracle_reproduces_golden_background_orders(order):
    from util import oracle_bg
    codes, off = GOLD["codes_gapped"], GOLD["col_offsets"]
    bg = oracle_bg(_newick("cat"), order, _bg_pis(order))
    for i in (0, 5):                                               # (the E-step vectors are checked on the GPU side)
        a = codes[off[i]:off[i + 1]]
        assert abs(bg.bg_range(a, 0, a.shape[0] - 1) - GOLD_BG[f"logprob_o{order}"][i]) <= 1e-12 * abs(GOLD_BG[f"logprob_o{order}"][i])
        assert abs(bg.bg_range(a, 7, 7 + order + 11) - GOLD_BG[f"range_o{order}"][i]) <= 1e-12 * abs(GOLD_BG[f"range_o{order}"][i])
    assert np.all(GOLD_BG[f"logprob_o{order}"] < 0) and np.isfinite(GOLD_BG[f"estep_llw_o{order}"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("order", [1, 2])
def test_gpu_matches_golden_background_orders(order):
    from phyfoo_b200 import core
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
    ctx = core.default_context()
    newick, codes, off = _newick("cat"), GOLD["codes_gapped"], GOLD["col_offsets"]
    smp = PhyloSample(codes, off)
    bg = PhyloBackground(order, newick, ctx); bg.setCondProbs(_bg_pis(order))
    assert rel_err(bg.getLogProbFor(smp), GOLD_BG[f"logprob_o{order}"]) < 1e-9
    got = [bg.getLogProbForRange(smp, i, 7, 7 + order + 11) for i in range(N_ALN)]
    assert rel_err(got, GOLD_BG[f"range_o{order}"]) < 1e-9
    fg = PhyloBayesModel(W, 0, newick, ctx)
    fg.setCondProbs([p.reshape(1, 4) for p in GOLD["pwm"]])
    res = SingleHiddenMotifMixture(fg, bg, zoops=0.9).getNewWeights(smp.slice(0, N_ESTEP))
    assert np.array_equal(res["argmax_off"], GOLD_BG[f"estep_argmax_off_o{order}"])
    assert np.max(np.abs(res["gamma"] - GOLD_BG[f"estep_gamma_o{order}"])) < 1e-9
    assert rel_err([res["ll"], *res["w"]], GOLD_BG[f"estep_llw_o{order}"]) < 1e-9
