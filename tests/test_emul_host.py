"""The kernels' per-thread arithmetic (phyfoo_b200/csrc/tree_pass.cuh, __host__ __device__) compiled for the
host and checked against the oracle: scores within 1e-9 relative, mean-field sweep counts identical.
This is the CPU-side guard for the GPU parity tests (tests/test_gpu_parity.py does the same through the C ABI)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from util import ROOT, orc, oracle_fg, random_codes, rel_err

from phyfoo_b200 import _abi, synth
from phyfoo_b200.newick import parse_newick

EMUL_DIR = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL_DIR, "libemul.so")
    srcs = [os.path.join(EMUL_DIR, "emul.cpp"), os.path.join(ROOT, "phyfoo_b200", "csrc", "tree_pass.cuh"),
            os.path.join(ROOT, "phyfoo_b200", "csrc", "model_host.hpp"), os.path.join(ROOT, "phyfoo_b200", "csrc", "fp64_math.cuh"),
            os.path.join(ROOT, "phyfoo_b200", "csrc", "fp64_math_tables.h"),
            os.path.join(ROOT, "phyfoo_b200", "csrc", "net1n.cuh"), os.path.join(ROOT, "phyfoo_b200", "csrc", "tree_pass_f81.cuh"),
            os.path.join(ROOT, "phyfoo_b200", "csrc", "ab_expand.hpp"), os.path.join(ROOT, "phyfoo_b200", "csrc", "netg.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", srcs[0], "-o", so])
    L = C.CDLL(so)
    L.emul_score_windows.argtypes = [C.POINTER(_abi.ModelDesc), C.POINTER(C.c_uint8), C.c_int, C.c_int,
                                     C.POINTER(C.c_double), C.POINTER(C.c_int)]
    for f in (L.emul_exp, L.emul_log):
        f.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        f.restype = None
    L.emul_ab_term.argtypes = [C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int]
    L.emul_ab_term.restype = C.c_double
    L.emul_ab_expand.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    return L


def run_emul(L, newick, pwm, codes, strand, mode, evol=_abi.EVOL_F81ALPHA, hky_ratio=0.0):
    tree = parse_newick(newick)
    W = len(pwm)
    desc, keep = _abi.make_desc(_abi.MODEL_FG, W, 0, tree, np.tile(tree.branch, (W, 1)), pwm, evol, mode, hky_ratio)
    n = codes.shape[0] - W + 1
    out = np.zeros(n)
    sw = np.zeros(n, np.int32)
    rc = L.emul_score_windows(C.byref(desc), codes.ctypes.data_as(C.POINTER(C.c_uint8)), codes.shape[0], strand,
                              out.ctypes.data_as(C.POINTER(C.c_double)), sw.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0
    return out, sw


TREES = [synth.star_newick(5), synth.caterpillar_newick(5), synth.random_binary_newick(12, 3),
         "((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.4):0.25):0.1,F:0.3)"]


@pytest.mark.parametrize("newick", TREES)
@pytest.mark.parametrize("gap_rate", [0.0, 0.1])
@pytest.mark.parametrize("strand", [0, 1])
def test_meanfield_pass_matches_oracle(emul, newick, gap_rate, strand):
    rng = np.random.default_rng(11)
    S = parse_newick(newick).n_species
    W = 6
    pwm = synth.random_pwm(W, 0, 29)
    codes = random_codes(rng, 40, S, gap_rate)
    got, sw = run_emul(emul, newick, pwm, codes, strand, _abi.MODE_MEANFIELD)
    ref, rsw = oracle_fg(newick, pwm).score_windows(codes, bool(strand))
    assert np.array_equal(sw, rsw)
    assert rel_err(got, ref) < 1e-9


@pytest.mark.parametrize("newick", TREES)
def test_pruning_matches_oracle(emul, newick):
    rng = np.random.default_rng(12)
    S = parse_newick(newick).n_species
    W = 5
    pwm = synth.random_pwm(W, 0, 31)
    codes = random_codes(rng, 30, S, 0.15)
    got, _ = run_emul(emul, newick, pwm, codes, 0, _abi.MODE_EXACT)
    ref, _ = oracle_fg(newick, pwm, mode=orc.EXACT_PRUNING).score_windows(codes)
    assert rel_err(got, ref) < 1e-12


def test_f81beta_tables(emul):
    rng = np.random.default_rng(13)
    newick = synth.caterpillar_newick(5, 0.3)
    pwm = synth.random_pwm(4, 0, 37)
    codes = random_codes(rng, 20, 5)
    got, sw = run_emul(emul, newick, pwm, codes, 0, _abi.MODE_MEANFIELD, _abi.EVOL_F81BETA)
    ref, rsw = oracle_fg(newick, pwm, evol=orc.F81BETA).score_windows(codes)
    assert np.array_equal(sw, rsw)
    assert rel_err(got, ref) < 1e-9


@pytest.mark.parametrize("ratio", [2.0, 0.5])
@pytest.mark.parametrize("gap_rate", [0.0, 0.15])
def test_hky_with_its_ratio_in_effect(emul, ratio, gap_rate):
    """PHYFOO_EVOL_HKY_RATIO (opt-in; evolution/HKY.java:77-106 with beta = alpha * ratio, the model of the reference's bundled
    HKY data): general 4x4 tables, so the generic pass scores -- mean field with identical sweep counts, pruning at 1e-12."""
    rng = np.random.default_rng(14)
    pwm = synth.random_pwm(4, 0, 39)
    for newick in (synth.caterpillar_newick(5, 0.2), TREES[3]):
        S = parse_newick(newick).n_species
        codes = random_codes(rng, 24, S, gap_rate)
        for strand in (0, 1):
            got, sw = run_emul(emul, newick, pwm, codes, strand, _abi.MODE_MEANFIELD, _abi.EVOL_HKY_RATIO, ratio)
            ref, rsw = oracle_fg(newick, pwm, evol=orc.HKY_RATIO, hky_ratio=ratio).score_windows(codes, bool(strand))
            assert np.array_equal(sw, rsw)
            assert rel_err(got, ref) < 1e-9
        got, _ = run_emul(emul, newick, pwm, codes, 0, _abi.MODE_EXACT, _abi.EVOL_HKY_RATIO, ratio)
        ref, _ = oracle_fg(newick, pwm, mode=orc.EXACT_PRUNING, evol=orc.HKY_RATIO, hky_ratio=ratio).score_windows(codes)
        assert rel_err(got, ref) < 1e-12
        lit, _ = oracle_fg(newick, pwm, mode=orc.EXACT_ELIMINATION, evol=orc.HKY_RATIO, hky_ratio=ratio).score_windows(codes)
        assert rel_err(lit, ref) < 1e-12


@pytest.mark.parametrize("newick", TREES)
@pytest.mark.parametrize("gap_rate", [0.0, 0.15])
def test_fixed_q_objective_matches_oracle(emul, newick, gap_rate):
    """store_q + expected_log_cpf_multi (the M-step kernels' arithmetic) against the oracle's FESet: f(lambda) within
    1e-9 relative; the forward-difference gradient within 1e-9 * |g| plus the cancellation noise of f/eps."""
    emul.emul_fe_values.argtypes = [C.POINTER(_abi.ModelDesc), C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_double), C.c_int,
                                    C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double)]
    rng = np.random.default_rng(17)
    tree = parse_newick(newick)
    S, W, n, pos, eps = tree.n_species, 4, 9, 2, 1e-4
    pwm = synth.random_pwm(W, 0, 41)
    cols = np.stack([random_codes(rng, W, S, gap_rate) for _ in range(n)])
    wts = rng.random(n)
    lam = np.log(rng.dirichlet(np.ones(4))) + 0.3
    desc, keep = _abi.make_desc(_abi.MODEL_FG, W, 0, tree, np.tile(tree.branch, (W, 1)), pwm)
    vals = np.zeros(5)
    rc = emul.emul_fe_values(C.byref(desc), cols.ctypes.data_as(C.POINTER(C.c_uint8)), n, wts.ctypes.data_as(C.POINTER(C.c_double)),
                             pos, lam.ctypes.data_as(C.POINTER(C.c_double)), eps, vals.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    fs = orc.FESet(oracle_fg(newick, pwm), cols, wts)
    f0, g = fs.grad(pos, lam, eps)
    assert abs(vals[0] - f0) <= 1e-9 * abs(f0)
    got_g = (vals[1:] - vals[0]) / eps
    noise = 64 * np.finfo(float).eps * abs(f0) / eps
    assert np.all(np.abs(got_g - g) <= 1e-9 * np.abs(g) + noise), (got_g, g)


WAVE_TREES = [synth.caterpillar_newick(5), "((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.4):0.25):0.1,F:0.3)",
              "((A:0.1,B:0.2):0.1,(C:0.3,D:0.1):0.2,(E:0.2,F:0.25):0.15)", synth.star_newick(4), synth.random_binary_newick(9, 2)]


@pytest.mark.parametrize("newick", WAVE_TREES)
@pytest.mark.parametrize("gap_rate", [0.0, 0.2])
def test_wavefront_trajectories_equal_the_sequential_pass(emul, newick, gap_rate):
    """The algorithm of the wavefront trajectory kernel (csrc/memo.cuh) modelled on the host: node i of sweep k at step
    2 k + depth(i), message slots per child, terms summed per sweep in node order, early stop at a bitwise fixed point or
    two-cycle with the periodic fill -- every trajectory entry equals the sequential pass bit for bit."""
    tree = parse_newick(newick)
    rng = np.random.default_rng(11)
    W, n, K1 = 3, 60, 101
    pwm = synth.random_pwm(W, 0, 9)
    cols = random_codes(rng, n, tree.n_species, gap_rate)
    desc, keep = _abi.make_desc(_abi.MODEL_FG, W, 0, tree, np.tile(tree.branch, (W, 1)), pwm, _abi.EVOL_F81ALPHA, _abi.MODE_MEANFIELD)
    dp = C.POINTER(C.c_double)
    seq = np.zeros((n, W, K1))
    assert emul.emul_seq_traj(C.byref(desc), cols.ctypes.data_as(C.POINTER(C.c_uint8)), n, K1, seq.ctypes.data_as(dp)) == 0
    for early in (0, 1):
        wave = np.zeros((n, W, K1))
        stopped = np.zeros((n, W), np.int32)
        rc = emul.emul_wave_traj(C.byref(desc), cols.ctypes.data_as(C.POINTER(C.c_uint8)), n, K1, early, wave.ctypes.data_as(dp),
                                 stopped.ctypes.data_as(C.POINTER(C.c_int)))
        assert rc == 0
        assert np.array_equal(wave.view(np.uint64), seq.view(np.uint64)), early
        if early:
            assert (stopped < K1 - 1).mean() > 0.5            # most trees become periodic long before the last sweep
        else:
            assert (stopped == K1 - 1).all()


@pytest.mark.parametrize("newick", [synth.caterpillar_newick(5), synth.random_binary_newick(12, 3), synth.random_binary_newick(30, 5),
                                    synth.star_newick(6), "((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.4):0.25):0.1,F:0.3)"])
@pytest.mark.parametrize("W", [1, 2, 12])
def test_order1_wavefront_schedule_respects_the_reference_sweep(emul, newick, W):
    """The list schedule of the order-1 wavefront kernel (csrc/model_host.hpp::wave_program): every node update (t, i) runs
    exactly once per sweep, at most four per step, AFTER the nodes whose new values the reference's Gauss-Seidel order gives
    it -- (t, parent), (t-1, i), (t-1, children) -- and BEFORE the nodes whose old values it reads -- (t, children), (t+1, i),
    (t+1, parent).  Any such schedule reproduces the reference's sweep bit for bit (csrc/net1w.cuh)."""
    tree = parse_newick(newick)
    pwm = synth.random_pwm(W, 1, 3)
    desc, keep = _abi.make_desc(_abi.MODEL_FG, W, 1, tree, np.tile(tree.branch, (W, 1)), pwm, _abi.EVOL_F81ALPHA, _abi.MODE_MEANFIELD)
    cap = 4 * W * tree.n_nodes + 64
    sched = np.zeros(cap, np.int32)
    par = np.zeros(tree.n_nodes, np.int32)
    ns, ni = C.c_int(), C.c_int()
    ip = C.POINTER(C.c_int)
    assert emul.emul_wave_schedule(C.byref(desc), sched.ctypes.data_as(ip), cap, C.byref(ns), C.byref(ni), par.ctypes.data_as(ip)) == 0
    I, n_steps = ni.value, ns.value
    par = par[:I]
    sched = sched[:4 * n_steps].reshape(n_steps, 4)
    step_of = {}
    for s in range(n_steps):
        live = [int(c) for c in sched[s] if c >= 0]
        assert 1 <= len(live) <= 4
        for c in live:
            key = (c >> 16, c & 0xffff)
            assert key not in step_of and key[0] < W and key[1] < I
            step_of[key] = s
    assert len(step_of) == W * I                                       # every node update once per sweep
    kids = {i: [c for c in range(1, I) if par[c] == i] for i in range(I)}
    for (t, i), s in step_of.items():
        new = ([(t, int(par[i]))] if i > 0 else []) + ([(t - 1, i)] + [(t - 1, c) for c in kids[i]] if t > 0 else [])
        old = [(t, c) for c in kids[i]] + ([(t + 1, i)] + ([(t + 1, int(par[i]))] if i > 0 else []) if t + 1 < W else [])
        assert all(step_of[k] < s for k in new), (t, i)
        assert all(step_of[k] > s for k in old), (t, i)
    # the chain is much shorter than the W * I sequential updates once the tree is not a path
    if I >= 4 and W == 12:
        assert n_steps <= 0.6 * W * I


@pytest.mark.parametrize("newick", [synth.caterpillar_newick(5), "((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.4):0.25):0.1,F:0.3)"])
@pytest.mark.parametrize("gap_rate", [0.0, 0.15])
def test_pattern_table_decomposition(emul, newick, gap_rate):
    """What the pattern-table path rests on (csrc/memo.cuh): for an order-0 model the free energy of a window after k sweeps is
    the tree-order sum of per-(column, position) trajectories, and the reference's stop rule applied to that sum gives the
    window's score and sweep count -- bit for bit what the window-by-window pass computes.  Also with a sweep cap: windows whose
    rule has not fired within the cap are exactly the ones the direct kernel has to take."""
    tree = parse_newick(newick)
    S = tree.n_species
    rng = np.random.default_rng(17)
    W, L, K1 = 4, 120, 101
    pwm = synth.random_pwm(W, 0, 21)
    codes = random_codes(rng, L, S, gap_rate)
    ref, ref_k = run_emul(emul, newick, pwm, codes, 0, _abi.MODE_MEANFIELD)
    desc, keep = _abi.make_desc(_abi.MODEL_FG, W, 0, tree, np.tile(tree.branch, (W, 1)), pwm, _abi.EVOL_F81ALPHA, _abi.MODE_MEANFIELD)
    traj = np.zeros((L, W, K1))
    assert emul.emul_seq_traj(C.byref(desc), codes.ctypes.data_as(C.POINTER(C.c_uint8)), L, K1, traj.ctypes.data_as(C.POINTER(C.c_double))) == 0
    cap = 12
    pending = 0
    for j in range(L - W + 1):
        def F(k):
            s = 0.0
            for t in range(W):
                s = s + traj[j + t, t, k]                   # tree-major, like calcFreeEnergy(0, W-1)
            return s
        old, k, conv = F(0), 0, False
        cur = old
        while (k < 100 and not conv) or k < 2:              # MeanFieldForBayesNet.java:107, :204-207
            cur = F(k + 1)
            if abs(old - cur) < 1e-5:
                conv = True
            old = cur
            k += 1
        assert k == ref_k[j] and -cur == ref[j], j
        pending += k > cap
    assert pending == int((ref_k > cap).sum())


# ---- order-1 network, node-per-thread kernel (csrc/net1n.cuh): its update function under its own program, on the host ----
@pytest.fixture(scope="module")
def emul_o1n(emul):
    emul.emul_score_o1n.argtypes = [C.POINTER(_abi.ModelDesc), C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int,
                                    C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return emul


O1N_TREES = {"star5": synth.star_newick(5), "cat5": synth.caterpillar_newick(5), "rand12": synth.random_binary_newick(12, 3),
             "mixed6": "((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.4):0.25):0.1,F:0.3)",
             "c3": synth.config_newick(synth.CONFIGS["c3"])}


@pytest.mark.parametrize("tree,W", [("star5", 1), ("star5", 6), ("cat5", 2), ("cat5", 7), ("mixed6", 5), ("rand12", 3), ("c3", 12)])
def test_order1_node_update_matches_oracle(emul_o1n, tree, W):
    """n1n_update under node_program's schedule and records = the reference's sweep: scores within 1e-9 relative and the
    same sweep counts as the oracle; the emulation also runs the node slots of every step in both orders and requires
    identical state (no node of a step reads what another node of the step writes)."""
    from util import oracle_fg
    newick = O1N_TREES[tree]
    tr = parse_newick(newick)
    rng = np.random.default_rng(5 + W)
    pwm = synth.random_pwm(W, 1, 61 + W)
    codes = random_codes(rng, W + 9, tr.n_species)
    desc, keep = _abi.make_desc(_abi.MODEL_FG, W, 1, tr, np.tile(tr.branch, (W, 1)), pwm, _abi.EVOL_F81ALPHA, _abi.MODE_MEANFIELD)
    om = oracle_fg(newick, pwm, order=1)
    n = codes.shape[0] - W + 1
    for strand in (0, 1):
        out, sw, ns = np.zeros(n), np.zeros(n, np.int32), C.c_int()
        rc = emul_o1n.emul_score_o1n(C.byref(desc), codes.ctypes.data_as(C.POINTER(C.c_uint8)), codes.shape[0], strand, (W + strand) % 8,
                                     out.ctypes.data_as(C.POINTER(C.c_double)), sw.ctypes.data_as(C.POINTER(C.c_int)), C.byref(ns))
        assert rc == 0
        ref, rsw = om.score_windows(codes, bool(strand))
        assert np.array_equal(sw, rsw)
        assert rel_err(out, ref) < 1e-9
        n_int = int((np.bincount(tr.parent[1:], minlength=tr.n_nodes) > 0).sum())
        assert ns.value >= -(-W * n_int // 4)            # at most four nodes per step


# ---- order 0, F81-structured pass (csrc/tree_pass_f81.cuh): compact tables, free energy without a logarithm per node ----
@pytest.mark.parametrize("tree", ["star5", "cat5", "mixed6", "rand12", "c3"])
@pytest.mark.parametrize("evol", [_abi.EVOL_F81ALPHA, _abi.EVOL_F81BETA])
@pytest.mark.parametrize("gap_rate", [0.0, 0.1])
def test_order0_f81_pass_matches_oracle(emul, tree, evol, gap_rate):
    from util import oracle_fg, orc
    emul.emul_score_o0f.argtypes = [C.POINTER(_abi.ModelDesc), C.POINTER(C.c_uint8), C.c_int, C.c_int, C.POINTER(C.c_double),
                                    C.POINTER(C.c_int)]
    newick = O1N_TREES[tree]
    tr = parse_newick(newick)
    rng = np.random.default_rng(17)
    W = 5
    pwm = synth.random_pwm(W, 0, 63)
    codes = random_codes(rng, W + 12, tr.n_species, gap_rate)
    desc, keep = _abi.make_desc(_abi.MODEL_FG, W, 0, tr, np.tile(tr.branch, (W, 1)), pwm, evol, _abi.MODE_MEANFIELD)
    om = oracle_fg(newick, pwm, order=0, evol=orc.F81ALPHA if evol == _abi.EVOL_F81ALPHA else orc.F81BETA)
    n = codes.shape[0] - W + 1
    for strand in (0, 1):
        out, sw = np.zeros(n), np.zeros(n, np.int32)
        rc = emul.emul_score_o0f(C.byref(desc), codes.ctypes.data_as(C.POINTER(C.c_uint8)), codes.shape[0], strand,
                                 out.ctypes.data_as(C.POINTER(C.c_double)), sw.ctypes.data_as(C.POINTER(C.c_int)))
        assert rc == 0
        ref, rsw = om.score_windows(codes, bool(strand))
        assert np.array_equal(sw, rsw)
        assert rel_err(out, ref) < 1e-9


@pytest.mark.parametrize("order", [1, 2])
def test_alignment_based_expansion_matches_the_java_statement_by_statement(emul, order):
    """csrc/ab_expand.hpp (what ab_windows_order_kernel / ab_columns_order_kernel / the count kernels run per row and column)
    against oracle.ab_motif_ln_cond_prob / ab_bg_ln_cond_prob, the literal restatements of AlignmentBasedModel.java:187-228 and
    AlignmentBasedBGModel.java:172-222 -- every gap pattern of a context, including the background's stale list entries when
    two predecessors are gaps."""
    import itertools
    rng = np.random.default_rng(order)
    W = order + 2
    cond = [rng.dirichlet(np.ones(4), size=4 ** min(p, order)).ravel() for p in range(W)]
    bg = cond[:order + 1]
    for pattern in itertools.product([0, 1, 2, 3, 4], repeat=order + 1):
        row = np.array(list(pattern[::-1]) + [1], np.uint8)          # row[l - k] = pattern[k] at l = order
        l = order
        rp = row.ctypes.data_as(C.POINTER(C.c_uint8))
        for pos_mot in range(order + 1):                              # motif position: context length min(order, pos_mot)
            got = emul.emul_ab_term(cond[pos_mot].ctypes.data_as(C.POINTER(C.c_double)), order, rp, l, pos_mot, 0)
            ref = orc.ab_motif_ln_cond_prob(cond, order, row, l, pos_mot)
            assert got == ref or abs(got - ref) <= 1e-15 * abs(ref), (pattern, pos_mot, got, ref)
        got = emul.emul_ab_term(bg[order].ctypes.data_as(C.POINTER(C.c_double)), order, rp, l, l, 1)
        ref = orc.ab_bg_ln_cond_prob(bg, order, row, l)
        assert got == ref or abs(got - ref) <= 1e-15 * abs(ref), (pattern, got, ref)
    # the background at the start of an alignment (l < order): shorter contexts, table min(l, order)
    row = np.array([4, 2, 4, 1], np.uint8)
    for l in range(order):
        got = emul.emul_ab_term(bg[l].ctypes.data_as(C.POINTER(C.c_double)), order, row.ctypes.data_as(C.POINTER(C.c_uint8)), l, l, 1)
        assert abs(got - orc.ab_bg_ln_cond_prob(bg, order, row, l)) <= 1e-15


# ---- generic net (csrc/netg.cuh): backgrounds of any order, incl. the stale trees of the short flank ranges ----------------
def _netg_items(emul, desc, codes, cols, lens):
    emul.emul_netg_items.argtypes = [C.POINTER(_abi.ModelDesc), C.POINTER(C.c_uint8), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    cols = np.ascontiguousarray(cols, np.int32); lens = np.ascontiguousarray(lens, np.int32)
    n = cols.shape[0]
    f, l, sw = np.zeros(n), np.zeros(n), np.zeros(n, np.int32)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    assert emul.emul_netg_items(C.byref(desc), codes.ctypes.data_as(C.POINTER(C.c_uint8)), n, cols.ctypes.data_as(ip),
                                lens.ctypes.data_as(ip), f.ctypes.data_as(dp), l.ctypes.data_as(dp), sw.ctypes.data_as(ip)) == 0
    return f, l, sw


def flank_items(L, W, k):
    """(left, right) items of every offset j: the column each of the k + 1 trees observes and the number of leading trees the
    range itself covers -- the host statement of bgk_items_flank_kernel (csrc/phyfoo.cu)."""
    left, right = [], []
    for j in range(L - W + 1):
        s, e = max(j - k, 0), min(j + W + k - 1, L - 1)
        n1, n2 = j - s, e - (j + W) + 1
        after_full = [e - k + t for t in range(k + 1)]
        after_left = [s + t if (n1 > 0 and t < n1) else after_full[t] for t in range(k + 1)]
        left.append((after_left, max(n1, 1)))
        right.append(([j + W + t if t < n2 else after_left[t] for t in range(k + 1)], max(n2, 1)))
    return left, right


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("gap_rate", [0.0, 0.12])
def test_generic_net_background_of_any_order_matches_oracle(emul, order, gap_rate):
    from util import oracle_bg
    rng = np.random.default_rng(40 + order)
    newick = TREES[3]
    tree = parse_newick(newick)
    T = order + 1
    pis = [rng.dirichlet(np.ones(4), size=4 ** min(order, t)) for t in range(T)]
    desc, keep = _abi.make_desc(_abi.MODEL_BG, 0, order, tree, np.tile(tree.branch, (T, 1)), pis)
    ob = oracle_bg(newick, order, pis)
    L, W = 13, 4
    codes = random_codes(rng, L, tree.n_species, gap_rate)
    first, last, sw = _netg_items(emul, desc, codes, [[u + t for t in range(T)] for u in range(L - order)], [T] * (L - order))
    full = lambda s, e: -(first[s] + last[s:e - order + 1].sum())                            # noqa: E731
    for s, e in [(0, L - 1), (2, 9), (3, 3 + order), (5, L - 1)]:
        ref = ob.bg_range(codes, s, e)
        assert abs(full(s, e) - ref) <= 1e-12 * abs(ref)
    # the three calls of every offset in the reference's order on ONE model object (SingleHiddenMotifMixture.java:536-540): the
    # short flank ranges see the trees the calls before them left behind
    left, right = flank_items(L, W, order)
    lf, _, _ = _netg_items(emul, desc, codes, [c for c, _ in left], [n for _, n in left])
    rf, _, _ = _netg_items(emul, desc, codes, [c for c, _ in right], [n for _, n in right])
    ob.bg_range(codes, 0, L - 1)                                                             # logPSeq, :526
    for j in range(L - W + 1):
        s, e = max(j - order, 0), min(j + W + order - 1, L - 1)
        r_full, r_left, r_right = ob.bg_range(codes, s, e), ob.bg_range(codes, s, j - 1), ob.bg_range(codes, j + W, e)
        assert abs(full(s, e) - r_full) <= 1e-12 * abs(r_full)
        if j > 0:
            assert abs(-lf[j] - r_left) <= 1e-12 * abs(r_left), (j, -lf[j], r_left)
        else:
            assert r_left == 0.0
        if j + W <= L - 1:
            assert abs(-rf[j] - r_right) <= 1e-12 * abs(r_right), (j, -rf[j], r_right)
        else:
            assert r_right == 0.0


@pytest.mark.parametrize("order", [1, 2])
def test_generic_net_expected_counts_give_the_fixed_q_objective(emul, order):
    """netg_counts: with q fixed, sum_e C[e] log table_theta[e] - ENT is the oracle's fixed-q objective for ANY parameter set
    (new stationary distributions at a position with 4^min(t, k) context rows, new branch lengths)."""
    from util import oracle_bg
    rng = np.random.default_rng(70 + order)
    newick = TREES[3]
    tree = parse_newick(newick)
    T = order + 1
    pis = [rng.dirichlet(np.ones(4), size=4 ** min(order, t)) for t in range(T)]
    branch = np.tile(tree.branch, (T, 1))
    desc, keep = _abi.make_desc(_abi.MODEL_BG, 0, order, tree, branch, pis)
    L = 20
    codes = random_codes(rng, L, tree.n_species, 0.1)
    n = L - order
    cols = np.ascontiguousarray([[u + t for t in range(T)] for u in range(n)], np.int32)
    w = rng.uniform(0.5, 2.0, n)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    emul.emul_netg_tab.argtypes = [C.POINTER(_abi.ModelDesc), dp, C.c_int]
    emul.emul_netg_counts.argtypes = [C.POINTER(_abi.ModelDesc), C.POINTER(C.c_uint8), C.c_int, ip, dp, dp, dp]
    ne = emul.emul_netg_tab(C.byref(desc), None, 0)
    counts, ent = np.zeros(ne), C.c_double()
    assert emul.emul_netg_counts(C.byref(desc), codes.ctypes.data_as(C.POINTER(C.c_uint8)), n, cols.ctypes.data_as(ip),
                                 w.ctypes.data_as(dp), counts.ctypes.data_as(dp), C.byref(ent)) == 0

    def objective(pis_, branch_):
        d2, k2 = _abi.make_desc(_abi.MODEL_BG, 0, order, tree, branch_, pis_)
        tab = np.zeros(ne)
        assert emul.emul_netg_tab(C.byref(d2), tab.ctypes.data_as(dp), ne) == ne
        nz = counts != 0
        return float((counts[nz] * tab[nz]).sum() - ent.value)

    ob = oracle_bg(newick, order, pis)
    ofe = orc.FESet(ob, np.stack([codes[u:u + T] for u in range(n)]), w)
    ref = ofe.sum()
    assert abs(objective(pis, branch) - ref) <= 1e-12 * abs(ref)
    for t in range(T):                                         # new parameters at one position after the other
        lam = rng.normal(size=pis[t].shape)
        new = np.exp(lam) / np.exp(lam).sum(axis=1, keepdims=True)
        pis = [new if i == t else p for i, p in enumerate(pis)]
        ref = ofe.eval(t, lam.ravel())
        assert abs(objective(pis, branch) - ref) <= 1e-12 * abs(ref), (t, objective(pis, branch), ref)
    b2 = branch.copy()
    b2[T - 1, 1:] = rng.uniform(0.05, 0.5, tree.n_nodes - 1)
    for k in range(1, tree.n_nodes):
        ob.set_branch(T - 1, k, b2[T - 1, k])
    ref = ofe.sum()
    assert abs(objective(pis, b2) - ref) <= 1e-12 * abs(ref)


# ---- the kernels' arithmetic against outputs of the reference's own statements (tests/golden/refexec_meanfield.npz) ----------
def test_kernel_arithmetic_reproduces_the_reference_statements(emul_o1n):
    """The per-thread code of the order-0 kernels (generic pass and F81-structured pass) and of the order-1 node kernel,
    compiled for the host, on the golden inputs of tests/test_reference_statements.py: scores within 1e-9 relative of what the
    reference's translated statements produced, sweep counts identical (what tests/test_zz_gpu_reference_statements.py asks of
    the library on the device)."""
    from test_reference_statements import cases, motif_pwm
    emul = emul_o1n
    emul.emul_score_o0f.argtypes = [C.POINTER(_abi.ModelDesc), C.POINTER(C.c_uint8), C.c_int, C.c_int, C.POINTER(C.c_double),
                                    C.POINTER(C.c_int)]
    checked = 0
    for c in cases("motif"):
        W, order, newick = int(c["W"]), int(c["order"]), str(c["newick"])
        tr = parse_newick(newick)
        pwm = motif_pwm(c)
        codes = np.ascontiguousarray(c["codes"], np.uint8)
        n = codes.shape[0] - W + 1
        desc, keep = _abi.make_desc(_abi.MODEL_FG, W, order, tr, np.tile(tr.branch, (W, 1)), pwm, _abi.EVOL_F81ALPHA, _abi.MODE_MEANFIELD)
        for strand in (0, 1):
            ref, rsw = c["score"][strand], c["sweeps"][strand]
            results = []
            if order == 0:
                results.append(run_emul(emul, newick, pwm, codes, strand, _abi.MODE_MEANFIELD))
                out, sw = np.zeros(n), np.zeros(n, np.int32)
                assert emul.emul_score_o0f(C.byref(desc), codes.ctypes.data_as(C.POINTER(C.c_uint8)), codes.shape[0], strand,
                                           out.ctypes.data_as(C.POINTER(C.c_double)), sw.ctypes.data_as(C.POINTER(C.c_int))) == 0
                results.append((out, sw))
            elif not (codes == 4).any():                       # the node kernel takes gap-free windows (the rest go to net1w.cuh)
                out, sw, ns = np.zeros(n), np.zeros(n, np.int32), C.c_int()
                assert emul.emul_score_o1n(C.byref(desc), codes.ctypes.data_as(C.POINTER(C.c_uint8)), codes.shape[0], strand, 3,
                                           out.ctypes.data_as(C.POINTER(C.c_double)), sw.ctypes.data_as(C.POINTER(C.c_int)), C.byref(ns)) == 0
                results.append((out, sw))
            for out, sw in results:
                assert np.array_equal(sw, rsw), (str(c["tree"]), order, strand)
                assert rel_err(out, ref) < 1e-9
                checked += out.size
    assert checked >= 2 * 88 + 40
