"""Host side of the M-step without a device: the sufficient-statistics objective (phyfoo_b200/optimize.py) on counts handed
in by a stand-in for the device pass, the vectorised substitution tables, and the BFGS stand-in on that objective."""
import numpy as np
import pytest

from phyfoo_b200 import _abi, evolution, optimize
from phyfoo_b200.newick import parse_newick


class FakeFE:
    """What core.FESet.suffstats returns: expected counts per table entry [W][E] and the entropy term."""

    def __init__(self, counts, entropy=0.0):
        self.counts, self.entropy = counts, entropy

    def suffstats(self, W, E):
        assert self.counts.shape == (W, E)
        return self.counts.copy(), self.entropy


@pytest.mark.parametrize("evol", [_abi.EVOL_F81ALPHA, _abi.EVOL_F81BETA])
def test_tables_entry_for_entry(evol):
    rng = np.random.default_rng(2)
    for _ in range(20):
        pi = rng.dirichlet(np.ones(4))
        br = np.concatenate([[0.0], rng.uniform(1e-4, 0.95, 10)])
        got = evolution.tables(pi, br, evol)
        ref = np.concatenate([pi] + [evolution.transition(pi, br[k], evol).ravel() for k in range(1, br.size)])
        assert np.array_equal(got, ref)
        T = got[4:].reshape(-1, 4, 4)
        assert np.allclose(T.sum(axis=2), 1.0)                       # every row of every table is a distribution


def test_objective_is_the_dot_product_and_ignores_empty_entries():
    tree = parse_newick("((A:0.1,B:0.3):0.15,(C:0.2,D:0.4):0.1)")
    W, N = 3, tree.n_nodes
    E = 4 + 16 * (N - 1)
    rng = np.random.default_rng(4)
    counts = rng.uniform(0, 5, (W, E))
    counts[:, rng.integers(0, E, 20)] = 0.0                            # entries no window ever used
    pi = [rng.dirichlet(np.ones(4)) for _ in range(W)]
    pi[1] = np.array([0.5, 0.5, 0.0, 0.0])                             # a zero probability: log table has -inf entries
    counts[1, [2, 3]] = 0.0                                            # ... which carry no count at the root
    nodes_23 = [4 + 16 * k + a * 4 + b for k in range(N - 1) for a in range(4) for b in (2, 3)]
    counts[1, nodes_23] = 0.0
    branch = np.tile(tree.branch, (W, 1))
    obj = optimize.SuffStatObjective(FakeFE(counts, 1.25), pi, branch, _abi.EVOL_F81ALPHA)
    want = -1.25
    for t in range(W):
        with np.errstate(divide="ignore"):
            lt = np.log(evolution.tables(pi[t], branch[t]))
        want += sum(c * l for c, l in zip(counts[t], lt) if c != 0.0)
    assert np.isfinite(obj.value()) and abs(obj.value() - want) < 1e-9 * abs(want)
    # counts replaced after construction (the all-reduced ones of a multi-GPU run) are picked up
    obj.counts = 2.0 * counts
    for t in range(W):
        obj.set_position(t)
    assert abs(obj.value() - (2.0 * (want + 1.25) - 1.25)) < 1e-9 * abs(want)


def test_bfgs_on_sufficient_statistics_finds_the_closed_form_optimum():
    """With branch length 1 under F81-alpha every node is an independent draw from pi, so the optimum of
    sum_e C[e] log table[e] is the normalised column sum of the counts."""
    tree = parse_newick("(A:1.0,B:1.0,C:1.0)")
    N = tree.n_nodes
    E = 4 + 16 * (N - 1)
    rng = np.random.default_rng(7)
    counts = rng.uniform(1, 10, (1, E))
    obj = optimize.SuffStatObjective(FakeFE(counts), [np.full(4, 0.25)], np.tile(tree.branch, (1, 1)), _abi.EVOL_F81ALPHA)
    per_symbol = counts[0, :4] + counts[0, 4:].reshape(-1, 4, 4).sum(axis=(0, 1))
    want = per_symbol / per_symbol.sum()
    fun = lambda z: -obj.eval_lambda(0, z)                             # noqa: E731
    x, f, it, ev = optimize.quasi_newton_bfgs(fun, lambda z: obj.grad(fun, z), np.log(np.full(4, 0.25)),
                                              optimize.TC_MOTIF_SEPARATED_POSITIONS())
    obj.eval_lambda(0, x)
    assert np.abs(obj.pi[0] - want).max() < 2e-3 and it >= 1 and ev > 5
    assert obj.value() > -fun(np.log(np.full(4, 0.25))) - 1e-12        # the objective did not get worse


# ---- the sufficient-statistics objective and its optimiser inside libphyfoo.so (csrc/suffstat_host.hpp): host code ----
def _ss_case(seed=3, W=5):
    from phyfoo_b200 import _abi, synth
    from phyfoo_b200.newick import parse_newick
    tree = parse_newick(synth.random_binary_newick(6, 9))
    rng = np.random.default_rng(seed)
    pwm = [rng.dirichlet(np.ones(4), size=1) for _ in range(W)]
    branch = np.tile(tree.branch, (W, 1))
    E = 4 + 16 * (tree.n_nodes - 1)
    counts = rng.gamma(2.0, 3.0, size=(W, E)) * (rng.random((W, E)) > 0.1)
    desc, keep = _abi.make_desc(_abi.MODEL_FG, W, 0, tree, branch, pwm, _abi.EVOL_F81ALPHA, _abi.MODE_MEANFIELD)
    return tree, pwm, branch, counts, desc, keep


class _FakeFE:
    def __init__(self, counts, entropy):
        self.c, self.e = counts, entropy

    def suffstats(self, W, E):
        return self.c.copy(), self.e


def test_library_suffstat_objective_matches_numpy_objective():
    from phyfoo_b200 import _abi, core, optimize
    tree, pwm, branch, counts, desc, keep = _ss_case()
    ss = core.SSObjective(desc=desc, counts=counts, entropy=1.25)
    ref = optimize.SuffStatObjective(_FakeFE(counts, 1.25), pwm, branch, _abi.EVOL_F81ALPHA)
    assert abs(ss.value() - ref.value()) <= 1e-13 * abs(ref.value())
    rng = np.random.default_rng(1)
    for t in (0, 3, 4):
        lam = rng.normal(size=4)
        f, g = ss.grad(t, lam)
        f_ref, g_ref = ref.grad(lambda z: ref.eval_lambda(t, z), lam)
        ref.eval_lambda(t, lam)                          # (the numpy gradient leaves the object at its last perturbed point)
        assert abs(f - f_ref) <= 1e-13 * abs(f_ref)
        assert np.max(np.abs(g - g_ref)) <= 1e-7 * np.max(np.abs(g_ref))      # both subtract sums of size |f| and divide by 1e-4
        assert np.allclose(ss.get_pi(t), ref.pi[t], rtol=1e-15, atol=0)
    b = np.concatenate([[0.0], rng.uniform(0.05, 0.4, tree.n_nodes - 1)])
    f = ss.eval_branches(2, b)
    ref.set_position(2, branch=b)
    assert abs(f - ref.value()) <= 1e-13 * abs(f)


def test_library_bfgs_matches_python_restatement(monkeypatch):
    """phyfoo_ss_optimize (C++, csrc/suffstat_host.hpp) against optimize.quasi_newton_bfgs (numpy) on the same objective: both
    restate jstacs Optimizer.quasiNewtonBFGS / findBracket / brentsMethod statement by statement, so they take the same
    iterates -- the same number of function evaluations and the same optimum to rounding."""
    from phyfoo_b200 import _abi, core, optimize
    monkeypatch.setenv("PHYFOO_SS_THREADS", "1")           # one position after the other, like the loop below
    tree, pwm, branch, counts, desc, keep = _ss_case(seed=11)
    W = len(pwm)
    ss = core.SSObjective(desc=desc, counts=counts, entropy=0.0)
    ref = optimize.SuffStatObjective(_FakeFE(counts, 0.0), pwm, branch, _abi.EVOL_F81ALPHA)
    order = np.array([2, 0, 4, 1, 3], np.int32)
    fb, fa, ev = ss.optimize(order)
    assert abs(fb - ref.value()) <= 1e-13 * abs(fb)
    n_ref = 0
    for pos in order:
        pos = int(pos)
        fun = lambda z: -ref.eval_lambda(pos, z)          # noqa: E731
        x, f, it, n = optimize.quasi_newton_bfgs(fun, lambda z: ref.grad(fun, z), np.log(ref.pi[pos]), optimize.TC_MOTIF_SEPARATED_POSITIONS())
        ref.eval_lambda(pos, x)
        n_ref += n + 1
    assert fa >= fb
    assert abs(fa - ref.value()) <= 1e-12 * abs(fa)
    assert ev == n_ref                                     # the same line searches, evaluation for evaluation
    for t in range(W):                                     # (the two objectives differ in the last bits, which the forward
        assert np.allclose(ss.get_pi(t), ref.pi[t], rtol=0, atol=1e-6)     # differences magnify by |f| / 1e-4)
    assert ev > 5 * W


def test_jstacs_line_search_pieces():
    """findBracket / brentsMethod / LimitedMedianStartDistance on functions with known answers."""
    from phyfoo_b200 import optimize
    f = lambda a: (a - 0.3) ** 2 + 1.0                     # noqa: E731
    br = optimize.find_bracket(f, 0.0, f(0.0), 0.1)       # downhill at the start: expand by the golden ratio
    assert br[0] < br[2] < br[4] and br[3] <= br[1] and br[3] <= br[5] and br[0] <= 0.3 <= br[4]
    x, fx = optimize.brents_method(f, br[0], br[2], br[3], br[4], 1e-3)
    assert abs(x - 0.3) < 1e-3 and abs(fx - 1.0) < 1e-6
    br = optimize.find_bracket(f, 0.0, f(0.0), 5.0)       # overshoots: shrink towards the lower end
    assert br[0] == 0.0 and br[2] < 5.0 and br[3] <= br[1] and br[0] <= 0.3 <= br[4]
    sdf = optimize.LimitedMedianStartDistance(10, 0.1)
    assert abs(sdf.get_new_start_distance() - 0.0667) < 1e-15
    for v in (1.0, 2.0, 3.0, 4.0, 5.0, 6.0):
        sdf.set_last_distance(v)
    assert abs(sdf.get_new_start_distance() - 0.667 * 0.5 * (1.0 + 2.0)) < 1e-15       # memory: 1..6 and four times 0.1
    # SmallStepCondition compares the SQUARED step length
    tc = optimize.CombinedCondition(1, small_step=1e-4)
    d = np.array([3.0, 4.0])
    assert tc.do_next_iteration(1, 0.0, 0.0, d, d, 0.0021) and not tc.do_next_iteration(1, 0.0, 0.0, d, d, 0.0019)
    # a quadratic bowl: BFGS finds the minimum
    A = np.array([[3.0, 1.0], [1.0, 2.0]]); b = np.array([1.0, -1.0])
    fun = lambda z: 0.5 * z @ A @ z - b @ z               # noqa: E731
    grad = lambda z: (fun(z), A @ z - b)                  # noqa: E731
    x, fmin, it, ev = optimize.quasi_newton_bfgs(fun, grad, np.zeros(2), optimize.CombinedCondition(1, small_gradient=1e-8))
    assert np.allclose(x, np.linalg.solve(A, b), atol=1e-5) and it >= 2


def test_library_mstep_on_host_threads_does_not_depend_on_their_number(monkeypatch):
    """phyfoo_ss_optimize runs distinct positions side by side (the objective is separable); every position is optimised against
    the other terms as they were on entry, so 2, 3 or 16 threads give the same bits, and the strictly sequential run
    (PHYFOO_SS_THREADS=1, each position seeing what the earlier ones left) the same optimum to rounding."""
    from phyfoo_b200 import core
    tree, pwm, branch, counts, desc, keep = _ss_case(seed=5)
    W = len(pwm)
    order = np.array([3, 1, 4, 0, 2], np.int32)
    res = {}
    for nt in ("1", "2", "3", "16"):
        monkeypatch.setenv("PHYFOO_SS_THREADS", nt)
        ss = core.SSObjective(desc=desc, counts=counts, entropy=0.5)
        fb, fa, ev = ss.optimize(order)
        res[nt] = (fb, fa, ev, np.stack([ss.get_pi(t) for t in range(W)]))
        ss.free()
    for nt in ("3", "16"):
        assert res[nt][1] == res["2"][1] and res[nt][2] == res["2"][2] and np.array_equal(res[nt][3], res["2"][3])
    assert res["1"][0] == res["2"][0] and res["2"][1] > res["2"][0]
    assert abs(res["1"][1] - res["2"][1]) <= 1e-9 * abs(res["1"][1])
    assert np.allclose(res["1"][3], res["2"][3], rtol=0, atol=1e-6)
    monkeypatch.delenv("PHYFOO_SS_THREADS")
    # a repeated position: one after the other
    ss = core.SSObjective(desc=desc, counts=counts, entropy=0.5)
    fb, fa, ev = ss.optimize(np.array([1, 1, 2], np.int32))
    assert fa > fb


def test_library_suffstat_objective_of_an_order1_network_on_the_host():
    """phyfoo_ss_create_host for an order-1 motif: per position 16 + 32 (N - 1) entries -- root [context][symbol], then per
    non-root node the off-diagonal log entries [context][symbol] and the diagonal ones (F81 alpha: log(alpha pi_c[b]) and
    log(1 - alpha + alpha pi_c[b])); the objective is the dot product with the counts, position t >= 1 has 16 parameters."""
    from phyfoo_b200 import _abi, core, synth
    from phyfoo_b200.newick import parse_newick
    tree = parse_newick("((A:0.1,B:0.3):0.15,(C:0.2,D:0.25):0.1)")
    N, W = tree.n_nodes, 3
    rng = np.random.default_rng(3)
    pwm = synth.random_pwm(W, 1, 5)
    branch = np.tile(tree.branch, (W, 1))
    E1 = 16 + 32 * (N - 1)
    counts = rng.uniform(0.0, 3.0, (W, E1))
    counts[0, 4:16] = 0.0                                           # position 0 has a single context row
    counts[0, 16:] = counts[0, 16:] * (np.arange(E1 - 16) % 16 < 4)
    desc, keep = _abi.make_desc(_abi.MODEL_FG, W, 1, tree, branch, pwm, _abi.EVOL_F81ALPHA, _abi.MODE_MEANFIELD)
    ss = core.SSObjective(desc=desc, counts=counts, entropy=2.5)

    def tab(t, pi):
        pi = np.asarray(pi).reshape(-1, 4)
        out = np.zeros(E1)
        out[:pi.size] = np.log(pi).ravel()
        for k in range(1, N):
            a = branch[t, k]
            for c in range(pi.shape[0]):
                out[16 + 32 * (k - 1) + 4 * c:16 + 32 * (k - 1) + 4 * c + 4] = np.log(a * pi[c])
                out[32 + 32 * (k - 1) + 4 * c:32 + 32 * (k - 1) + 4 * c + 4] = np.log(1 - a + a * pi[c])
        return out

    def objective(pis):
        return sum(float(counts[t] @ tab(t, pis[t])) for t in range(W)) - 2.5
    assert abs(ss.value() - objective(pwm)) <= 1e-12 * abs(ss.value())
    lam = rng.normal(size=16)
    new = np.exp(lam.reshape(4, 4)); new /= new.sum(axis=1, keepdims=True)
    f = ss.eval(2, lam)
    assert abs(f - objective([pwm[0], pwm[1], new])) <= 1e-12 * abs(f)
    assert np.allclose(ss.get_pi(2, 16), new.ravel(), rtol=1e-14, atol=0)
    with pytest.raises(ValueError):
        ss.eval(2, np.zeros(4))
    f0, g = ss.grad(0, np.log(np.asarray(pwm[0]).ravel()), 1e-4)
    assert g.size == 4 and np.isfinite(g).all()
    fb, fa, ev = ss.optimize(np.array([1, 0, 2], np.int32))
    assert fa >= fb and ev > 0


def test_library_suffstat_objective_under_hky_with_its_ratio():
    """phyfoo_ss_create_host for an order-0 motif under PHYFOO_EVOL_HKY_RATIO: the objective is the dot product of the counts with
    the logs of the general 4x4 tables of evolution/HKY.java:77-106 (beta = alpha * ratio), entry layout [pi (4)] then 16 per
    non-root node row-major [parent][own]; a new lambda re-builds the tables of that position; branch lengths likewise."""
    from phyfoo_b200 import _abi, core, synth
    from phyfoo_b200.newick import parse_newick
    tree = parse_newick("((A:0.1,B:0.3):0.15,(C:0.2,D:0.25):0.1)")
    N, W, ratio = tree.n_nodes, 2, 2.0
    rng = np.random.default_rng(4)
    pwm = synth.random_pwm(W, 0, 6)
    branch = np.tile(tree.branch, (W, 1))
    E = 4 + 16 * (N - 1)
    counts = rng.uniform(0.0, 3.0, (W, E))
    desc, keep = _abi.make_desc(_abi.MODEL_FG, W, 0, tree, branch, pwm, _abi.EVOL_HKY_RATIO, _abi.MODE_MEANFIELD, ratio)
    ss = core.SSObjective(desc=desc, counts=counts, entropy=1.5)

    def hky(pi, alpha):
        a, b = alpha * pi, alpha * ratio * pi
        return np.array([[1 - (a[2] + b[3] + b[1]), b[1], a[2], b[3]], [b[0], 1 - (a[3] + b[0] + b[2]), b[2], a[3]],
                         [a[0], b[1], 1 - (a[0] + b[3] + b[1]), b[3]], [b[0], a[1], b[2], 1 - (a[1] + b[0] + b[2])]])

    def tab(pi, br):
        pi = np.asarray(pi, np.float64).ravel()
        return np.log(np.concatenate([pi] + [hky(pi, br[k]).ravel() for k in range(1, N)]))

    def objective(pis, brs):
        return sum(float(counts[t] @ tab(pis[t], brs[t])) for t in range(W)) - 1.5
    assert abs(ss.value() - objective(pwm, branch)) <= 1e-12 * abs(ss.value())
    lam = rng.normal(size=4)
    new = np.exp(lam) / np.exp(lam).sum()
    f = ss.eval(1, lam)
    assert abs(f - objective([pwm[0], new], branch)) <= 1e-12 * abs(f)
    b2 = branch.copy()
    b2[1, 1:] = rng.uniform(0.05, 0.3, N - 1)
    f = ss.eval_branches(1, b2[1])
    assert abs(f - objective([pwm[0], new], b2)) <= 1e-12 * abs(f)
    # order 1 is refused for this evolution model (the order-1 kernels and their table layout assume F81 structure)
    d1, k1 = _abi.make_desc(_abi.MODEL_FG, W, 1, tree, branch, synth.random_pwm(W, 1, 6), _abi.EVOL_HKY_RATIO, _abi.MODE_MEANFIELD, ratio)
    with pytest.raises(Exception):
        core.SSObjective(desc=d1, counts=np.zeros((W, 16 + 32 * (N - 1))), entropy=0.0)


def test_numpy_tables_under_hky_with_its_ratio_equal_the_library():
    """evolution.tables / SuffStatObjective (numpy mirror used by trainEdges) under EVOL_HKY_RATIO against phyfoo_ss_* on the
    same counts: the objective at the start, after a new stationary distribution and after new branch lengths."""
    from phyfoo_b200 import _abi, core, evolution, optimize, synth
    from phyfoo_b200.newick import parse_newick
    tree = parse_newick("((A:0.1,B:0.3):0.15,(C:0.2,(D:0.05,E:0.2):0.25):0.1)")
    N, W, ratio = tree.n_nodes, 3, 0.5
    rng = np.random.default_rng(8)
    pwm = synth.random_pwm(W, 0, 9)
    branch = np.tile(tree.branch, (W, 1))
    E = 4 + 16 * (N - 1)
    counts = rng.uniform(0.0, 2.0, (W, E))
    T = evolution.transition(np.ravel(pwm[0]), 0.2, _abi.EVOL_HKY_RATIO, ratio)
    assert np.allclose(T.sum(axis=1), 1) and abs(T[0, 1] / T[0, 2] - ratio * np.ravel(pwm[0])[1] / np.ravel(pwm[0])[2]) < 1e-15

    class Counts:                                    # what core.FESet.suffstats hands over after the device pass
        def suffstats(self, w, e):
            return counts.copy(), 0.75

    obj = optimize.SuffStatObjective(Counts(), pwm, branch, _abi.EVOL_HKY_RATIO, ratio)
    desc, keep = _abi.make_desc(_abi.MODEL_FG, W, 0, tree, branch, pwm, _abi.EVOL_HKY_RATIO, _abi.MODE_MEANFIELD, ratio)
    ss = core.SSObjective(desc=desc, counts=counts, entropy=0.75)
    assert abs(obj.value() - ss.value()) <= 1e-12 * abs(ss.value())
    lam = rng.normal(size=4)
    new = np.exp(lam) / np.exp(lam).sum()
    f_lib = ss.eval(2, lam)
    obj.set_position(2, pi=new)
    f_np = sum(obj.position_term(t, obj.pi[t], obj.branch[t]) for t in range(W)) - 0.75
    assert abs(f_np - f_lib) <= 1e-12 * abs(f_lib)
    b = branch[1].copy()
    b[1:] = rng.uniform(0.05, 0.3, N - 1)
    f_lib = ss.eval_branches(1, b)
    obj.set_position(1, branch=b)
    f_np = sum(obj.position_term(t, obj.pi[t], obj.branch[t]) for t in range(W)) - 0.75
    assert abs(f_np - f_lib) <= 1e-12 * abs(f_lib)
