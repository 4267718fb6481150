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
