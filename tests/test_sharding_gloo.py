"""world_size-2 gloo test of the multi-GPU host logic: shard bounds, the sum all-reduce of the
E-step partials and the rank-order gather.  The per-rank compute is the oracle here (CPU box, test infrastructure);
on the GPUs the same plumbing carries the CUDA partials (bench.py --gpus N, tests/test_gpu_parity.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT, orc

from phyfoo_b200 import sharding, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _models(newick, pwm):
    fg = orc.Model(orc.FG, len(pwm), 0, newick)
    for t, p in enumerate(pwm):
        fg.set_pi(t, np.asarray(p).ravel())
    return fg, orc.Model(orc.BG, 0, 0, newick)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        newick = synth.caterpillar_newick(5)
        pwm = synth.random_pwm(6, 0, 3)
        smp, _ = synth.make_sample(newick, 11, 40, 5, pwm=pwm, length_jitter=12)
        assert sharding.world() == (rank, world)
        lo, hi = sharding.shard_bounds(smp.col_offsets, world)[rank]
        local = smp.subset(range(lo, hi))
        fg, bg = _models(newick, pwm)
        res = orc.estep(fg, bg, local.col_offsets, local.codes, np.log([0.7, 0.3]))
        partial = torch.tensor([res["ll"], res["w"][0], res["w"][1]], dtype=torch.float64)
        sharding.allreduce_sum_(partial)
        calls = sharding.gather_in_rank_order(res["argmax_off"])
        gamma = sharding.gather_in_rank_order(res["gamma"])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), partial=partial.numpy(), calls=calls, gamma=gamma, lo=lo, hi=hi)
    finally:
        dist.destroy_process_group()


def _order1_case(rank):
    """Expected counts of an order-1 motif's table entries as a rank would hold them (rank < 0: both ranks' sum)."""
    from phyfoo_b200.newick import parse_newick
    tree = parse_newick(synth.caterpillar_newick(4))
    W, E1 = 3, 16 + 32 * (tree.n_nodes - 1)
    pwm = synth.random_pwm(W, 1, 9)

    def part(r):
        c = np.random.default_rng(20 + r).uniform(0.0, 2.0, (W, E1))
        c[0, 4:16] = 0.0
        c[0, 16:] *= (np.arange(E1 - 16) % 16 < 4)
        return c
    if rank < 0:
        return pwm, part(0) + part(1), 0.25 * 3, np.array([2, 0, 1], np.int32)
    return pwm, part(rank), 0.25 * (rank + 1), np.array([2, 0, 1], np.int32)


def _order1_mstep(pwm, counts, entropy, order):
    from phyfoo_b200 import _abi, core
    from phyfoo_b200.newick import parse_newick
    tree = parse_newick(synth.caterpillar_newick(4))
    W = len(pwm)
    desc, keep = _abi.make_desc(_abi.MODEL_FG, W, 1, tree, np.tile(tree.branch, (W, 1)), pwm, _abi.EVOL_F81ALPHA, _abi.MODE_MEANFIELD)
    ss = core.SSObjective(desc=desc, counts=counts, entropy=float(entropy))
    ss.optimize(order)
    return np.concatenate([ss.get_pi(t, 4 if t == 0 else 16) for t in range(W)])


def _mstep_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        newick = synth.caterpillar_newick(5)
        W = 5
        pwm = synth.random_pwm(W, 0, 3)
        rng = np.random.default_rng(4)
        cols = rng.integers(0, 4, size=(21, W, 5)).astype(np.uint8)
        wts = rng.random(21)
        # selected windows are sharded like alignments: contiguous blocks
        lo, hi = sharding.shard_bounds(np.arange(22) * W, world)[rank]
        fg, _ = _models(newick, pwm)
        fe = orc.FESet(fg, cols[lo:hi], wts[lo:hi])
        lam = np.log([0.1, 0.2, 0.3, 0.4])
        f0, g = fe.grad(2, lam, 1e-4)
        vals = torch.tensor(np.concatenate([[f0], f0 + g * 1e-4]), dtype=torch.float64)      # [f, f_1..f_dim] of this shard
        sharding.allreduce_sum_(vals)
        counts = np.full((W, 3), float(rank + 1))
        c, e = sharding.allreduce_counts_(counts, 0.5 * (rank + 1))
        # the sharded M-step of an order-1 motif (models.PhyloBayesModel.train_selected, sharded=True): every rank's expected
        # counts summed, then the same host optimisation on every rank -- no further collective, identical parameters
        pi1, c1, e1, order1 = _order1_case(rank)
        c1, e1 = sharding.allreduce_counts_(c1, e1)
        np.savez(os.path.join(out_dir, f"m{rank}.npz"), vals=vals.numpy(), c=c, e=e, pi1=_order1_mstep(pi1, c1, e1, order1))
    finally:
        dist.destroy_process_group()


def test_two_rank_mstep_values_allreduce(tmp_path):
    """[f, f_1..f_dim] per shard summed over ranks gives the full-sample forward-difference gradient."""
    world = 2
    mp.spawn(_mstep_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    newick = synth.caterpillar_newick(5)
    W = 5
    pwm = synth.random_pwm(W, 0, 3)
    rng = np.random.default_rng(4)
    cols = rng.integers(0, 4, size=(21, W, 5)).astype(np.uint8)
    wts = rng.random(21)
    fg, _ = _models(newick, pwm)
    f0, g = orc.FESet(fg, cols, wts).grad(2, np.log([0.1, 0.2, 0.3, 0.4]), 1e-4)
    for r in range(world):
        z = np.load(tmp_path / f"m{r}.npz")
        v = z["vals"]
        assert abs(v[0] - f0) <= 1e-12 * abs(f0)
        assert np.max(np.abs((v[1:] - v[0]) / 1e-4 - g)) <= 1e-9 * np.max(np.abs(g)) + 64 * np.finfo(float).eps * abs(f0) / 1e-4
        assert np.array_equal(z["c"], np.full((W, 3), 3.0)) and float(z["e"]) == 1.5
    pi1, c1, e1, order1 = _order1_case(-1)
    ref = _order1_mstep(pi1, c1, e1, order1)
    a, b = np.load(tmp_path / "m0.npz")["pi1"], np.load(tmp_path / "m1.npz")["pi1"]
    assert np.array_equal(a, b) and np.array_equal(a, ref)       # same counts on every rank, same optimiser: same bits


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(0)
    lens = rng.integers(300, 401, 750)
    off = np.concatenate([[0], np.cumsum(lens)])
    for world in (1, 2, 3, 4, 8):
        b = sharding.shard_bounds(off, world)
        assert b[0][0] == 0 and b[-1][1] == 750 and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        cols = [off[hi] - off[lo] for lo, hi in b]
        assert max(cols) - min(cols) <= 2 * 400
    # more ranks than alignments: trailing ranks get empty shards, nothing is lost
    b = sharding.shard_bounds(np.array([0, 10, 20]), 4)
    assert sum(hi - lo for lo, hi in b) == 2 and all(hi >= lo for lo, hi in b)


def test_two_rank_estep_allreduce_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    newick = synth.caterpillar_newick(5)
    pwm = synth.random_pwm(6, 0, 3)
    smp, _ = synth.make_sample(newick, 11, 40, 5, pwm=pwm, length_jitter=12)
    fg, bg = _models(newick, pwm)
    ref = orc.estep(fg, bg, smp.col_offsets, smp.codes, np.log([0.7, 0.3]))
    want = np.array([ref["ll"], ref["w"][0], ref["w"][1]])
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in range(world))
    assert r0["hi"] == r1["lo"] and r0["lo"] == 0 and r1["hi"] == 11
    for r in (r0, r1):
        assert np.max(np.abs(r["partial"] - want) / np.abs(want)) < 1e-12      # same sums, re-associated over ranks
        assert np.array_equal(r["calls"], ref["argmax_off"])                   # bit-exact site calls in rank order
        assert np.array_equal(r["gamma"], ref["gamma"])
    assert np.array_equal(r0["partial"], r1["partial"])                        # every rank holds the same reduced values
