"""bench.py's host-side bookkeeping without a device: the roofline object for each scoring path from a list of kernel times
(the part of the bench line that cannot be exercised on the CPU otherwise), the flop models, the traffic lookup."""
import json
import os
import sys
from types import SimpleNamespace

import pytest

from util import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _run(name, order, S, I, n_windows=1000, patterns=100):
    cfg = dict(W=12, S=S, order=order, n_aln=10, length=100, tree="random")
    return SimpleNamespace(cfg=cfg, I=I, n_windows=n_windows, name=name, ds=SimpleNamespace(patterns=lambda: patterns))


@pytest.mark.parametrize("path,kernels,order,expect", [
    (0, {(0, "score_o0j_kernel"): 1.0, (1, "score_o0j_kernel"): 20.0, (2, "zoops_kernel"): 0.1}, 0, "score_o0j_kernel"),
    (0, {(0, "score_o0f_kernel"): 1.0, (1, "score_o0f_kernel"): 20.0, (2, "zoops_kernel"): 0.1}, 0, "score_o0f_kernel"),
    (0, {(0, "score_o0_kernel"): 1.0, (1, "score_o0_kernel"): 20.0}, 0, "score_o0_kernel"),
    (2, {(0, "score_o0j_kernel"): 1.0, (1, "score_o1n_kernel"): 50.0, (2, "score_o1w_kernel"): 2.0}, 1, "score_o1n_kernel"),
    (1, {(0, "memo_traj_kernel"): 0.2, (1, "memo_window_kernel"): 0.05, (2, "memo_window_kernel"): 0.18,
         (3, "score_o0_kernel(fallback list)"): 0.01, (4, "zoops_kernel"): 0.03}, 0, "memo_window_kernel"),
    (1, {(0, "memo_traj_kernel"): 0.4, (1, "memo_window_kernel"): 0.05, (2, "memo_window_kernel"): 0.18}, 0, "memo_traj_kernel"),
])
def test_roofline_object_for_every_scoring_path(path, kernels, order, expect):
    run = _run("c3" if order else "c3o0", order, 12, 11)
    traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    step_ms = sum(kernels.values())
    rl, extra = bench.roofline_of(run, path, kernels, step_ms, 15000, 33.9, 6550.0, "measured", traffic)
    assert rl["kernel"].startswith(expect) and 0 < rl["frac"] and 0 < rl["kernel_share_of_step"] <= 1
    assert rl["bound"] in ("fp64", "hbm") and rl["unit"] in ("TFLOP/s", "GB/s")
    if path == 1:
        assert extra["pattern_table"]["patterns"] == 100
    elif expect in ("score_o0j_kernel", "score_o1n_kernel"):
        assert rl["traffic"] and rl["traffic"] > 0                     # an ncu capture of this kernel is committed under profiles/


def test_flop_models_are_consistent():
    u0, f0 = bench.flops_order0(11, 12)
    assert u0 == 8 * (8 * 11 - 7 + 12) + 7 * 11 and f0 == 8 * 11 + 8 + 48 * 10 + 8 * 12
    u1, f1 = bench.flops_order1(11, 12, 12)
    assert u1 > 12 * u0 and f1 > 12 * f0                               # the network adds the horizontal terms
    cfg = dict(W=12, S=12, order=0)
    assert bench.algorithmic_flops(cfg, 11, 10, 150) == 12 * (150 * u0 + 160 * f0)
    assert bench.algorithmic_bytes(cfg, 10) == 10 * (5 + 8)
