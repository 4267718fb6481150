#!/usr/bin/env python
"""bench.py -- site log-likelihoods/s of the ZOOPS E-step (BASELINE.json's metric) on 1..8 B200.

A "step" is one pass of the hot path over one batch of synthetic alignments: background column scores, motif
window scores of every offset, the ZOOPS posterior (gamma, w, log-likelihood, arg-max) and, for N > 1, the
all-reduce of [ll, w0, w1] over NCCL.  Default workload: BASELINE.json configs[2] = north_star's target shape
(C3: order-1 phylogenetic Bayesian network, W = 12, 12-species tree, 10 000 alignments x 1 kb).  One process per
GPU; at N > 1 the SAME 10 000 alignments are split into contiguous blocks balanced by column count
(sharding.shard_bounds, the partition PhyloSample.shard makes): "strong" scaling, every rank scores different
alignments.  Before the timed region rank 0 scores the first alignments of its shard with the oracle and asserts
parity (scores 1e-9 relative, sweep counts and arg-max calls exact); the line carries `parity_checked`.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3|c1|c2|c2star|c3o0|c4|c5] [--impl reference]

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` goes through the public
host API with host buffers (pack + copies inside the timed region); `roofline` is for the dominant kernel
(score_o0_kernel on the motif windows) against the fp64 FMA peak measured on this GPU in the same run;
`cpu_baseline` is the oracle (C++ restatement of the Java algorithm; no JVM exists here) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

# the oracle (parity check, cpu_baseline) is OpenMP code: its idle workers must sleep, not spin, or they take the host cores
# the optimiser-iteration timings run on (measured: an order-0 M-step 3.3 ms, 9 to 200 ms next to spinning workers)
os.environ.setdefault("OMP_WAIT_POLICY", "passive")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "site_loglik_per_sec"
UNIT = "window scores/s"
DEFAULT_CONFIG = "c3"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=DEFAULT_CONFIG)
    ap.add_argument("--n-aln", type=int, default=None, help="override the number of alignments (debug only)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mstep", action="store_true", help="skip the secondary optimizer-iteration timings")
    ap.add_argument("--no-memo", action="store_true", help="switch the pattern-table path off (direct kernel only)")
    ap.add_argument("--memo-split", type=int, default=None, help="pattern_memo_split of the library (tuning only)")
    ap.add_argument("--memo-wave", type=int, default=None, help="pattern_memo_wave of the library (tuning only)")
    ap.add_argument("--order1-kernel", type=int, default=None, help="order1_kernel of the library (tuning only)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity check before the timed region")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configurations (c2, c3o0, c5 on 3000 alignments, c1)")
    return ap.parse_args()


def workload_name(cfg):
    kind = "phylogenetic PWM" if cfg["order"] == 0 else "phylogenetic Bayesian network (order 1)"
    if cfg.get("motifs"):
        kind = f"{cfg['motifs']} phylogenetic PWMs + homogeneous background (ZOOPS mixture of several motifs)"
    src = "bundled (reference data/synthetic_data/trees/data_tree1_1.0.fg, first 200)" if cfg.get("bundled") else "synthetic"
    return (f"{kind}, width {cfg['W']}, {cfg['S']}-species {cfg['tree']} tree, {cfg['n_aln']} {src} alignments x "
            f"{cfg['length']} bp, ZOOPS E-step, mean-field scoring (reference default)")


# ---------------------------------------------------------------------------------------------------------
# flop model of SURVEY.md section 8d (order 0, no gaps): per tree and sweep
# ---------------------------------------------------------------------------------------------------------
def flops_order0(I, S):
    update = 8 * (8 * I - 7 + S) + 7 * I
    fe = 8 * I + 8 + 48 * (I - 1) + 8 * S
    return update, fe


def flops_order1(I, S, W):
    """Order-1 network, no gaps, per WINDOW: (flop of one update sweep, flop of one free-energy evaluation) under the
    counting rule of SURVEY.md section 8d (an accumulation = 2 flop, a q factor beyond the first = 1 flop, normalisation
    = 7 flop per hidden node).  tests/test_oracle_pins.py checks these closed forms against the oracle's term counter."""
    upd = fe = 0
    for t in range(W):
        nxt = (32 + 192 * (I - 1)) if t + 1 < W else 0
        if t == 0:
            upd += 8 + 32 * (I - 1) + 32 * (I - 1) + 8 * S + nxt + 7 * I
            fe += 8 * I + 8 + 48 * (I - 1) + 8 * S
        else:
            upd += 32 + 192 * (I - 1) + 192 * (I - 1) + 8 * S + nxt + 7 * I
            fe += 8 * I + 48 + 256 * (I - 1) + 8 * S
    return upd, fe


def algorithmic_flops(cfg, I, n_windows, sweeps_total):
    W = cfg["W"]
    if cfg["order"] == 1:
        upd, fe = flops_order1(I, cfg["S"], W)
        return sweeps_total * upd + (sweeps_total + n_windows) * fe
    upd, fe = flops_order0(I, cfg["S"])
    return W * (sweeps_total * upd + (sweeps_total + n_windows) * fe)


def algorithmic_bytes(cfg, n_windows):
    # ceil(3S/8) bytes per column read once + 8 bytes per window score
    return n_windows * (-(-3 * cfg["S"] // 8) + 8)


# ---------------------------------------------------------------------------------------------------------
# clocks sampler (NVML) during the timed region
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, str(e)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (test infrastructure; executed here only as the reported baseline / reference arm)
# ---------------------------------------------------------------------------------------------------------
def oracle_models(cfg, newick, pwm):
    from oracle import oracle as orc
    fg = orc.Model(orc.FG, cfg["W"], cfg["order"], newick)
    for t, p in enumerate(pwm):
        fg.set_pi(t, np.asarray(p).ravel())
    bg = orc.Model(orc.BG, 0, 0, newick)
    # the reference's cost model: Math.log inside the inner loops, background ranges recomputed per offset
    fg.set_options(log_in_loop=True)
    bg.set_options(log_in_loop=True)
    return orc, fg, bg


def cpu_run(cfg, newick, pwm, sample, n_aln, threads):
    """E-step of the first n_aln alignments on `threads` host threads; returns (windows/s, seconds, windows)."""
    orc, fg, bg = oracle_models(cfg, newick, pwm)
    off = sample.col_offsets[: n_aln + 1]
    codes = sample.codes[: off[-1]]
    t0 = time.perf_counter()
    res = orc.estep(fg, bg, off, codes, np.log([0.9, 0.1]), faithful_cost=True, threads=threads, want=())
    dt = time.perf_counter() - t0
    nwin = int(np.maximum(np.diff(off) - cfg["W"] + 1, 0).sum())
    assert np.isfinite(res["ll"])
    return nwin / dt, dt, nwin


def cpu_sample_size(cfg, newick, pwm, sample, threads, target_s):
    """Pick how many alignments give about target_s seconds of CPU work (calibrated on a short run)."""
    n0 = min(sample.n_alignments, max(threads, 2))
    _, dt, _ = cpu_run(cfg, newick, pwm, sample, n0, threads)
    n = int(max(n0, min(sample.n_alignments, n0 * target_s / max(dt, 1e-3))))
    return n


def cpu_bounded_sample(cfg, newick, pwm, sample, threads, target_s):
    """(sample, n, note): the first n alignments of `sample` as about target_s seconds of CPU work.  When one alignment per thread
    already takes longer (C3: ~11 s), the alignments are cut to a prefix of their columns instead -- the cost of a window does not
    depend on the length of its alignment, and the rate is per window."""
    from phyfoo_b200.data import PhyloSample
    n0 = min(sample.n_alignments, max(threads, 2))
    _, dt, _ = cpu_run(cfg, newick, pwm, sample, n0, threads)
    if dt <= 1.5 * target_s:
        return sample, int(max(n0, min(sample.n_alignments, n0 * target_s / max(dt, 1e-3)))), ""
    keep = [max(4 * cfg["W"], int(sample.lengths()[i] * target_s / dt)) for i in range(n0)]
    cut = PhyloSample.from_alignments([sample.getElementAt(i)[:keep[i]] for i in range(n0)])
    return cut, n0, f", each cut to its first {keep[0]} columns"


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (oracle port; the Java cannot run: no JVM) on all host
    threads, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from phyfoo_b200 import synth
    cores = os.cpu_count() or 1
    full = dict(synth.CONFIGS[args.config])
    if args.n_aln:
        full["n_aln"] = args.n_aln
    # the pool the bounded samples are drawn from: the first alignments of the same seeded workload
    cfg, newick, pwm, sample, _ = synth.make_config(args.config, n_aln=min(full["n_aln"], 512))
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    budget = 120.0 / (steps + warmup)            # whole run within a few minutes
    sample, n, cut_note = cpu_bounded_sample(cfg, newick, pwm, sample, cores, min(budget, 10.0))
    for _ in range(warmup):
        cpu_run(cfg, newick, pwm, sample, n, cores)
    t_total, w_total = 0.0, 0
    for _ in range(steps):
        _, dt, nw = cpu_run(cfg, newick, pwm, sample, n, cores)
        t_total += dt
        w_total += nw
    value = w_total / t_total
    samp = (f"first {n} of {full['n_aln']} alignments{cut_note} per step ({w_total // steps} windows), oracle E-step with the "
            "reference's cost model")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * t_total / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(full), "sample": samp},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": samp},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
PARITY_ALIGNMENTS = {0: 16, 1: 3}      # alignments of the parity check, by motif order (the order-1 oracle is ~30x slower)


def parity_check(cfg, newick, pwm, sample, shm, k):
    """The first k alignments of this rank's shard through the public API against the oracle: gamma / ll at 1e-9 relative,
    mean-field sweep counts and arg-max site calls exact (the bars of tests/test_gpu_parity.py).  Raises on a mismatch."""
    from oracle import oracle as orc
    sub = sample.slice(0, k)
    res = shm.getNewWeights(sub, want_gamma=True)
    ofg = orc.Model(orc.FG, cfg["W"], cfg["order"], newick)
    for t, p in enumerate(pwm):
        ofg.set_pi(t, np.asarray(p).ravel())
    obg = orc.Model(orc.BG, 0, 0, newick)
    ref = orc.estep(ofg, obg, sub.col_offsets, sub.codes, np.log(shm.weights), threads=os.cpu_count() or 1, want=("gamma",))
    err_ll = abs(res["ll"] - ref["ll"]) / abs(ref["ll"])
    err_g = float(np.max(np.abs(res["gamma"] - ref["gamma"])))
    ok = (err_ll <= 1e-9 and err_g <= 1e-9 and np.array_equal(res["argmax_off"], ref["argmax_off"])
          and res["stats"]["sweeps_fg"] == ref["sweeps_fg"] and 2 * res["stats"]["sweeps_bg"] == ref["sweeps_bg"])
    if not ok:
        raise SystemExit(f"bench.py: parity check against the oracle FAILED on {cfg}: ll rel err {err_ll:.3e}, gamma abs err "
                         f"{err_g:.3e}, sweeps {res['stats']['sweeps_fg']} vs {ref['sweeps_fg']}, arg-max equal: "
                         f"{np.array_equal(res['argmax_off'], ref['argmax_off'])}")
    sub.device(shm.motif.ctx).free()
    return {"alignments": int(k), "windows": int(res["gamma"].size), "ll_rel_err": err_ll, "gamma_abs_err": err_g,
            "sweep_counts": "identical", "argmax_calls": "identical",
            "against": "the CPU oracle, which reproduces to the last bit the reference's own statements executed from its source "
                       "files (tests/test_reference_statements.py; no JVM exists in the image)"}


class EStepRunner:
    """One configuration resident on this rank's GPU: models, packed shard, the device-side E-step."""

    def __init__(self, ctx, torch, dist, world, rank, name, n_aln=None, stream=None, flush=None):
        from phyfoo_b200 import sharding, synth
        from phyfoo_b200.models import PhyloBackground, PhyloBayesModel, SingleHiddenMotifMixture
        self.ctx, self.torch, self.dist, self.world, self.rank, self.name = ctx, torch, dist, world, rank, name
        self.cfg, self.newick, self.pwm, full, self.planted = synth.make_config(name, n_aln=n_aln)
        W = self.cfg["W"]
        self.total_windows = int(np.maximum(full.lengths() - W + 1, 0).sum())
        self.total_alignments = full.n_alignments
        if world > 1:
            self.bounds = sharding.shard_bounds(full.col_offsets, world)
            lo, hi = self.bounds[rank]
            self.sample = full.slice(lo, hi)
        else:
            self.bounds = [(0, full.n_alignments)]
            self.sample = full
        del full
        self.fg = PhyloBayesModel(W, self.cfg["order"], self.newick, ctx)
        self.fg.setCondProbs(self.pwm)
        self.bg = PhyloBackground(0, self.newick, ctx)
        self.shm = SingleHiddenMotifMixture(self.fg, self.bg, zoops=0.9)
        self.logw = np.log(self.shm.weights)
        self.ds = self.sample.device(ctx)
        self.n_windows = self.ds.windows(W)
        self.I = self.fg.tree.n_nodes - self.fg.tree.n_species
        self.partial = torch.zeros(3, dtype=torch.float64, device="cuda")
        self.stream, self.flush = stream, flush

    def step(self):
        from phyfoo_b200 import core
        core.zoops_estep_device(self.ctx, self.ds, self.fg.device_model(), self.bg.device_model(), self.logw, None, None,
                                self.partial.data_ptr(), self.stream.cuda_stream)
        if self.world > 1:
            self.dist.all_reduce(self.partial)

    def timed(self, n_steps):
        """n_steps E-steps, L2 flushed before each (untimed), CUDA events on the launching stream.  Returns the summed
        device ms and the per-kernel device ms (library events), keyed by launch order."""
        torch = self.torch
        total, per_kernel = 0.0, {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(n_steps):
            self.flush.fill_(1)
            e0.record(self.stream)
            self.step()
            e1.record(self.stream)
            e1.synchronize()
            total += e0.elapsed_time(e1)
            for i, (name, ms) in enumerate(self.ctx.last_kernel_times()):
                key = (i, name)
                per_kernel[key] = per_kernel.get(key, 0.0) + ms
        return total, {k: v / n_steps for k, v in per_kernel.items()}

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def free(self):
        self.ds.free()


def fp64_roofline(flops, abytes, kernel_name, kernel_ms, peak_tflops, hbm_peak, hbm_of, note=None):
    achieved = flops / (kernel_ms * 1e-3) / 1e12
    r = {"bound": "fp64", "kernel": kernel_name, "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
         "frac": achieved / peak_tflops,
         "peak_source": "DFMA microbenchmark in this run (MEASURED_PEAKS.json has no fp64 entry)",
         "flops_per_launch": flops, "kernel_ms": kernel_ms,
         "hbm": {"achieved": abytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                 "frac": abytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak, "of": hbm_of},
         "traffic": None}
    if note:
        r["note"] = note
    return r


def roofline_of(run, path, kernels, step_ms, sweeps_fg, peak_tflops, hbm_peak, hbm_of, traffic_db, direct=None):
    """roofline of the kernel(s) with the largest share of the step + extra sub-objects (pattern-table path)."""
    cfg, I, S, W, n_windows = run.cfg, run.I, run.cfg["S"], run.cfg["W"], run.n_windows
    flops = algorithmic_flops(cfg, I, n_windows, sweeps_fg)
    abytes = algorithmic_bytes(cfg, n_windows)
    names = {name for (_, name) in kernels}
    extra = {}

    def traffic_of(*keys):
        # profiles/traffic.json: "<profile label>:<kernel name as ncu prints it>" -> dram bytes read + written per launch
        for key in keys:
            label, _, kernel = key.partition(":")
            for have, v in sorted(traffic_db.items(), reverse=True):          # the latest capture of a configuration first
                hl, _, hk = have.partition(":")
                if (hl == label or hl.startswith(label + "_")) and (hk == kernel or hk.startswith(kernel + "<")):
                    return v
        return None

    o1 = [n for n in ("score_o1n_kernel", "score_o1w_kernel", "score_o1_kernel") if n in names]
    o0 = [n for n in ("score_o0j_kernel", "score_o0f_kernel", "score_o0_kernel") if n in names]
    if path != 1:
        fg_name = o1[0] if cfg["order"] == 1 and o1 else o0[0]
        # the motif-window launches of the scoring kernel (an order-1 step may add a fallback launch for windows with gaps)
        fg_ms = max(ms for (_, name), ms in kernels.items() if name == fg_name)
        roofline = fp64_roofline(flops, abytes, fg_name + " (motif windows)", fg_ms, peak_tflops, hbm_peak, hbm_of)
        roofline["kernel_share_of_step"] = fg_ms / step_ms
        roofline["traffic"] = traffic_of(f"r02_{run.name}:{fg_name}", f"r01_{run.name}_direct_kernel:score_o0_kernel<0, 0>")
        if cfg["order"] == 1:
            roofline["note"] = ("flops_per_launch is the ALGORITHMIC count of SURVEY.md section 8d (the reference's generic 16-row "
                                "contractions, measured sweep counts); the kernel executes fewer fp64 operations than that (F81 "
                                "structure of the tables, shared context sums, no logarithm per node), so frac can exceed the share "
                                "of the fp64 pipe ncu reports")
    else:
        # pattern-table path: per-pattern trajectories (fp64, a small latency-bound grid) + gathers per window
        D = run.ds.patterns()
        win_ms = sum(ms for (_, name), ms in kernels.items() if name == "memo_window_kernel")     # background + motif launches
        traj_ms = sum(ms for (_, name), ms in kernels.items() if name.startswith("memo_traj"))    # both phases
        upd, fe = flops_order0(I, S)
        K1 = 65              # pattern_memo_sweep_cap (64) + 1 passes per tree
        traj_flops = D * (W + 1) * K1 * (upd + fe)            # motif positions + the background tree
        gather_bytes = W * (sweeps_fg + n_windows) * 8         # trajectory entries the stop rule consumes
        if win_ms >= traj_ms:
            roofline = {"bound": "hbm", "kernel": "memo_window_kernel (background + motif launches; gathers from an L2-resident table)",
                        "achieved": abytes / (win_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": abytes / (win_ms * 1e-3) / 1e9 / hbm_peak, "of": hbm_of,
                        "bytes_per_launch": abytes, "kernel_ms": win_ms, "kernel_share_of_step": win_ms / step_ms,
                        "table_gather": {"bytes": gather_bytes, "GB/s": gather_bytes / (win_ms * 1e-3) / 1e9,
                                         "what": "8 B x W x (K_w + 1) trajectory entries per window, served by L2"},
                        "traffic": traffic_of("r01_c2_pattern_table:memo_window_kernel<0, 4>", "r01_c2_pattern_table:memo_window_kernel")}
        else:
            achieved = traj_flops / (traj_ms * 1e-3) / 1e12
            roofline = {"bound": "fp64", "kernel": "memo_traj_kernel (mean-field trajectories of the occurring column patterns)",
                        "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                        "peak_source": "DFMA microbenchmark in this run (MEASURED_PEAKS.json has no fp64 entry)",
                        "flops_per_launch": traj_flops, "kernel_ms": traj_ms, "kernel_share_of_step": traj_ms / step_ms,
                        "note": f"{D} patterns x {W + 1} trees x {K1} trajectory entries each; flops count every entry although a "
                                "tree stops once its q repeats bit for bit (the rest of its row is filled: about half of the "
                                "entries are executed); a small grid of dependent steps, bound by issue slots and instruction "
                                "latency, not by the fp64 pipe",
                        "traffic": traffic_of("r01_c2_pattern_table:memo_traj_wave_kernel<8>", "r01_c2_pattern_table:memo_traj_kernel")}
        extra["pattern_table"] = {"patterns": D, "trajectory_flops_per_step": traj_flops, "direct_flops_per_step": flops,
                                  "work_ratio": flops / max(traj_flops, 1), "trajectory_ms": traj_ms, "gather_ms": win_ms}
        if direct:
            d_ms, d_kernels = direct
            d_name = [n for n in ("score_o0j_kernel", "score_o0f_kernel", "score_o0_kernel") if n in {nm for (_, nm) in d_kernels}][0]
            d_fg = max(ms for (_, name), ms in d_kernels.items() if name == d_name)
            extra["roofline_direct"] = fp64_roofline(
                flops, abytes, d_name + " (motif windows, direct path: pattern table switched off)", d_fg, peak_tflops,
                hbm_peak, hbm_of, "what every configuration with more than 6 species runs; measured in this run, 3 steps, not part of `value`")
            extra["roofline_direct"]["step_ms"] = d_ms
    return roofline, extra


def optimizer_timings(run, args, step_ms):
    """one forward-difference gradient of the M-step objective (dim + 1 objective sums in one pass,
    plus the all-reduce at N > 1) and one whole EM iteration (E-step + M-step of the motif model: selection, fixing q,
    position-wise BFGS).  Wall clock, max over ranks.  They change the model, so they come last."""
    from phyfoo_b200 import core, sharding
    ctx, fg, bg, ds, cfg, world = run.ctx, run.fg, run.bg, run.ds, run.cfg, run.world
    W = cfg["W"]
    core.zoops_estep(ctx, ds, fg.device_model(), bg.device_model(), run.logw, None, None, True, False, False)   # gamma stays resident
    prepare_ms = None
    for _ in range(2):                                  # the second time: scratch buffers exist (as in every later EM iteration)
        if prepare_ms is not None:
            fe.free()
        t0 = time.perf_counter()
        sa, so, sw = core.select(ctx, ds, W, None, fg.PROB_THRESH_FOR_WEIGHTS, fg.MOTIF_QUALITY_THRESHOLD)
        fe = core.FESet(ctx, ds, fg.device_model(), sa, so, sw)
        ctx.synchronize()
        prepare_ms = (time.perf_counter() - t0) * 1e3
    obj = sharding.ShardedObjective(fe, ctx.device) if world > 1 else fe
    pos = 1 if cfg["order"] == 1 else 0                 # an order-1 motif: a position with four context rows (dim 16)
    lam = np.log(np.asarray(fg.getCondProb(pos), np.float64).ravel())
    for _ in range(3):
        obj.grad(pos, lam)
    if world > 1:
        run.dist.barrier()
    t0 = time.perf_counter()
    n_grad = 20
    for _ in range(n_grad):
        obj.grad(pos, lam)
    grad_ms = (time.perf_counter() - t0) * 1e3 / n_grad
    fe.free()

    def mstep(suffstats):
        fg.setCondProbs(run.pwm)
        core.zoops_estep(ctx, ds, fg.device_model(), bg.device_model(), run.logw, None, None, True, False, False)
        fg.trainingStep = fg.SKIP_TRAINING_STEPS          # the reference ignores the first call; time a real M-step
        fg.MSTEP_SUFFICIENT_STATISTICS = suffstats
        if world > 1:
            run.dist.barrier()
        t0 = time.perf_counter()
        st = fg.train(run.sample, None, rng=np.random.default_rng(1), sharded=world > 1)
        return (time.perf_counter() - t0) * 1e3, st

    repeats = 2                                         # the first M-step of a process pays for its scratch allocations
    out = {"gradient_ms": grad_ms, "dim": int(lam.size), "position": pos, "selected_windows_per_gpu": int(sa.size),
           "select_and_fix_q_ms": prepare_ms}
    # default M-step: sufficient statistics, one library call (phyfoo_ss_optimize); beside it the per-window evaluation on the
    # device under the same optimiser (the reference's own evaluation order)
    mstep_ms, st = min((mstep(True) for _ in range(repeats)), key=lambda r: r[0])
    mstep_dev_ms, st_dev = min((mstep(False) for _ in range(repeats)), key=lambda r: r[0])
    fg.MSTEP_SUFFICIENT_STATISTICS = True
    grad_ms, mstep_ms, prepare_ms, mstep_dev_ms = run.max_over_ranks([grad_ms, mstep_ms, prepare_ms, mstep_dev_ms])
    out.update({"mstep_device_evaluation_ms": mstep_dev_ms, "mstep_device_evaluation_evaluations": int(st_dev["evaluations"]),
                "em_iteration_device_evaluation_ms": step_ms + mstep_dev_ms})
    fg.setCondProbs(run.pwm)
    out.update({"gradient_ms": grad_ms, "select_and_fix_q_ms": prepare_ms, "mstep_ms": mstep_ms,
                "mstep_evaluations": int(st["evaluations"]), "em_iteration_ms": step_ms + mstep_ms,
                "what": "wall clock, max over ranks: evaluateGradientOfFunction of one motif position on the device (forward "
                        "differences, dim + 1 objective sums over the selected windows in one pass" + (", one all-reduce" if world > 1 else "") +
                        "); M-step = selection + mean fields of the selected windows + BFGS position by position; EM iteration = "
                        "E-step (ms_per_step) + M-step.  The M-step runs on sufficient statistics (one device pass for the "
                        "expected counts" + (", one all-reduce of them" if world > 1 else "") + ", then the optimiser and its "
                        "evaluations on the host inside libphyfoo.so: phyfoo_ss_optimize); mstep_device_evaluation_ms is the same "
                        "optimiser with a device pass per evaluation"})
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (phyfoo_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        # a short collective timeout: a rank-asymmetric bug must abort the run in minutes, not hold 8 GPUs for the default 10
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=300))

    from phyfoo_b200 import core
    from phyfoo_b200.data import PhyloSample

    ctx = core.Context(local)
    if args.memo_split is not None:
        ctx.set_option("pattern_memo_split", args.memo_split)
    if args.memo_wave is not None:
        ctx.set_option("pattern_memo_wave", args.memo_wave)
    if args.no_memo:
        ctx.set_option("pattern_memo", 0)
    if args.order1_kernel is not None:
        ctx.set_option("order1_kernel", args.order1_kernel)
    warmup = max(args.warmup, 3)                 # timing rule: at least three untimed steps; the line reports what ran
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    # a non-default torch stream: handle 0 would mean "the library's own stream" to the C ABI
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)

    run = EStepRunner(ctx, torch, dist, world, rank, args.config, args.n_aln, stream, flush)
    cfg, sample, W, S = run.cfg, run.sample, run.cfg["W"], run.cfg["S"]

    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_check(cfg, run.newick, run.pwm, sample, run.shm, min(PARITY_ALIGNMENTS[cfg["order"]], sample.n_alignments))

    for _ in range(warmup):
        run.step()
    torch.cuda.synchronize()
    launches_per_step = ctx.last_launches()
    path = ctx.last_score_path()
    peak_tflops = ctx.fp64_peak()

    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms, kernels = run.timed(args.steps)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    sweeps_fg, sweeps_bg = ctx.estep_sweeps()
    ll_w = run.partial.cpu().numpy().copy()

    dev_ms_max, = run.max_over_ranks([dev_ms])
    total_windows = run.total_windows                  # strong scaling: the ranks' shards add up to the one data set
    value = total_windows * args.steps / (dev_ms_max * 1e-3)
    step_ms = dev_ms_max / args.steps

    # the direct kernel on the same workload (the path every configuration with more than 6 species takes), for the roofline
    # (every rank runs it: step() contains the all-reduce, so a rank-0-only run would leave rank 0 alone in a collective)
    direct = None
    if path == 1:
        ctx.set_option("pattern_memo", 0)
        run.step(); torch.cuda.synchronize()
        d_ms, d_kernels = run.timed(3)
        ctx.set_option("pattern_memo", 1)
        run.step(); torch.cuda.synchronize()
        direct = (d_ms / 3, d_kernels)

    # ---- e2e: public host API, host buffers, pack + copies inside the timed region (per rank, max over ranks).
    # Inputs come from page-locked host memory, results land in page-locked host memory that the caller reuses.
    pinned = ctx.pinned_empty(sample.codes.shape, np.uint8)
    pinned[:] = sample.codes
    n_windows = run.n_windows
    out = {"gamma": ctx.pinned_empty(n_windows), "argmax_off": ctx.pinned_empty(sample.n_alignments, np.int32),
           "argmax_strand": ctx.pinned_empty(sample.n_alignments, np.uint8)}
    e2e_steps = max(3, min(args.steps, 10 if step_ms < 100 else 5))

    def e2e_step():
        smp = PhyloSample(pinned, sample.col_offsets)
        res = run.shm.getNewWeights(smp, want_gamma=True, out=out)
        smp.device(ctx).free()
        return res

    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = e2e_step()
    torch.cuda.synchronize()
    e2e_s, = run.max_over_ranks([time.perf_counter() - t0])
    e2e_value = total_windows * e2e_steps / e2e_s
    h2d = int(sample.codes.nbytes + sample.col_offsets.nbytes)
    d2h = int(res["gamma"].nbytes + res["argmax_off"].nbytes + res["argmax_strand"].nbytes + 24)
    assert world > 1 or abs(res["ll"] - ll_w[0]) <= 1e-9 * abs(res["ll"])

    # arg-max site calls against the planted sites of this rank's alignments (c1: the MOTIF_POS annotations of the bundled data)
    planted_hits = None
    pl = np.asarray(run.planted)[run.bounds[rank][0]:run.bounds[rank][1]]
    has = pl >= 0
    if has.any():
        d = np.abs(res["argmax_off"][has].astype(np.int64) - pl[has])
        planted_hits = {"alignments_with_a_site": int(has.sum()), "of": f"rank 0's {sample.n_alignments} alignments",
                        "argmax_exact": float(np.mean(d == 0)), "argmax_within_half_width": float(np.mean(d <= W // 2)),
                        "model": "the generating PWM and tree"}

    optimizer = None
    if not args.no_mstep:
        optimizer = optimizer_timings(run, args, step_ms)

    # ---- secondary configurations (N = 1 only, short): C2 (BASELINE.json configs[1]) and C3's data with an order-0 motif
    extras = {}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_of = "measured (MEASURED_PEAKS.json)" if peaks else "fallback"
    try:
        traffic_db = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        traffic_db = {}
    if world == 1 and not args.no_extra and args.config == DEFAULT_CONFIG and not args.n_aln:
        from phyfoo_b200.synth import CONFIGS as synth_configs
        run.free()
        # C2 (configs[1], pattern table), C3's data with an order-0 motif, the 30-species tree of C5 (configs[4]) on 3000 of its
        # alignments, and C1 (configs[0]: the reference's bundled alignments)
        for name, n_ex_aln in (("c2", None), ("c3o0", None), ("c5", 3000), ("c1", None)):
            ex = EStepRunner(ctx, torch, dist, 1, 0, name, n_ex_aln, stream, flush)
            pc = parity_check(ex.cfg, ex.newick, ex.pwm, ex.sample, ex.shm, PARITY_ALIGNMENTS[0]) if not args.no_parity else None
            for _ in range(3):
                ex.step()
            torch.cuda.synchronize()
            n_ex = 20 if name in ("c2", "c1") else 5 if name == "c3o0" else 3
            ms, kern = ex.timed(n_ex)
            sw_fg, _ = ctx.estep_sweeps()
            rl, ex_extra = roofline_of(ex, ctx.last_score_path(), kern, ms / n_ex, sw_fg, peak_tflops, hbm_peak, hbm_of, traffic_db)
            extras[name] = {"workload": workload_name(ex.cfg), "value": ex.n_windows * n_ex / (ms * 1e-3), "unit": UNIT,
                            "ms_per_step": ms / n_ex, "steps": n_ex, "warmup": 3, "parity_checked": pc,
                            "scoring_path": PATHS[ctx.last_score_path()], "roofline": rl,
                            "kernels_ms": [{"kernel": nm, "ms": round(v, 5)} for (_, nm), v in sorted(kern.items())]}
            extras[name].update(ex_extra)
            am = ex.shm.getNewWeights(ex.sample, want_gamma=False)["argmax_off"]
            pl = np.asarray(ex.planted)
            if (pl >= 0).any():
                d = np.abs(am[pl >= 0].astype(np.int64) - pl[pl >= 0])
                extras[name]["planted_sites"] = {"alignments_with_a_site": int((pl >= 0).sum()), "argmax_exact": float(np.mean(d == 0)),
                                                 "argmax_within_half_width": float(np.mean(d <= ex.cfg["W"] // 2))}
            if n_ex_aln:
                extras[name]["alignments"] = f"{n_ex_aln} of the configuration's {synth_configs[name]['n_aln']}"
            ex.free()

    if rank == 0:
        kern_list = [{"kernel": name, "ms": round(ms, 5)} for (_, name), ms in sorted(kernels.items())]
        roofline, extra = roofline_of(run, path, kernels, step_ms, sweeps_fg, peak_tflops, hbm_peak, hbm_of, traffic_db, direct)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "bundled with the reference (synthetic, its own generator)" if cfg.get("bundled") else "synthetic",
            "config": {"workload": workload_name(cfg), "name": args.config, "windows_total": total_windows,
                       "windows_per_gpu": n_windows, "alignments_per_gpu": [hi - lo for lo, hi in run.bounds],
                       "sharding": "contiguous alignment blocks balanced by column count (PhyloSample.shard); every rank "
                                   "scores different alignments of the one data set" if world > 1 else "none (one GPU)",
                       "mean_sweeps_per_window": sweeps_fg / max(n_windows, 1), "seed": hex(cfg["seed"]),
                       "scoring_path": PATHS[path],
                       "l2": "flushed between timed iterations (512 MB fill)",
                       "collective": "NCCL all-reduce of 3 fp64 per step" if world > 1 else "none"},
            "parity_checked": parity,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "what": "PhyloSample (page-locked host codes) -> pack on device -> E-step -> gamma/arg-max/ll to page-locked host buffers"},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
            "roofline": roofline,
            "kernels_ms": kern_list,
            "wall_s": t_wall,
        }
        if args.warmup != warmup:
            line["warmup_requested"] = args.warmup
        if planted_hits:
            line["planted_sites"] = planted_hits
        if optimizer:
            line["optimizer"] = optimizer
        line.update(extra)
        if extras:
            line["extra"] = extras
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n = cpu_sample_size(cfg, run.newick, run.pwm, sample, cores, args.cpu_seconds)
            v, dt, nw = cpu_run(cfg, run.newick, run.pwm, sample, n, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {n} of {cfg['n_aln']} alignments ({nw} windows, {dt:.1f} s), "
                                              "C++ restatement of the Java algorithm with its cost model (JVM unavailable)"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: ZOOPS mixture of several motifs (c4), alignments sharded over the ranks
# ---------------------------------------------------------------------------------------------------------
def run_multi_motif(args):
    """--config c4: 3 phylogenetic PWMs + homogeneous background, 100 000 alignments x 1 kb over the ranks (every rank
    generates and scores its own block of n_aln / N alignments; weak in memory, strong in the problem: the blocks add up to
    the one data set).  A step = background columns + K motif scoring launches + the K-component posterior
    (phyfoo_zoops_estep_multi) and, at N > 1, the all-reduce of [ll, w_0..w_K]."""
    import torch
    import torch.distributed as dist
    from phyfoo_b200 import core, sharding, synth
    from phyfoo_b200.data import PhyloSample
    from phyfoo_b200.models import MultiMotifMixture, PhyloBackground, PhyloBayesModel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (phyfoo_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=600))
    ctx = core.Context(local)
    warmup = max(args.warmup, 3)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    cfg, newick, pwms, sample, planted, kind = synth.make_config_multi(args.config, n_aln=args.n_aln, shard=(rank, world))
    W, K = cfg["W"], cfg["motifs"]
    fgs = []
    for p in pwms:
        m = PhyloBayesModel(W, cfg["order"], newick, ctx)
        m.setCondProbs(p)
        fgs.append(m)
    bg = PhyloBackground(0, newick, ctx)
    wts = np.concatenate([np.full(K, cfg["zoops"] / K), [1.0 - cfg["zoops"]]])
    mix = MultiMotifMixture(fgs, bg, weights=wts)
    logw = np.log(wts)
    I = fgs[0].tree.n_nodes - fgs[0].tree.n_species

    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import oracle as orc
        sub = sample.slice(0, min(8, sample.n_alignments))
        res = mix.getNewWeights(sub)
        ofgs = []
        for p in pwms:
            o = orc.Model(orc.FG, W, cfg["order"], newick)
            for t, row in enumerate(p):
                o.set_pi(t, np.asarray(row).ravel())
            ofgs.append(o)
        ref = orc.estep_multi(ofgs, orc.Model(orc.BG, 0, 0, newick), sub.col_offsets, sub.codes, logw)
        err_ll = abs(res["ll"] - ref["ll"]) / abs(ref["ll"])
        err_g = max(float(np.max(np.abs(g - r))) for g, r in zip(res["gamma"], ref["gamma"]))
        same = np.array_equal(res["argmax_off"], ref["argmax_off"]) and np.array_equal(res["argmax_motif"], ref["argmax_motif"])
        if not (err_ll <= 1e-9 and err_g <= 1e-9 and same):
            raise SystemExit(f"bench.py: c4 parity check against the specification FAILED: ll {err_ll:.3e}, gamma {err_g:.3e}, arg-max equal {same}")
        sub.device(ctx).free()
        parity = {"alignments": sub.n_alignments, "windows": int(sum(g.size for g in res["gamma"])), "ll_rel_err": err_ll,
                  "gamma_abs_err": err_g, "argmax_calls": "identical",
                  "against": "oracle/oracle.py::estep_multi (the specification of DESIGN.md on the oracle's window and column scores)"}

    ds = sample.device(ctx)
    handles = [m.device_model() for m in fgs]

    def step():
        r = core.zoops_estep_multi(ctx, ds, handles, bg.device_model(), logw, None, want_gamma=False, want_argmax=False)
        v = np.concatenate([[r["ll"]], r["w"]])
        if world > 1:
            sharding.allreduce_host_(v)
        return r, v

    for _ in range(warmup):
        step()
    peak_tflops = ctx.fp64_peak()
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    dev_ms, per_kernel, sweeps = 0.0, {}, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        r, v = step()
        dev_ms += r["stats"]["gpu_ms"]
        sweeps = r["stats"]["sweeps_fg"]
        for i, (name, ms) in enumerate(ctx.last_kernel_times()):
            per_kernel[(i, name)] = per_kernel.get((i, name), 0.0) + ms / args.steps
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    n_win_local = K * ds.windows(W)
    tot = np.array([float(n_win_local), float(sample.n_alignments)])
    mx = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        sharding.allreduce_host_(tot)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dev_ms_max = float(mx.item())
    step_ms = dev_ms_max / args.steps
    value = tot[0] * args.steps / (dev_ms_max * 1e-3)

    # e2e: host codes -> pack -> E-step -> gamma of every component / arg-max / ll to the host
    pinned = ctx.pinned_empty(sample.codes.shape, np.uint8)
    pinned[:] = sample.codes
    e2e_steps = 3

    def e2e_step():
        smp = PhyloSample(pinned, sample.col_offsets)
        r = mix.getNewWeights(smp, want_gamma=True)
        smp.device(ctx).free()
        return r

    res = e2e_step()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = e2e_step()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = tot[0] * e2e_steps / float(e2e_s.item())
    hit = None
    has = planted >= 0
    if has.any():
        hit = {"alignments_with_a_site": int(has.sum()), "of": f"rank 0's {sample.n_alignments} alignments",
               "argmax_motif_and_offset_exact": float(np.mean((res["argmax_off"][has] == planted[has]) & (res["argmax_motif"][has] == kind[has])))}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        o0_ms = [ms for (_, name), ms in sorted(per_kernel.items()) if name.startswith("score_o0")]
        fg_ms = sum(o0_ms) - min(o0_ms)                   # minus the background launch
        fg_kernel = [name for (_, name), _ in sorted(per_kernel.items()) if name.startswith("score_o0")][-1]
        one = dict(cfg)
        flops = algorithmic_flops(one, I, ds.windows(W), sweeps // K) * K
        roofline = fp64_roofline(flops, K * algorithmic_bytes(one, ds.windows(W)), f"{fg_kernel} ({K} motif launches per step)",
                                 fg_ms, peak_tflops, peaks.get("hbm_gbs", 6650.0),
                                 "measured (MEASURED_PEAKS.json)" if peaks else "fallback")
        roofline["kernel_share_of_step"] = fg_ms / (dev_ms / args.steps)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(cfg), "name": args.config, "windows_total": int(tot[0]),
                           "alignments_total": int(tot[1]), "windows_per_gpu": int(n_win_local),
                           "what_counts": "one window score per motif component and offset",
                           "sharding": "every rank generates and scores its own block of n_aln / N alignments" if world > 1 else "none (one GPU)",
                           "mean_sweeps_per_window": sweeps / max(n_win_local, 1),
                           "l2": "flushed between timed iterations (512 MB fill)",
                           "collective": f"NCCL all-reduce of {K + 2} fp64 per step" if world > 1 else "none"},
                "parity_checked": parity,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(sample.codes.nbytes + sample.col_offsets.nbytes),
                        "d2h_bytes_per_step": int(sum(g.nbytes for g in res["gamma"]) + 8 * sample.n_alignments + 8 * (K + 2)),
                        "steps": e2e_steps, "what": "host codes -> pack on device -> K-motif E-step -> gamma of every component / arg-max / ll on the host"},
                "gpu_launches": int(ctx.last_launches() * args.steps), "clocks": clocks, "roofline": roofline,
                "kernels_ms": [{"kernel": nm, "ms": round(v, 5)} for (_, nm), v in sorted(per_kernel.items())],
                "wall_s": wall, "ll": float(v[0]), "component_weights": [float(x) for x in v[1:]]}
        if hit:
            line["planted_sites"] = hit
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


PATHS = {0: "direct kernel", 1: "pattern table (bit-identical to the direct kernel)", 2: "order-1 network kernel"}


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else a library prints (e.g. NCCL's version banner) was
    re-routed to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)            # C-level stdout of this process (NCCL, CUDA libraries) -> stderr
    args = parse_args()
    from phyfoo_b200 import synth
    if args.impl == "reference":
        run_reference(args)
    elif synth.CONFIGS.get(args.config, {}).get("motifs"):
        run_multi_motif(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
