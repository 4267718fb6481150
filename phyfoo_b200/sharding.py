"""Data parallelism over alignments: one process per GPU, a contiguous block of alignments per
rank balanced by column count, the model replicated, and ONE exchange per consumer -- a sum all-reduce of the few
doubles the consumer needs ([ll, w0, w1] per E-step, [f, f_1..f_dim] per gradient).  gamma, arg-max calls and
per-window scores stay local and are concatenated in rank order only when the caller asks.

torch.distributed is the plumbing (NCCL on the GPUs, gloo in the CPU tests); nothing here computes scores.
"""
import numpy as np


def shard_bounds(col_offsets, world):
    """[(lo, hi)] per rank: contiguous alignment blocks whose column counts are as equal as the boundaries allow.  With
    at least `world` alignments every rank gets at least one (very unequal lengths would otherwise leave a rank without
    work); with fewer, the last ranks get empty blocks and must contribute zeros to the all-reduce (ShardedEStep does)."""
    col_offsets = np.asarray(col_offsets, np.int64)
    n = col_offsets.size - 1
    total = int(col_offsets[-1])
    cuts = [int(np.searchsorted(col_offsets, total * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, n
    for r in range(1, world):
        lo = r if n >= world else min(r, n)                # rank r-1 keeps at least one alignment ...
        hi = n - (world - r) if n >= world else n          # ... and so do the ranks behind this cut
        cuts[r] = min(max(cuts[r], cuts[r - 1], lo), hi)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def world():
    d = _dist()
    return (d.get_rank(), d.get_world_size()) if d else (0, 1)


def allreduce_sum_(tensor):
    """In-place sum over ranks; a no-op in a single process.  The reduction order over ranks is NCCL's/gloo's: results
    agree with the single-GPU sum to fp64 re-association (compare at 1e-12 relative)."""
    d = _dist()
    if d and d.get_world_size() > 1:
        d.all_reduce(tensor, op=d.ReduceOp.SUM)
    return tensor


def allreduce_host_(values):
    """Sum of a small host vector over the ranks (numpy float64 in, the same array out, identical on every rank)."""
    d = _dist()
    if not d or d.get_world_size() == 1:
        return values
    import torch
    t = torch.from_numpy(np.ascontiguousarray(values, np.float64).copy())
    if d.get_backend() == "nccl":
        t = t.cuda()
    d.all_reduce(t, op=d.ReduceOp.SUM)
    values[...] = t.cpu().numpy().reshape(np.shape(values))
    return values


def broadcast_host_(values, src=0):
    """Rank `src`'s copy of a small host array (int64 or float64) on every rank -- e.g. the random order of the motif
    positions in the M-step, which every rank must walk alike (util/PWMUtil.java:131-144 draws it from an unseeded RNG)."""
    d = _dist()
    if not d or d.get_world_size() == 1:
        return values
    import torch
    t = torch.from_numpy(np.ascontiguousarray(values).copy())
    if d.get_backend() == "nccl":
        t = t.cuda()
    d.broadcast(t, src=src)
    values[...] = t.cpu().numpy().reshape(np.shape(values))
    return values


def gather_in_rank_order(array):
    """Concatenation of per-rank numpy arrays in rank order on every rank (arg-max calls, gamma when asked for)."""
    d = _dist()
    if not d or d.get_world_size() == 1:
        return np.asarray(array)
    parts = [None] * d.get_world_size()
    d.all_gather_object(parts, np.asarray(array))
    return np.concatenate(parts)


class ShardedEStep:
    """The ZOOPS E-step of a SingleHiddenMotifMixture over this rank's shard with the all-reduce of [ll, w0, w1]
    enqueued on the same CUDA stream right behind the reduction kernel (no host round trip in between)."""

    def __init__(self, shm, sample, rank=None, world_size=None):
        import torch
        r, w = world()
        self.rank = r if rank is None else rank
        self.world = w if world_size is None else world_size
        self.shm = shm
        self.bounds = shard_bounds(sample.col_offsets, self.world)
        lo, hi = self.bounds[self.rank]
        self.local = sample.slice(lo, hi) if self.world > 1 else sample
        self.ctx = shm.motif.ctx
        self.partial = torch.zeros(3, dtype=torch.float64, device=f"cuda:{self.ctx.device}")
        self.stream = torch.cuda.Stream(device=self.ctx.device)

    def step(self):
        """Asynchronous: returns after enqueueing; call result() for the host values."""
        import torch
        from . import core
        shm = self.shm
        with np.errstate(divide="ignore"):
            logw = np.log(shm.weights)
        with torch.cuda.stream(self.stream):
            if self.local.n_alignments == 0:
                self.partial.zero_()                   # an empty shard (fewer alignments than ranks) contributes zeros
            else:
                core.zoops_estep_device(self.ctx, self.local.device(self.ctx), shm.motif.device_model(),
                                        shm.flanking.device_model(), logw,
                                        None if shm.strand is None else shm.strand.logWeights(), None,
                                        self.partial.data_ptr(), self.stream.cuda_stream)
            allreduce_sum_(self.partial)

    def result(self):
        self.stream.synchronize()
        v = self.partial.cpu().numpy()
        return {"ll": float(v[0]), "w": v[1:].copy()}


class ShardedObjective:
    """The M-step objective over this rank's shard of the selected windows with ONE all-reduce per evaluation:
    [f(lambda), f(lambda + eps e_0), ...] (dim + 1 doubles) per gradient, left on the device by phyfoo_fe_values_device and
    summed over the ranks on the same stream.  Every rank ends up with the same f and g."""

    def __init__(self, fe, device):
        import torch
        self.fe = fe
        self.values = torch.zeros(64, dtype=torch.float64, device=f"cuda:{device}")
        self.stream = torch.cuda.Stream(device=device)

    def grad(self, position, lam, eps=1e-4):
        import torch
        lam = np.ascontiguousarray(lam, np.float64)
        n = lam.size + 1
        if n > self.values.numel():
            self.values = torch.zeros(n, dtype=torch.float64, device=self.values.device)
        with torch.cuda.stream(self.stream):
            self.fe.values_device(position, lam, eps, self.values.data_ptr(), self.stream.cuda_stream)
            v = self.values[:n]
            allreduce_sum_(v)
        self.stream.synchronize()
        v = v.cpu().numpy()
        return float(v[0]), (v[1:] - v[0]) / eps

    def eval(self, position, lam):
        import torch
        lam = np.ascontiguousarray(lam, np.float64)
        with torch.cuda.stream(self.stream):
            self.fe.values_device(position, lam, 0.0, self.values.data_ptr(), self.stream.cuda_stream)
            v = self.values[:1]
            allreduce_sum_(v)
        self.stream.synchronize()
        return float(v.cpu().numpy()[0])


def allreduce_counts_(counts, entropy):
    """Sufficient statistics of the M-step objective summed over the ranks: one all-reduce of W*E + 1 doubles, after which
    every rank optimises on the host without further communication.  counts: numpy [W][E]; returns (counts, entropy)."""
    d = _dist()
    if not d or d.get_world_size() == 1:
        return counts, entropy
    import torch
    t = torch.from_numpy(np.concatenate([np.asarray(counts, np.float64).ravel(), [entropy]]))
    if d.get_backend() == "nccl":
        t = t.cuda()
    d.all_reduce(t, op=d.ReduceOp.SUM)
    t = t.cpu().numpy()
    return t[:-1].reshape(np.asarray(counts).shape), float(t[-1])
