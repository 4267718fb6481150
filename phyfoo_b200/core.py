"""Thin object wrappers over the C ABI handles (context, packed dataset, device model, FE set).

numpy arrays in, numpy arrays out; all arithmetic happens in libphyfoo.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _abi
from .lib import lib, raise_for

_dp = C.POINTER(C.c_double)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class Context:
    """phyfoo_ctx: one GPU, one caller thread at a time."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        rc = lib().phyfoo_init(int(device), C.byref(self.h))
        if rc != _abi.OK:
            raise_for(rc, None)
        self.device = int(device)

    def close(self):
        if self.h:
            lib().phyfoo_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        raise_for(lib().phyfoo_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)), self.h)
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value}

    def fp64_peak(self):
        v = C.c_double()
        raise_for(lib().phyfoo_fp64_peak(self.h, C.byref(v)), self.h)
        return v.value

    def stream(self):
        return lib().phyfoo_stream(self.h)

    def synchronize(self):
        raise_for(lib().phyfoo_synchronize(self.h), self.h)

    def last_launches(self):
        return lib().phyfoo_last_launches(self.h)

    def set_option(self, name, value):
        raise_for(lib().phyfoo_set_option(self.h, name.encode(), int(value)), self.h)

    def jit_note(self):
        """Why the last request for a run-time compiled order-0 kernel fell back to the precompiled one ('' if it did not)."""
        return lib().phyfoo_jit_note(self.h).decode()

    def pinned_empty(self, shape, dtype=np.float64):
        """numpy array in page-locked host memory (phyfoo_host_alloc); copies to/from it run at full PCIe speed.
        The memory is released when the array (and every view of it) is garbage collected."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = lib().phyfoo_host_alloc(self.h, n)
        if not ptr:
            raise MemoryError("phyfoo_host_alloc failed")
        buf = (C.c_char * max(n, 1)).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        import weakref
        weakref.finalize(buf, lib().phyfoo_host_free, None, ptr)
        return arr

    def last_kernel_times(self):
        """[(kernel name, device ms)] of the most recent call on this context, in launch order."""
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = lib().phyfoo_last_kernel_times(self.h, 64, names, ms)
        if n < 0:
            raise_for(n, self.h)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def last_score_path(self):
        """0 = direct kernel, 1 = pattern table, 2 = order-1 network."""
        return lib().phyfoo_last_score_path(self.h)

    def estep_kernel_ms(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        raise_for(lib().phyfoo_estep_kernel_ms(self.h, C.byref(a), C.byref(b), C.byref(c)), self.h)
        return {"bg_ms": a.value, "fg_ms": b.value, "zoops_ms": c.value}

    def estep_sweeps(self):
        a, b = C.c_int64(), C.c_int64()
        raise_for(lib().phyfoo_estep_sweeps(self.h, C.byref(a), C.byref(b)), self.h)
        return a.value, b.value


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class Dataset:
    """phyfoo_dataset: packed columns resident on the device."""

    def __init__(self, ctx, col_offsets, codes=None, n_species=None, codes_device_ptr=None):
        self.ctx = ctx
        self.col_offsets = np.ascontiguousarray(col_offsets, np.int64)
        self.n_aln = self.col_offsets.size - 1
        self.h = C.c_void_p()
        if codes_device_ptr is not None:
            self.S = int(n_species)
            rc = lib().phyfoo_dataset_create_device(ctx.h, self.n_aln, self.S, _p(self.col_offsets, C.c_int64),
                                                    C.c_void_p(codes_device_ptr), C.byref(self.h))
        else:
            codes = np.ascontiguousarray(codes, np.uint8)
            if codes.ndim != 2:
                raise ValueError("codes must be [columns][species]")
            self.S = codes.shape[1]
            if codes.shape[0] != self.col_offsets[-1]:
                raise ValueError("col_offsets[-1] must equal the number of columns")
            rc = lib().phyfoo_dataset_create(ctx.h, self.n_aln, self.S, _p(self.col_offsets, C.c_int64),
                                             _p(codes, C.c_uint8), C.byref(self.h))
        raise_for(rc, ctx.h)
        self.n_cols = int(self.col_offsets[-1])

    def free(self):
        if self.h:
            lib().phyfoo_dataset_free(self.ctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def lengths(self):
        return np.diff(self.col_offsets)

    def windows(self, width):
        return int(lib().phyfoo_dataset_windows(self.h, int(width)))

    def patterns(self):
        """Distinct column patterns (forward + complemented), or -1 when the pattern table does not apply."""
        return int(lib().phyfoo_dataset_patterns(self.ctx.h, self.h))

    def window_offsets(self, width):
        return np.concatenate([[0], np.cumsum(np.maximum(self.lengths() - width + 1, 0))]).astype(np.int64)

    def packed(self):
        nb, ng = (self.S + 15) // 16, (self.S + 31) // 32
        bases = np.zeros((nb, self.n_cols), np.uint32)
        gaps = np.zeros((ng, self.n_cols), np.uint32)
        raise_for(lib().phyfoo_dataset_packed(self.ctx.h, self.h, _p(bases, C.c_uint32), _p(gaps, C.c_uint32)), self.ctx.h)
        return bases, gaps


class DeviceModel:
    """phyfoo_model: tree + parameters of one PhyloBayesModel / PhyloBackground."""

    def __init__(self, ctx, kind, width, order, tree, branch, pi_list, evol=_abi.EVOL_F81ALPHA,
                 mode=_abi.MODE_MEANFIELD, hky_ratio=0.0, max_sweeps=0, min_sweeps=0, tol=0.0):
        self.ctx = ctx
        desc, keep = _abi.make_desc(kind, width, order, tree, branch, pi_list, evol, mode, hky_ratio, max_sweeps, min_sweeps, tol)
        self.h = C.c_void_p()
        raise_for(lib().phyfoo_model_create(ctx.h, C.byref(desc), C.byref(self.h)), ctx.h)
        del keep
        self.kind, self.order = kind, order
        self.W = width if kind == _abi.MODEL_FG else order + 1

    def free(self):
        if self.h:
            lib().phyfoo_model_free(self.ctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def set_pi(self, position, pi):
        pi = np.ascontiguousarray(pi, np.float64).ravel()
        raise_for(lib().phyfoo_model_set_pi(self.ctx.h, self.h, int(position), _p(pi, C.c_double), pi.size), self.ctx.h)

    def get_pi(self, position, n):
        out = np.zeros(n)
        raise_for(lib().phyfoo_model_get_pi(self.ctx.h, self.h, int(position), _p(out, C.c_double), n), self.ctx.h)
        return out

    def set_branch(self, position, node, length):
        raise_for(lib().phyfoo_model_set_branch(self.ctx.h, self.h, int(position), int(node), float(length)), self.ctx.h)

    def set_mode(self, mode):
        raise_for(lib().phyfoo_model_set_mode(self.ctx.h, self.h, int(mode)), self.ctx.h)


def _stats_dict(st):
    return {"n_windows": st.n_windows, "n_columns": st.n_columns, "sweeps_fg": st.sweeps_fg, "sweeps_bg": st.sweeps_bg,
            "kernel_launches": st.kernel_launches, "gpu_ms": st.gpu_ms}


def score_windows(ctx, ds, model, strand=0, want_sweeps=False):
    n = ds.windows(model.W)
    out = np.zeros(n)
    sw = np.zeros(n, np.int32) if want_sweeps else None
    st = _abi.Stats()
    raise_for(lib().phyfoo_score_windows(ctx.h, ds.h, model.h, int(strand), _p(out, C.c_double), _p(sw, C.c_int32),
                                         C.byref(st)), ctx.h)
    return out, sw, _stats_dict(st)


def bg_columns(ctx, ds, model, want_sweeps=False):
    out = np.zeros(ds.n_cols)
    sw = np.zeros(ds.n_cols, np.int32) if want_sweeps else None
    st = _abi.Stats()
    raise_for(lib().phyfoo_bg_columns(ctx.h, ds.h, model.h, _p(out, C.c_double), _p(sw, C.c_int32), C.byref(st)), ctx.h)
    return out, sw, _stats_dict(st)


def zoops_estep(ctx, ds, fg, bg, logw, strand_logw=None, aln_weights=None, want_gamma=True, want_profile=False,
                want_argmax=True, out=None):
    """out: optional dict of preallocated result arrays (gamma / profile [windows] float64, argmax_off [n_aln] int32,
    argmax_strand [n_aln] uint8), e.g. from Context.pinned_empty, reused across calls."""
    n = ds.windows(fg.W)
    out = out or {}
    logw = np.ascontiguousarray(logw, np.float64)
    sl = None if strand_logw is None else np.ascontiguousarray(strand_logw, np.float64)
    aw = None if aln_weights is None else np.ascontiguousarray(aln_weights, np.float64)

    def buf(key, size, dtype):
        a = out.get(key)
        if a is None:
            return np.zeros(size, dtype)
        if a.dtype != dtype or a.size < size or not a.flags.c_contiguous:
            raise ValueError(f"out[{key!r}] must be a contiguous {np.dtype(dtype).name} array of at least {size} elements")
        return a.ravel()[:size]

    gamma = buf("gamma", n, np.float64) if want_gamma else None
    prof = buf("profile", n, np.float64) if want_profile else None
    w = np.zeros(2)
    ll = C.c_double()
    ao = buf("argmax_off", ds.n_aln, np.int32) if want_argmax else None
    astr = buf("argmax_strand", ds.n_aln, np.uint8) if want_argmax else None
    st = _abi.Stats()
    raise_for(lib().phyfoo_zoops_estep(ctx.h, ds.h, fg.h, bg.h, _p(logw, C.c_double), _p(sl, C.c_double),
                                       _p(aw, C.c_double), _p(gamma, C.c_double), _p(prof, C.c_double),
                                       _p(w, C.c_double), C.byref(ll), _p(ao, C.c_int32), _p(astr, C.c_uint8),
                                       C.byref(st)), ctx.h)
    return {"gamma": gamma, "profile": prof, "w": w, "ll": ll.value, "argmax_off": ao, "argmax_strand": astr,
            "stats": _stats_dict(st)}


def zoops_estep_device(ctx, ds, fg, bg, logw, strand_logw, aln_weights_ptr, partial_ptr, stream_ptr):
    """Asynchronous E-step that leaves [ll, w0, w1] (local sums) in device memory for an NCCL all-reduce."""
    logw = np.ascontiguousarray(logw, np.float64)
    sl = None if strand_logw is None else np.ascontiguousarray(strand_logw, np.float64)
    raise_for(lib().phyfoo_zoops_estep_device(ctx.h, ds.h, fg.h, bg.h, _p(logw, C.c_double), _p(sl, C.c_double),
                                              C.c_void_p(aln_weights_ptr) if aln_weights_ptr else None,
                                              C.c_void_p(partial_ptr), C.c_void_p(stream_ptr) if stream_ptr else None),
              ctx.h)


def select(ctx, ds, width, weights=None, t=0.8, q=0.0):
    """io/SequenceSpecificDataSelector.java:37-99 over all alignments.  weights = None: the gamma of the most
    recent zoops_estep on this context (still resident on the device).  Returns (aln, off, weight)."""
    w = None if weights is None else np.ascontiguousarray(weights, np.float64)
    n = C.c_int64(0)
    cap = max(4 * ds.n_aln, 1024)
    while True:
        sa, so, sw = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
        rc = lib().phyfoo_select(ctx.h, ds.h, int(width), _p(w, C.c_double), float(t), float(q), _p(sa, C.c_int32),
                                 _p(so, C.c_int32), _p(sw, C.c_double), cap, C.byref(n))
        if rc == _abi.ENOMEM and n.value > cap:
            cap = int(n.value)
            continue
        raise_for(rc, ctx.h)
        return sa[:n.value].copy(), so[:n.value].copy(), sw[:n.value].copy()


class FESet:
    """phyfoo_feset: the fixed-q free-energy objective of the M-step over a list of windows
    (optimizing/AbstractFreeEnergyOptimizer.java:95-114); q lives on the device."""

    def __init__(self, ctx, ds, model, sel_aln=None, sel_off=None, sel_weight=None, sel_strand=None, aln_weights=None):
        self.ctx, self.ds, self.model = ctx, ds, model
        self.h = C.c_void_p()
        if sel_aln is None:
            aw = None if aln_weights is None else np.ascontiguousarray(aln_weights, np.float64)
            rc = lib().phyfoo_fe_prepare_all(ctx.h, ds.h, model.h, _p(aw, C.c_double), C.byref(self.h))
        else:
            sa = np.ascontiguousarray(sel_aln, np.int32)
            so = np.ascontiguousarray(sel_off, np.int32)
            sw = np.ascontiguousarray(sel_weight, np.float64)
            ss = None if sel_strand is None else np.ascontiguousarray(sel_strand, np.uint8)
            rc = lib().phyfoo_fe_prepare(ctx.h, ds.h, model.h, sa.size, _p(sa, C.c_int32), _p(so, C.c_int32),
                                         _p(ss, C.c_uint8), _p(sw, C.c_double), C.byref(self.h))
        raise_for(rc, ctx.h)

    def free(self):
        if self.h:
            lib().phyfoo_fe_free(self.ctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __len__(self):
        return int(lib().phyfoo_fe_size(self.h))

    def eval(self, position, lam):
        """evaluateFunction(lambda): f = -sum_u w_u F(q_u; theta(lambda)); the model is left at lambda."""
        lam = np.ascontiguousarray(lam, np.float64)
        f = C.c_double()
        raise_for(lib().phyfoo_fe_eval(self.ctx.h, self.h, int(position), _p(lam, C.c_double), lam.size, C.byref(f)), self.ctx.h)
        return f.value

    def grad(self, position, lam, eps=1e-4):
        """evaluateGradientOfFunction(lambda): forward differences (jstacs NumericalDifferentiableFunction.java:64-84)."""
        lam = np.ascontiguousarray(lam, np.float64)
        f = C.c_double()
        g = np.zeros(lam.size)
        raise_for(lib().phyfoo_fe_grad(self.ctx.h, self.h, int(position), _p(lam, C.c_double), lam.size, float(eps),
                                       C.byref(f), _p(g, C.c_double)), self.ctx.h)
        return f.value, g

    def suffstats(self, W, E):
        """(counts [W][E], entropy): expected counts of the log-table entries per position, from one device pass
        (phyfoo_fe_suffstats; order 0).  f(theta) = sum(counts * log tables_theta) - entropy for any parameters."""
        counts = np.zeros((W, E))
        ent = C.c_double()
        raise_for(lib().phyfoo_fe_suffstats(self.ctx.h, self.h, _p(counts, C.c_double), C.byref(ent)), self.ctx.h)
        return counts, ent.value

    def values_device(self, position, lam, eps, values_ptr, stream_ptr):
        lam = np.ascontiguousarray(lam, np.float64)
        raise_for(lib().phyfoo_fe_values_device(self.ctx.h, self.h, int(position), _p(lam, C.c_double), lam.size, float(eps),
                                                C.c_void_p(values_ptr), C.c_void_p(stream_ptr) if stream_ptr else None),
                  self.ctx.h)


# ---------------------------------------------------------------------------------------------------------------------------
# several GPUs behind one caller (include/phyfoo.h "phyfoo_multi_*", csrc/multi.cuh): the library shards the sample over the
# listed devices and all-reduces the few doubles each consumer needs over NCCL; the arguments and results are those of the
# single-GPU calls, with per-window / per-alignment arrays in the order of the whole sample
# ---------------------------------------------------------------------------------------------------------------------------
def _mraise(code, mh):
    if code == _abi.OK:
        return
    msg = lib().phyfoo_multi_last_error(mh)
    from .lib import IllegalArgumentException, PhyfooError
    msg = msg.decode() if msg else ""
    raise (IllegalArgumentException if code == _abi.EINVAL else PhyfooError)(code, msg)


class MultiContext:
    """phyfoo_mctx: one caller, several GPUs (one worker thread and one NCCL rank per device inside the library)."""

    def __init__(self, devices):
        import os
        if "PHYFOO_NCCL_LIB" not in os.environ:
            try:                                     # the NCCL that ships with the torch wheel, if there is one
                import nvidia.nccl as _n
                cand = os.path.join(os.path.dirname(_n.__file__ or list(_n.__path__)[0]), "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    os.environ["PHYFOO_NCCL_LIB"] = cand
            except Exception:
                pass
        devices = [int(d) for d in devices]
        arr = (C.c_int * len(devices))(*devices)
        self.h = C.c_void_p()
        rc = lib().phyfoo_init_multi(arr, len(devices), C.byref(self.h))
        if rc != _abi.OK:
            _mraise(rc, None)
        self.devices = devices

    def close(self):
        if self.h:
            lib().phyfoo_destroy_multi(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return lib().phyfoo_multi_size(self.h)

    def set_option(self, name, value):
        _mraise(lib().phyfoo_multi_set_option(self.h, name.encode(), int(value)), self.h)


class MultiDataset:
    def __init__(self, mctx, col_offsets, codes):
        self.mctx = mctx
        self.col_offsets = np.ascontiguousarray(col_offsets, np.int64)
        codes = np.ascontiguousarray(codes, np.uint8)
        self.n_aln, self.S = self.col_offsets.size - 1, codes.shape[1]
        self.h = C.c_void_p()
        _mraise(lib().phyfoo_multi_dataset_create(mctx.h, self.n_aln, self.S, _p(self.col_offsets, C.c_int64), _p(codes, C.c_uint8),
                                                  C.byref(self.h)), mctx.h)
        b = np.zeros(len(mctx) + 1, np.int32)
        lib().phyfoo_multi_dataset_bounds(self.h, _p(b, C.c_int32))
        self.bounds = b

    def windows(self, width):
        return int(np.maximum(np.diff(self.col_offsets) - width + 1, 0).sum())

    def free(self):
        if self.h:
            lib().phyfoo_multi_dataset_free(self.mctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MultiModel:
    def __init__(self, mctx, kind, width, order, tree, branch, pi_list, evol=_abi.EVOL_F81ALPHA, mode=_abi.MODE_MEANFIELD,
                 hky_ratio=0.0):
        self.mctx = mctx
        desc, keep = _abi.make_desc(kind, width, order, tree, branch, pi_list, evol, mode, hky_ratio)
        self.h = C.c_void_p()
        _mraise(lib().phyfoo_multi_model_create(mctx.h, C.byref(desc), C.byref(self.h)), mctx.h)
        del keep
        self.W = width if kind == _abi.MODEL_FG else order + 1

    def set_pi(self, position, pi):
        pi = np.ascontiguousarray(pi, np.float64).ravel()
        _mraise(lib().phyfoo_multi_model_set_pi(self.mctx.h, self.h, int(position), _p(pi, C.c_double), pi.size), self.mctx.h)

    def get_pi(self, position, n):
        out = np.zeros(n)
        _mraise(lib().phyfoo_multi_model_get_pi(self.mctx.h, self.h, int(position), _p(out, C.c_double), n), self.mctx.h)
        return out

    def free(self):
        if self.h:
            lib().phyfoo_multi_model_free(self.mctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def multi_zoops_estep(mctx, mds, fg, bg, logw, strand_logw=None, aln_weights=None, want_gamma=True, want_argmax=True):
    n = mds.windows(fg.W)
    logw = np.ascontiguousarray(logw, np.float64)
    sl = None if strand_logw is None else np.ascontiguousarray(strand_logw, np.float64)
    aw = None if aln_weights is None else np.ascontiguousarray(aln_weights, np.float64)
    gamma = np.zeros(n) if want_gamma else None
    w, ll = np.zeros(2), C.c_double()
    ao = np.zeros(mds.n_aln, np.int32) if want_argmax else None
    astr = np.zeros(mds.n_aln, np.uint8) if want_argmax else None
    st = _abi.Stats()
    _mraise(lib().phyfoo_multi_zoops_estep(mctx.h, mds.h, fg.h, bg.h, _p(logw, C.c_double), _p(sl, C.c_double), _p(aw, C.c_double),
                                           _p(gamma, C.c_double), None, _p(w, C.c_double), C.byref(ll), _p(ao, C.c_int32),
                                           _p(astr, C.c_uint8), C.byref(st)), mctx.h)
    return {"gamma": gamma, "w": w, "ll": ll.value, "argmax_off": ao, "argmax_strand": astr, "stats": _stats_dict(st)}


class MultiFESet:
    """The M-step objective over all devices: select (on the resident gamma, or on `weights`) + fix q, then
    evaluateFunction / evaluateGradientOfFunction with one all-reduce each."""

    def __init__(self, mctx, mds, model, weights=None, t=0.8, q=0.0):
        self.mctx = mctx
        w = None if weights is None else np.ascontiguousarray(weights, np.float64)
        self.h = C.c_void_p()
        n = C.c_int64()
        _mraise(lib().phyfoo_multi_fe_prepare_selected(mctx.h, mds.h, model.h, _p(w, C.c_double), float(t), float(q), C.byref(self.h),
                                                       C.byref(n)), mctx.h)
        self.n_selected = n.value

    def __len__(self):
        return int(lib().phyfoo_multi_fe_size(self.h))

    def eval(self, position, lam):
        lam = np.ascontiguousarray(lam, np.float64)
        f = C.c_double()
        _mraise(lib().phyfoo_multi_fe_eval(self.mctx.h, self.h, int(position), _p(lam, C.c_double), lam.size, C.byref(f)), self.mctx.h)
        return f.value

    def grad(self, position, lam, eps=1e-4):
        lam = np.ascontiguousarray(lam, np.float64)
        f, g = C.c_double(), np.zeros(lam.size)
        _mraise(lib().phyfoo_multi_fe_grad(self.mctx.h, self.h, int(position), _p(lam, C.c_double), lam.size, float(eps), C.byref(f),
                                           _p(g, C.c_double)), self.mctx.h)
        return f.value, g

    def free(self):
        if self.h:
            lib().phyfoo_multi_fe_free(self.mctx.h, self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class SSObjective:
    """phyfoo_ssobj: the fixed-q objective on sufficient statistics, evaluated on the host inside the library (order 0).
    From an FESet (one device pass), a MultiFESet (counts all-reduced over the devices) or given counts."""

    def __init__(self, fe=None, desc=None, counts=None, entropy=0.0, background=None):
        """background = (ctx, dataset, device model of a background of order >= 1, per-alignment weights or None):
        phyfoo_ss_create_bg -- the mean fields of every (order + 1)-column window, fixed in one pass of the generic net kernel."""
        self.h = C.c_void_p()
        if background is not None:
            ctx, ds, model, w = background
            w = None if w is None else np.ascontiguousarray(w, np.float64)
            raise_for(lib().phyfoo_ss_create_bg(ctx.h, ds.h, model.h, _p(w, C.c_double), C.byref(self.h)), ctx.h)
        elif isinstance(fe, MultiFESet):
            _mraise(lib().phyfoo_multi_ss_create(fe.mctx.h, fe.h, C.byref(self.h)), fe.mctx.h)
        elif fe is not None:
            raise_for(lib().phyfoo_ss_create(fe.ctx.h, fe.h, C.byref(self.h)), fe.ctx.h)
        else:
            c = np.ascontiguousarray(counts, np.float64)
            rc = lib().phyfoo_ss_create_host(C.byref(desc), _p(c, C.c_double), float(entropy), C.byref(self.h))
            if rc != _abi.OK:
                raise ValueError(f"phyfoo_ss_create_host failed ({rc})")

    def free(self):
        if self.h:
            lib().phyfoo_ss_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def value(self):
        return lib().phyfoo_ss_value(self.h)

    def evaluations(self):
        return int(lib().phyfoo_ss_evaluations(self.h))

    def eval(self, position, lam):
        lam = np.ascontiguousarray(lam, np.float64)
        f = C.c_double()
        if lib().phyfoo_ss_eval(self.h, int(position), _p(lam, C.c_double), lam.size, C.byref(f)) != _abi.OK:
            raise ValueError("phyfoo_ss_eval: bad argument")
        return f.value

    def grad(self, position, lam, eps=1e-4):
        lam = np.ascontiguousarray(lam, np.float64)
        f, g = C.c_double(), np.zeros(lam.size)
        if lib().phyfoo_ss_grad(self.h, int(position), _p(lam, C.c_double), lam.size, float(eps), C.byref(f), _p(g, C.c_double)) != _abi.OK:
            raise ValueError("phyfoo_ss_grad: bad argument")
        return f.value, g

    def eval_branches(self, position, branch):
        b = np.ascontiguousarray(branch, np.float64)
        f = C.c_double()
        if lib().phyfoo_ss_eval_branches(self.h, int(position), _p(b, C.c_double), C.byref(f)) != _abi.OK:
            raise ValueError("phyfoo_ss_eval_branches: bad argument")
        return f.value

    def get_pi(self, position, n=4):
        out = np.zeros(int(n))
        if lib().phyfoo_ss_get_pi(self.h, int(position), _p(out, C.c_double), int(n)) != _abi.OK:
            raise ValueError("phyfoo_ss_get_pi: bad argument (n must be 4 x the context rows of the position)")
        return out

    def optimize(self, order):
        """One M-step of the motif model: the positions in `order`, each by the library's BFGS.  Returns (f_before, f_after,
        evaluations)."""
        o = np.ascontiguousarray(order, np.int32)
        fb, fa, ev = C.c_double(), C.c_double(), C.c_int64()
        if lib().phyfoo_ss_optimize(self.h, _p(o, C.c_int32), o.size, C.byref(fb), C.byref(fa), C.byref(ev)) != _abi.OK:
            raise ValueError("phyfoo_ss_optimize: bad argument")
        return fb.value, fa.value, ev.value


# ---- ZOOPS mixture of several motif models (phyfoo_zoops_estep_multi) ----
def zoops_estep_multi(ctx, ds, fgs, bg, logw, aln_weights=None, want_gamma=True, want_argmax=True):
    """fgs: list of K DeviceModels; logw[K + 1] (motifs, then "no motif").  Returns gamma (list of K arrays), w[K + 1], ll,
    argmax_motif / argmax_off per alignment."""
    K = len(fgs)
    logw = np.ascontiguousarray(logw, np.float64)
    if logw.size != K + 1:
        raise ValueError("logw must hold one entry per motif plus one for 'no motif'")
    aw = None if aln_weights is None else np.ascontiguousarray(aln_weights, np.float64)
    handles = (C.c_void_p * K)(*[m.h for m in fgs])
    gam = [np.zeros(ds.windows(m.W)) for m in fgs] if want_gamma else None
    gptr = (C.POINTER(C.c_double) * K)(*[_p(g, C.c_double) for g in gam]) if want_gamma else None
    w, ll = np.zeros(K + 1), C.c_double()
    am = np.zeros(ds.n_aln, np.int32) if want_argmax else None
    ao = np.zeros(ds.n_aln, np.int32) if want_argmax else None
    st = _abi.Stats()
    raise_for(lib().phyfoo_zoops_estep_multi(ctx.h, ds.h, K, handles, bg.h, _p(logw, C.c_double), _p(aw, C.c_double), gptr, _p(w, C.c_double),
                                             C.byref(ll), _p(am, C.c_int32), _p(ao, C.c_int32), C.byref(st)), ctx.h)
    return {"gamma": gam, "w": w, "ll": ll.value, "argmax_motif": am, "argmax_off": ao, "stats": _stats_dict(st)}


def select_multi(ctx, ds, motif, t=0.8, q=0.0):
    """SequenceSpecificDataSelector on the gamma of component `motif` of the last zoops_estep_multi (resident on the device)."""
    n = C.c_int64(0)
    cap = max(4 * ds.n_aln, 1024)
    while True:
        sa, so, sw = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
        rc = lib().phyfoo_select_multi(ctx.h, ds.h, int(motif), float(t), float(q), _p(sa, C.c_int32), _p(so, C.c_int32), _p(sw, C.c_double),
                                       cap, C.byref(n))
        if rc == _abi.ENOMEM and n.value > cap:
            cap = int(n.value)
            continue
        raise_for(rc, ctx.h)
        return sa[:n.value].copy(), so[:n.value].copy(), sw[:n.value].copy()
