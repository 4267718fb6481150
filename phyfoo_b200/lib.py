"""Loader for libphyfoo.so (the C ABI of include/phyfoo.h) and its build recipe.

The library is the product: there is no Python or CPU implementation behind these calls.  If the
shared object is missing or no CUDA device is present, everything here raises.
"""
import ctypes as C
import os
import shutil
import subprocess

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.path.join(_HERE, "libphyfoo.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-ldl"]

_LIB = None

# every symbol include/phyfoo.h declares (tests check that the .so exports each one)
SYMBOLS = [
    "phyfoo_init", "phyfoo_destroy", "phyfoo_last_error", "phyfoo_device_info", "phyfoo_fp64_peak",
    "phyfoo_dataset_create", "phyfoo_dataset_create_device", "phyfoo_dataset_packed", "phyfoo_dataset_columns",
    "phyfoo_dataset_windows", "phyfoo_dataset_patterns", "phyfoo_dataset_free",
    "phyfoo_model_create", "phyfoo_model_set_pi", "phyfoo_model_get_pi", "phyfoo_model_set_branch",
    "phyfoo_model_set_mode", "phyfoo_model_is_degenerate", "phyfoo_model_free",
    "phyfoo_score_windows", "phyfoo_bg_columns", "phyfoo_bg_columns_order1", "phyfoo_zoops_estep", "phyfoo_zoops_estep_device",
    "phyfoo_zoops_estep_multi", "phyfoo_select_multi",
    "phyfoo_ab_score_windows", "phyfoo_ab_bg_columns", "phyfoo_ab_train", "phyfoo_ab_table_size", "phyfoo_ab_score_windows_order", "phyfoo_ab_bg_columns_order", "phyfoo_ab_train_order", "phyfoo_ab_bg_train_order", "phyfoo_scan_hits", "phyfoo_score_order_stats", "phyfoo_select", "phyfoo_fe_prepare", "phyfoo_fe_prepare_all", "phyfoo_fe_eval", "phyfoo_fe_grad",
    "phyfoo_fe_values_device", "phyfoo_fe_suffstats", "phyfoo_fe_size", "phyfoo_fe_free",
    "phyfoo_set_option", "phyfoo_last_score_path", "phyfoo_last_kernel_times", "phyfoo_host_alloc", "phyfoo_host_free",
    "phyfoo_last_launches", "phyfoo_stream", "phyfoo_synchronize", "phyfoo_estep_kernel_ms", "phyfoo_estep_sweeps",
    "phyfoo_init_multi", "phyfoo_destroy_multi", "phyfoo_multi_last_error", "phyfoo_multi_size", "phyfoo_multi_ctx",
    "phyfoo_multi_set_option", "phyfoo_multi_dataset_create", "phyfoo_multi_dataset_bounds", "phyfoo_multi_dataset_free",
    "phyfoo_multi_model_create", "phyfoo_multi_model_set_pi", "phyfoo_multi_model_get_pi", "phyfoo_multi_model_set_branch",
    "phyfoo_multi_model_set_mode", "phyfoo_multi_model_free", "phyfoo_multi_zoops_estep", "phyfoo_multi_fe_prepare_selected",
    "phyfoo_multi_fe_eval", "phyfoo_multi_fe_grad", "phyfoo_multi_fe_size", "phyfoo_multi_fe_free",
    "phyfoo_ss_create", "phyfoo_ss_create_host", "phyfoo_ss_eval", "phyfoo_ss_grad", "phyfoo_ss_eval_branches", "phyfoo_ss_get_pi",
    "phyfoo_ss_value", "phyfoo_ss_evaluations", "phyfoo_ss_optimize", "phyfoo_ss_free", "phyfoo_multi_ss_create",
]


def _sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".hpp"))] + \
           [os.path.join(os.path.dirname(_HERE), "include", "phyfoo.h")]


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> phyfoo_b200/libphyfoo.so (in-tree)."""
    if not force and os.path.exists(SO_PATH):
        newest = max(os.path.getmtime(s) for s in _sources())
        if os.path.getmtime(SO_PATH) >= newest:
            return SO_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(SO_PATH):
            return SO_PATH   # GPU box without a toolchain mismatch: use the prebuilt object
        raise RuntimeError("nvcc not found and libphyfoo.so is not built")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH, os.path.join(CSRC, "phyfoo.cu")]
    subprocess.check_call(cmd)
    return SO_PATH


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise RuntimeError(f"{SO_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(phyfoo_b200 has no CPU fallback)")
    L = C.CDLL(SO_PATH)
    vp, dp, ip32 = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    i64p, u8p, u32p = C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
    sp = C.POINTER(_abi.Stats)
    L.phyfoo_init.argtypes = [C.c_int, C.POINTER(vp)]
    L.phyfoo_destroy.argtypes = [vp]
    L.phyfoo_destroy.restype = None
    L.phyfoo_last_error.argtypes = [vp]
    L.phyfoo_last_error.restype = C.c_char_p
    L.phyfoo_device_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), i64p]
    L.phyfoo_fp64_peak.argtypes = [vp, dp]
    L.phyfoo_dataset_create.argtypes = [vp, C.c_int32, C.c_int32, i64p, u8p, C.POINTER(vp)]
    L.phyfoo_dataset_create_device.argtypes = [vp, C.c_int32, C.c_int32, i64p, vp, C.POINTER(vp)]
    L.phyfoo_dataset_packed.argtypes = [vp, vp, u32p, u32p]
    L.phyfoo_dataset_columns.argtypes = [vp]
    L.phyfoo_dataset_columns.restype = C.c_int64
    L.phyfoo_dataset_windows.argtypes = [vp, C.c_int32]
    L.phyfoo_dataset_windows.restype = C.c_int64
    L.phyfoo_dataset_patterns.argtypes = [vp, vp]
    L.phyfoo_dataset_patterns.restype = C.c_int64
    L.phyfoo_dataset_free.argtypes = [vp, vp]
    L.phyfoo_dataset_free.restype = None
    L.phyfoo_model_create.argtypes = [vp, C.POINTER(_abi.ModelDesc), C.POINTER(vp)]
    L.phyfoo_model_set_pi.argtypes = [vp, vp, C.c_int32, dp, C.c_int32]
    L.phyfoo_model_get_pi.argtypes = [vp, vp, C.c_int32, dp, C.c_int32]
    L.phyfoo_model_set_branch.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_double]
    L.phyfoo_model_set_mode.argtypes = [vp, vp, C.c_int32]
    L.phyfoo_model_is_degenerate.argtypes = [vp]
    L.phyfoo_model_free.argtypes = [vp, vp]
    L.phyfoo_model_free.restype = None
    L.phyfoo_score_windows.argtypes = [vp, vp, vp, C.c_int32, dp, ip32, sp]
    L.phyfoo_bg_columns.argtypes = [vp, vp, vp, dp, ip32, sp]
    L.phyfoo_bg_columns_order1.argtypes = [vp, vp, vp, dp, dp]
    L.phyfoo_zoops_estep.argtypes = [vp, vp, vp, vp, dp, dp, dp, dp, dp, dp, dp, ip32, u8p, sp]
    L.phyfoo_zoops_estep_device.argtypes = [vp, vp, vp, vp, dp, dp, vp, vp, vp]
    L.phyfoo_zoops_estep_multi.argtypes = [vp, vp, C.c_int32, C.POINTER(vp), vp, dp, dp, C.POINTER(dp), dp, dp, ip32, ip32, sp]
    L.phyfoo_select_multi.argtypes = [vp, vp, C.c_int32, C.c_double, C.c_double, ip32, ip32, dp, C.c_int64, i64p]
    L.phyfoo_ab_score_windows.argtypes = [vp, vp, C.c_int32, dp, C.c_int32, dp]
    L.phyfoo_ab_bg_columns.argtypes = [vp, vp, dp, dp]
    L.phyfoo_ab_train.argtypes = [vp, vp, C.c_int32, C.c_int64, ip32, ip32, u8p, dp, C.c_double, dp]
    L.phyfoo_ab_table_size.argtypes = [C.c_int32, C.c_int32]
    L.phyfoo_ab_score_windows_order.argtypes = [vp, vp, C.c_int32, C.c_int32, dp, C.c_int32, dp]
    L.phyfoo_ab_bg_columns_order.argtypes = [vp, vp, C.c_int32, dp, dp]
    L.phyfoo_ab_train_order.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int64, ip32, ip32, u8p, dp, C.c_double, dp]
    L.phyfoo_ab_bg_train_order.argtypes = [vp, vp, C.c_int32, dp, C.c_double, dp]
    L.phyfoo_scan_hits.argtypes = [vp, vp, vp, C.c_int32, C.c_double, ip32, ip32, dp, ip32]
    L.phyfoo_score_order_stats.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, i64p, dp, i64p]
    L.phyfoo_select.argtypes = [vp, vp, C.c_int32, dp, C.c_double, C.c_double, ip32, ip32, dp, C.c_int64, i64p]
    L.phyfoo_fe_prepare.argtypes = [vp, vp, vp, C.c_int64, ip32, ip32, u8p, dp, C.POINTER(vp)]
    L.phyfoo_fe_prepare_all.argtypes = [vp, vp, vp, dp, C.POINTER(vp)]
    L.phyfoo_fe_eval.argtypes = [vp, vp, C.c_int32, dp, C.c_int32, dp]
    L.phyfoo_fe_grad.argtypes = [vp, vp, C.c_int32, dp, C.c_int32, C.c_double, dp, dp]
    L.phyfoo_fe_values_device.argtypes = [vp, vp, C.c_int32, dp, C.c_int32, C.c_double, vp, vp]
    L.phyfoo_fe_suffstats.argtypes = [vp, vp, dp, dp]
    L.phyfoo_fe_size.argtypes = [vp]
    L.phyfoo_fe_size.restype = C.c_int64
    L.phyfoo_fe_free.argtypes = [vp, vp]
    L.phyfoo_fe_free.restype = None
    L.phyfoo_last_launches.argtypes = [vp]
    L.phyfoo_host_alloc.argtypes = [vp, C.c_int64]
    L.phyfoo_host_alloc.restype = vp
    L.phyfoo_host_free.argtypes = [vp, vp]
    L.phyfoo_host_free.restype = None
    L.phyfoo_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.phyfoo_last_score_path.argtypes = [vp]
    L.phyfoo_last_kernel_times.argtypes = [vp, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_float)]
    L.phyfoo_stream.argtypes = [vp]
    L.phyfoo_stream.restype = vp
    L.phyfoo_synchronize.argtypes = [vp]
    fp = C.POINTER(C.c_float)
    L.phyfoo_estep_kernel_ms.argtypes = [vp, fp, fp, fp]
    L.phyfoo_estep_sweeps.argtypes = [vp, i64p, i64p]
    # ---- sufficient-statistics objective on the host inside the library (csrc/suffstat_host.hpp)
    L.phyfoo_ss_create.argtypes = [vp, vp, C.POINTER(vp)]
    L.phyfoo_ss_create_host.argtypes = [C.POINTER(_abi.ModelDesc), dp, C.c_double, C.POINTER(vp)]
    L.phyfoo_ss_eval.argtypes = [vp, C.c_int32, dp, C.c_int32, dp]
    L.phyfoo_ss_grad.argtypes = [vp, C.c_int32, dp, C.c_int32, C.c_double, dp, dp]
    L.phyfoo_ss_eval_branches.argtypes = [vp, C.c_int32, dp, dp]
    L.phyfoo_ss_get_pi.argtypes = [vp, C.c_int32, dp, C.c_int32]
    L.phyfoo_ss_value.argtypes = [vp]
    L.phyfoo_ss_value.restype = C.c_double
    L.phyfoo_ss_evaluations.argtypes = [vp]
    L.phyfoo_ss_evaluations.restype = C.c_int64
    L.phyfoo_ss_optimize.argtypes = [vp, ip32, C.c_int32, dp, dp, i64p]
    L.phyfoo_ss_free.argtypes = [vp]
    L.phyfoo_ss_free.restype = None
    L.phyfoo_multi_ss_create.argtypes = [vp, vp, C.POINTER(vp)]
    # ---- several GPUs behind one caller (csrc/multi.cuh)
    L.phyfoo_init_multi.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    L.phyfoo_destroy_multi.argtypes = [vp]
    L.phyfoo_destroy_multi.restype = None
    L.phyfoo_multi_last_error.argtypes = [vp]
    L.phyfoo_multi_last_error.restype = C.c_char_p
    L.phyfoo_multi_size.argtypes = [vp]
    L.phyfoo_multi_ctx.argtypes = [vp, C.c_int]
    L.phyfoo_multi_ctx.restype = vp
    L.phyfoo_multi_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.phyfoo_multi_dataset_create.argtypes = [vp, C.c_int32, C.c_int32, i64p, u8p, C.POINTER(vp)]
    L.phyfoo_multi_dataset_bounds.argtypes = [vp, ip32]
    L.phyfoo_multi_dataset_free.argtypes = [vp, vp]
    L.phyfoo_multi_dataset_free.restype = None
    L.phyfoo_multi_model_create.argtypes = [vp, C.POINTER(_abi.ModelDesc), C.POINTER(vp)]
    L.phyfoo_multi_model_set_pi.argtypes = [vp, vp, C.c_int32, dp, C.c_int32]
    L.phyfoo_multi_model_get_pi.argtypes = [vp, vp, C.c_int32, dp, C.c_int32]
    L.phyfoo_multi_model_set_branch.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_double]
    L.phyfoo_multi_model_set_mode.argtypes = [vp, vp, C.c_int32]
    L.phyfoo_multi_model_free.argtypes = [vp, vp]
    L.phyfoo_multi_model_free.restype = None
    L.phyfoo_multi_zoops_estep.argtypes = [vp, vp, vp, vp, dp, dp, dp, dp, dp, dp, dp, ip32, u8p, sp]
    L.phyfoo_multi_fe_prepare_selected.argtypes = [vp, vp, vp, dp, C.c_double, C.c_double, C.POINTER(vp), i64p]
    L.phyfoo_multi_fe_eval.argtypes = [vp, vp, C.c_int32, dp, C.c_int32, dp]
    L.phyfoo_multi_fe_grad.argtypes = [vp, vp, C.c_int32, dp, C.c_int32, C.c_double, dp, dp]
    L.phyfoo_multi_fe_size.argtypes = [vp]
    L.phyfoo_multi_fe_size.restype = C.c_int64
    L.phyfoo_multi_fe_free.argtypes = [vp, vp]
    L.phyfoo_multi_fe_free.restype = None
    _LIB = L
    return L


class PhyfooError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libphyfoo error {code}: {msg}")
        self.code = code


class IllegalArgumentException(PhyfooError, ValueError):
    """PHYFOO_EINVAL: what the reference's check() methods throw (models/PhyloBayesModel.java:408-424)."""


def raise_for(code, ctx_handle):
    if code == _abi.OK:
        return
    msg = lib().phyfoo_last_error(ctx_handle)
    msg = msg.decode() if msg else ""
    if code == _abi.EINVAL:
        raise IllegalArgumentException(code, msg)
    raise PhyfooError(code, msg)
