"""ctypes view of include/phyfoo.h (struct layouts and constants only; no logic)."""
import ctypes as C

import numpy as np

OK, EINVAL, ENODEVICE, ECUDA, ENOMEM, EUNSUPPORTED = 0, -1, -2, -3, -4, -5
MODEL_FG, MODEL_BG = 0, 1
EVOL_F81ALPHA, EVOL_F81BETA, EVOL_HKY_AS_IMPLEMENTED, EVOL_HKY_RATIO = 0, 1, 2, 3
MODE_MEANFIELD, MODE_EXACT = 0, 1


class ModelDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("width", C.c_int32), ("order", C.c_int32), ("n_nodes", C.c_int32),
        ("parent", C.POINTER(C.c_int32)), ("leaf_species", C.POINTER(C.c_int32)),
        ("evol", C.c_int32), ("hky_ratio", C.c_double),
        ("branch", C.POINTER(C.c_double)), ("pi", C.POINTER(C.c_double)),
        ("mode", C.c_int32), ("allow_independent_leaf", C.c_int32),
        ("max_sweeps", C.c_int32), ("min_sweeps", C.c_int32), ("tol", C.c_double),
    ]


class Stats(C.Structure):
    _fields_ = [("n_windows", C.c_int64), ("n_columns", C.c_int64), ("sweeps_fg", C.c_int64),
                ("sweeps_bg", C.c_int64), ("kernel_launches", C.c_int32), ("gpu_ms", C.c_float)]


def make_desc(kind, width, order, tree, branch, pi_list, evol=EVOL_F81ALPHA, mode=MODE_MEANFIELD, hky_ratio=0.0,
              max_sweeps=0, min_sweeps=0, tol=0.0):
    """Returns (desc, keepalive): keepalive holds the numpy arrays the struct points into."""
    parent = np.ascontiguousarray(tree.parent, np.int32)
    leaf = np.ascontiguousarray(tree.leaf_species, np.int32)
    branch = np.ascontiguousarray(branch, np.float64)
    pi = np.ascontiguousarray(np.concatenate([np.asarray(p, np.float64).ravel() for p in pi_list]))
    d = ModelDesc(kind, width, order, parent.size, parent.ctypes.data_as(C.POINTER(C.c_int32)),
                  leaf.ctypes.data_as(C.POINTER(C.c_int32)), evol, hky_ratio,
                  branch.ctypes.data_as(C.POINTER(C.c_double)), pi.ctypes.data_as(C.POINTER(C.c_double)),
                  mode, 1, int(max_sweeps), int(min_sweeps), float(tol))
    return d, (parent, leaf, branch, pi)
