"""phy_exp / phy_log (phyfoo_b200/csrc/fp64_math.cuh, the transcendentals of the mean-field kernels) against the C
library on the host: error in ulps over the ranges the kernels use, and the special cases."""
import ctypes as C

import numpy as np

from test_emul_host import emul  # noqa: F401  (fixture)


def _call(fn, x):
    x = np.ascontiguousarray(x, np.float64)
    out = np.zeros_like(x)
    fn(x.ctypes.data_as(C.POINTER(C.c_double)), x.size, out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def _ulps(got, ref):
    return np.abs(got - ref) / np.spacing(np.abs(ref))


def test_exp_accuracy_and_special_cases(emul):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-80, 5, 400000), rng.uniform(-745, 709, 200000), rng.normal(0, 1e-3, 50000),
                        np.linspace(-1, 1, 20001), [0.0, -0.0, 1.0, -1.0, 709.0, -708.0]])
    got = _call(emul.emul_exp, x)
    ref = np.exp(x.astype(np.longdouble)).astype(np.float64)
    normal = ref > 2.3e-308
    assert _ulps(got[normal], ref[normal]).max() <= 1.0
    assert np.abs(got[~normal] - ref[~normal]).max() <= 5e-324 * 2       # denormal results: within one denormal step
    sp = _call(emul.emul_exp, [-np.inf, np.inf, np.nan, -746.0, -800.0, 710.0, -745.0])
    assert sp[0] == 0 and np.isinf(sp[1]) and np.isnan(sp[2]) and sp[3] == 0 and sp[4] == 0 and np.isinf(sp[5])
    assert sp[6] == np.exp(-745.0)


def test_log_accuracy_and_special_cases(emul):
    rng = np.random.default_rng(2)
    z = np.concatenate([np.exp(rng.uniform(-700, 3, 400000)), rng.uniform(0.5, 2.0, 200000), np.linspace(0.99, 1.01, 20001),
                        rng.uniform(0, 4, 100000), [1.0, 4.0, 2.0, 0.703125, 1.40625, 5e-324, 2.2250738585072014e-308, 1e-310]])
    got = _call(emul.emul_log, z)
    ref = np.log(z.astype(np.longdouble)).astype(np.float64)
    # absolute accuracy is what the free energy needs; near z = 1 the result is tiny and ulps are relative to it
    err = np.abs(got - ref)
    assert (err <= np.maximum(1.0 * np.spacing(np.abs(ref)), 2.3e-16)).all(), (err.max(), z[np.argmax(err)])
    assert abs(_call(emul.emul_log, [1.0])[0]) < 1e-17          # not exactly 0: table + polynomial, 2e-18 off
    sp = _call(emul.emul_log, [0.0, -0.0, np.inf, -1.0, np.nan])
    assert np.isneginf(sp[0]) and np.isneginf(sp[1]) and np.isinf(sp[2]) and sp[2] > 0 and np.isnan(sp[3]) and np.isnan(sp[4])
