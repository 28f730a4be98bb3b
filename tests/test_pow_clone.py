"""The product's pow()/exp() clone (ode_solvers_b200/csrc/ode_pow.h), compiled for the HOST,
must be bit-identical to the libm that the reference's f64::powf / f64::exp call
(src/controller.rs:53-54, src/dopri5.rs:236, src/dop853.rs:244). The device build of the
same header is checked by tests/test_gpu_pow.py."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    oracle.lib()  # (re)builds oracle/ if needed
    l = C.CDLL(os.path.join(ROOT, "oracle", "libode_powcheck.so"))
    l.ode_powcheck_pow.restype = C.c_int64
    l.ode_powcheck_exp.restype = C.c_int64
    return l


def pow_mismatches(x, y):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(np.broadcast_to(y, x.shape), dtype=np.float64)
    return _lib().ode_powcheck_pow(C.c_int64(x.size), C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), None,
                                   None)


def exp_mismatches(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return _lib().ode_powcheck_exp(C.c_int64(x.size), C.c_void_p(x.ctypes.data), None, None)


def test_tables_match_running_libm():
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "extract_libm_tables.py"), "--check"], check=True,
                   stdout=subprocess.DEVNULL)


def test_pow_controller_domain():
    rng = np.random.default_rng(11)
    n = 500_000
    for a in (0.2 - 0.04 * 0.75, 0.125, -0.04, 0.2, 0.125 - 0.1 * 0.2):
        assert pow_mismatches(10 ** rng.uniform(-8, 2, n), a) == 0


def test_pow_full_range_and_specials():
    rng = np.random.default_rng(12)
    n = 500_000
    assert pow_mismatches(10 ** rng.uniform(-320, 308, n), rng.uniform(-3, 3, n)) == 0
    assert pow_mismatches(rng.uniform(-10, 10, n), rng.integers(-8, 9, n).astype(float)) == 0
    assert pow_mismatches(rng.uniform(-10, 10, n), rng.uniform(-3, 3, n)) == 0
    assert pow_mismatches(np.exp(rng.uniform(-2, 2, n)), rng.uniform(-600, 600, n)) == 0
    assert pow_mismatches(np.exp(rng.uniform(-1e-3, 1e-3, n)), rng.uniform(-1e6, 1e6, n)) == 0
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, 2.2e-308, 1.7e308, 0.5, 2.0, 3.0, -3.0,
                   1e-70, -1e-70, 1e70, 9e18, 1e19])
    X, Y = np.meshgrid(sp, sp)
    assert pow_mismatches(X.ravel(), Y.ravel()) == 0


def test_exp():
    rng = np.random.default_rng(13)
    n = 500_000
    assert exp_mismatches(rng.uniform(-5, 5, n)) == 0
    assert exp_mismatches(rng.uniform(-750, 720, n)) == 0
    sp = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-20, -1e-20, 709.78, 709.79, -745.1, -745.2, 1e-300])
    assert exp_mismatches(sp) == 0
    assert exp_mismatches(np.concatenate([10 ** rng.uniform(-30, 3, 100000), -10 ** rng.uniform(-30, 3, 100000)])) == 0


# ---- f32: ode_powf.h (groundwork for the f32 instantiation of the adaptive steppers)
def powf_mismatches(x, y):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(np.broadcast_to(y, x.shape), dtype=np.float32)
    f = _lib().ode_powcheck_powf
    f.restype = C.c_int64
    return f(C.c_int64(x.size), C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), None, None)


def test_powf_tables_match_running_libm():
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "extract_libm_powf_tables.py"), "--check"],
                   check=True, stdout=subprocess.DEVNULL)


def test_powf_matches_libm_bitwise():
    """controller exponents (f32 alpha / -beta, hinit's 1/5 and 1/8), the whole float range, integer
    exponents of negative bases, the overflow / underflow / may-underflow edges, every special case"""
    rng = np.random.default_rng(14)
    n = 1_000_000
    for a in (0.2 - 0.04 * 0.75, 0.125, -0.04, 0.2, 0.125 - 0.1 * 0.2):
        assert powf_mismatches(10 ** rng.uniform(-12, 3, n), a) == 0
    wide = np.ldexp(rng.uniform(1, 2, n), rng.integers(-149, 128, n)).astype(np.float32) * rng.choice([-1, 1], n)
    assert powf_mismatches(wide, rng.uniform(-40, 40, n)) == 0
    assert powf_mismatches(rng.uniform(-10, 10, n), rng.integers(-30, 30, n).astype(np.float32)) == 0
    assert powf_mismatches(np.exp(rng.uniform(-2, 2, n)), rng.uniform(-600, 600, n)) == 0
    assert powf_mismatches(np.full(n, 2.0), rng.uniform(-151, 129, n)) == 0
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, 1.1e-38, 3.4e38, 0.5, 2.0, 3.0, -3.0, 1e-30,
                   1e30, 16777216.0, 16777217.0, -2.5, 0.99999994, 1.0000001], dtype=np.float32)
    X, Y = np.meshgrid(sp, sp)
    assert powf_mismatches(X.ravel(), Y.ravel()) == 0
