"""Device pow()/exp() clones must be BIT-identical to the host libm the reference's
f64::powf / f64::exp resolve to (controller.rs:53-54, dopri5.rs:236, dop853.rs:244)."""
import os

import numpy as np
import pytest

import ode_solvers_b200 as ob
import oracle
from ode_solvers_b200 import _abi
from parity_util import same_f64

pytestmark = pytest.mark.gpu


def _dev_pow(x, y):
    lib = _abi.load()
    ctx = ob.default_context(0)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(np.broadcast_to(y, x.shape), dtype=np.float64)
    out = np.zeros_like(x)
    assert lib.ode_b200_debug_pow(ctx.handle, x.size, x.ctypes.data, y.ctypes.data, out.ctypes.data) == 0
    return out


def _dev_exp(x):
    lib = _abi.load()
    ctx = ob.default_context(0)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    assert lib.ode_b200_debug_exp(ctx.handle, x.size, x.ctypes.data, out.ctypes.data) == 0
    return out


def _libm_pow(x, y):
    y = np.broadcast_to(y, np.shape(x))
    f = oracle.lib().ode_oracle_pow
    return np.array([f(float(a), float(b)) for a, b in zip(np.ravel(x), np.ravel(y))]).reshape(np.shape(x))


def _libm_exp(x):
    f = oracle.lib().ode_oracle_exp
    return np.array([f(float(a)) for a in np.ravel(x)]).reshape(np.shape(x))


def test_pow_controller_domain():
    rng = np.random.default_rng(3)
    n = 200_000
    for a in (0.2 - 0.04 * 0.75, 0.125, -0.04, 0.2):
        x = 10 ** rng.uniform(-8, 2, n)
        assert same_f64(_dev_pow(x, a), _libm_pow(x, a)).all()


def test_pow_wide_and_special():
    rng = np.random.default_rng(4)
    n = 200_000
    x = 10 ** rng.uniform(-320, 308, n)
    y = rng.uniform(-3, 3, n)
    assert same_f64(_dev_pow(x, y), _libm_pow(x, y)).all()
    x = rng.uniform(-10, 10, n)
    y = rng.integers(-8, 9, n).astype(float)
    assert same_f64(_dev_pow(x, y), _libm_pow(x, y)).all()
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, 2.2e-308, 1.7e308, 0.5, 2.0, 3.0, -3.0,
                   1e-70, 1e70, 9e18, 1e19])
    X, Y = np.meshgrid(sp, sp)
    assert same_f64(_dev_pow(X.ravel(), Y.ravel()), _libm_pow(X.ravel(), Y.ravel())).all()
    x = np.exp(rng.uniform(-2, 2, n))
    y = rng.uniform(-600, 600, n)
    assert same_f64(_dev_pow(x, y), _libm_pow(x, y)).all()


def test_exp():
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-5, 5, 200_000), rng.uniform(-750, 720, 200_000),
                        [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-20, -1e-20, 709.78, 709.79, -745.1, -745.2]])
    assert same_f64(_dev_exp(x), _libm_exp(x)).all()


def _dev_div(a, b, shared=0):
    lib = _abi.load()
    ctx = ob.default_context(0)
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(np.broadcast_to(b, a.shape), dtype=np.float64)
    out = np.zeros_like(a)
    assert lib.ode_b200_debug_div(ctx.handle, a.size, a.ctypes.data, b.ctypes.data, out.ctypes.data, shared) == 0
    return out


def test_div_matches_ieee():
    """the kernels' interleavable division must be the correctly rounded IEEE quotient (what
    Rust's `/` and the oracle's `/` produce) for every operand class"""
    rng = np.random.default_rng(6)
    n = 20_000_000
    with np.errstate(all="ignore"):
        for rep in range(5):
            # random mantissas over a wide exponent range (inside and outside the fast window)
            a = np.ldexp(rng.uniform(1, 2, n), rng.integers(-1070, 1023, n)) * rng.choice([-1.0, 1.0], n)
            b = np.ldexp(rng.uniform(1, 2, n), rng.integers(-1070, 1023, n)) * rng.choice([-1.0, 1.0], n)
            if rep >= 2:  # typical solver magnitudes
                a = rng.standard_normal(n) * 10 ** rng.uniform(-12, 6, n)
                b = rng.uniform(1e-12, 1e3, n)
            if rep == 4:  # hard cases: mantissas at the ends, near-ties
                a = np.ldexp(1 + rng.integers(0, 8, n) * 2.0 ** -52, rng.integers(-300, 300, n))
                b = np.ldexp(2 - rng.integers(1, 8, n) * 2.0 ** -52, rng.integers(-300, 300, n))
            assert same_f64(_dev_div(a, b), a / b).all()
            assert same_f64(_dev_div(a, b, shared=1), a / b).all()
        sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, 2.2e-308, 1.7e308, 3.0, 1e-200, 1e200,
                       -7.5, 0.9])
        A, B = np.meshgrid(sp, sp)
        assert same_f64(_dev_div(A.ravel(), B.ravel()), A.ravel() / B.ravel()).all()


FLAGGED = 0x7ff8f1a900000000


def _spec(a, b, mode):
    out = _dev_div(a, b, shared=mode)
    flagged = out.view(np.uint64) == FLAGGED
    return out, flagged


def test_speculative_div_and_sqrt_are_ieee_exact_or_flagged():
    """MathFast (ode_device.cuh): every value the branch-free fast paths return must be the correctly
    rounded IEEE result; operands they cannot handle must be flagged, never answered wrongly; and
    ordinary solver magnitudes (and zero numerators) must NOT be flagged, or the step loop would
    live on its out-of-line exact path"""
    rng = np.random.default_rng(16)
    n = 10_000_000
    with np.errstate(all="ignore"):
        for rep in range(5):
            a = np.ldexp(rng.uniform(1, 2, n), rng.integers(-1070, 1023, n)) * rng.choice([-1.0, 1.0], n)
            b = np.ldexp(rng.uniform(1, 2, n), rng.integers(-1070, 1023, n)) * rng.choice([-1.0, 1.0], n)
            ordinary = rep >= 2
            if ordinary:
                a = rng.standard_normal(n) * 10 ** rng.uniform(-30, 30, n)
                b = rng.uniform(1e-30, 1e30, n) * rng.choice([-1.0, 1.0], n)
                a[rng.integers(0, n, n // 10)] = 0.0
                a[rng.integers(0, n, n // 10)] = -0.0
            if rep == 4:  # near-ties, mantissas at the ends
                a = np.ldexp(1 + rng.integers(0, 8, n) * 2.0 ** -52, rng.integers(-300, 300, n))
                b = np.ldexp(2 - rng.integers(1, 8, n) * 2.0 ** -52, rng.integers(-300, 300, n))
            for mode in (2, 3, 6):
                out, fl = _spec(a, b, mode)
                assert same_f64(out[~fl], (a / b)[~fl]).all(), "mode %d rep %d" % (mode, rep)
                if ordinary:
                    nz = a != 0.0
                    assert not fl[nz].any()
                    assert fl[~nz].all() if mode == 3 else not fl[~nz].any()
                else:
                    assert 0.2 < fl.mean() < 0.9  # the wide exponent sweep leaves the window often
            x = np.abs(a)
            for mode in (4, 5):
                out, fl = _spec(x, x, mode)
                assert same_f64(out[~fl], np.sqrt(x)[~fl]).all()
                if ordinary:
                    nz = x != 0.0
                    assert not fl[nz].any()
                    assert fl[~nz].all() if mode == 5 else not fl[~nz].any()
        sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, 2.2e-308, 1.7e308, 3.0, 1e-200, 1e200,
                       -7.5, 0.9, 1e-300, 4e-292, 1e-291, 2.0 ** -970, 2.0 ** -969, 2.0 ** 1016, 2.0 ** 1017])
        A, B = (v.ravel() for v in np.meshgrid(sp, sp))
        for mode in (2, 3, 6):
            out, fl = _spec(A, B, mode)
            assert same_f64(out[~fl], (A / B)[~fl]).all()
            assert fl[~np.isfinite(A / B)].all() and fl[~np.isfinite(B)].all()
        for mode in (4, 5):
            out, fl = _spec(sp, sp, mode)
            assert same_f64(out[~fl], np.sqrt(sp)[~fl]).all()
            assert fl[~(sp >= 0)].all() and fl[np.isinf(sp)].all()


def test_powf_clone_on_the_device_matches_libm():
    """ode_powf.h (groundwork for the f32 adaptive steppers), device build, against the host's libm powf
    (the host build of the same header is compared with libm in tests/test_pow_clone.py)"""
    import ctypes as C
    libm = C.CDLL("libm.so.6")
    lib = _abi.load()
    lib.ode_b200_debug_powf.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ode_b200_debug_powf.restype = C.c_int
    chk = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "libode_powcheck.so"))
    chk.ode_powcheck_powf.restype = C.c_int64
    ctx = ob.default_context(0)
    rng = np.random.default_rng(21)
    n = 2_000_000
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, 1.1e-38, 3.4e38, 0.5, 2.0, 3.0, -3.0, 1e-30,
                   1e30, 16777216.0, -2.5], dtype=np.float32)
    SX, SY = np.meshgrid(sp, sp)
    cases = [(10 ** rng.uniform(-12, 3, n), rng.choice([0.17, 0.125, -0.04, 0.2, 0.105], n)),
             (np.ldexp(rng.uniform(1, 2, n), rng.integers(-149, 128, n)) * rng.choice([-1, 1], n), rng.uniform(-40, 40, n)),
             (rng.uniform(-10, 10, n), rng.integers(-30, 30, n)), (np.full(n, 2.0), rng.uniform(-151, 129, n)),
             (SX.ravel(), SY.ravel())]
    for x, y in cases:
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        dev = np.zeros_like(x)
        ref = np.zeros_like(x)
        assert lib.ode_b200_debug_powf(ctx.handle, x.size, x.ctypes.data, y.ctypes.data, dev.ctypes.data) == 0
        chk.ode_powcheck_powf(C.c_int64(x.size), C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), None,
                              C.c_void_p(ref.ctypes.data))  # ref = libm powf
        same = (dev.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(dev) & np.isnan(ref))
        assert same.all(), (x[~same][:3], y[~same][:3], dev[~same][:3], ref[~same][:3])
    del libm
