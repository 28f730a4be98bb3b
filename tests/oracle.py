"""Test-side ctypes wrapper of the CPU oracle (oracle/libode_oracle.so).

TEST INFRASTRUCTURE ONLY: nothing under ode_solvers_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from ode_solvers_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
# ODE_ORACLE_LIB selects another build of the SAME oracle (the ASan/UBSan build of tests/test_sanitizers.py)
ORACLE_LIB = os.environ.get("ODE_ORACLE_LIB") or os.path.join(ORACLE_DIR, "libode_oracle.so")

_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
        if "ODE_ORACLE_LIB" not in os.environ and (
                not os.path.exists(ORACLE_LIB)
                or os.path.getmtime(ORACLE_LIB) < max(os.path.getmtime(s) for s in srcs)):
            build()
        l = C.CDLL(ORACLE_LIB)
        _abi.declare_config_fns(l, "ode_oracle_")
        _abi.declare_config_fns_f32(l, "ode_oracle_")
        l.ode_oracle_powf.argtypes = [C.c_float, C.c_float]
        l.ode_oracle_powf.restype = C.c_float
        l.ode_oracle_integrate_batch.argtypes = [C.c_char_p, C.POINTER(_abi.Config), C.c_int64, C.c_void_p,
                                                 C.c_void_p, C.c_int, C.POINTER(_abi.Outputs), C.c_int]
        l.ode_oracle_integrate_batch.restype = C.c_int
        l.ode_oracle_com_evaluate.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_double, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_double, C.c_void_p]
        l.ode_oracle_com_evaluate.restype = C.c_int
        l.ode_oracle_tableau.argtypes = [C.c_char_p, C.c_int, C.c_int]
        l.ode_oracle_tableau.restype = C.c_double
        l.ode_oracle_max_threads.restype = C.c_int
        l.ode_oracle_pow.argtypes = [C.c_double, C.c_double]
        l.ode_oracle_pow.restype = C.c_double
        l.ode_oracle_exp.argtypes = [C.c_double]
        l.ode_oracle_exp.restype = C.c_double

        class Sys(C.Structure):
            _fields_ = [("name", C.c_char_p), ("dim", C.c_int), ("n_params", C.c_int),
                        ("rhs", C.c_void_p), ("solout", C.c_void_p)]
        l.ode_oracle_system_by_name.argtypes = [C.c_char_p]
        l.ode_oracle_system_by_name.restype = C.POINTER(Sys)
        l.ode_oracle_system_count.restype = C.c_int
        l.ode_oracle_system_at.argtypes = [C.c_int]
        l.ode_oracle_system_at.restype = C.POINTER(Sys)
        _lib = l
    return _lib


def system_info(name):
    s = lib().ode_oracle_system_by_name(name.encode())
    if not s:
        raise KeyError(name)
    return s.contents.dim, s.contents.n_params


def systems():
    l = lib()
    return [l.ode_oracle_system_at(i).contents.name.decode() for i in range(l.ode_oracle_system_count())]


def config_new(method, x, x_end, dx, rtol, atol, out_type=None):
    c = _abi.Config()
    assert lib().ode_oracle_config_new(method, x, x_end, dx, rtol, atol, C.byref(c)) == 0
    if out_type is not None:
        c.out_type = out_type
    return c


def config_from_param(method, x, x_end, dx, rtol, atol, safety, beta, fac_min, fac_max, h_max, h, n_max,
                      n_stiff, out_type):
    c = _abi.Config()
    assert lib().ode_oracle_config_from_param(method, x, x_end, dx, rtol, atol, safety, beta, fac_min, fac_max,
                                              h_max, h, n_max, n_stiff, out_type, C.byref(c)) == 0
    return c


def config_rk4(x, x_end, step):
    c = _abi.Config()
    assert lib().ode_oracle_config_rk4(x, x_end, step, C.byref(c)) == 0
    return c


def integrate_batch(system, cfg, y0, params=None, out_cap=0, com_cap=0, n_threads=0):
    """y0: [dim][n] (or [dim] for one trajectory); params: [n_params] or [n_params][n]."""
    dim, n_params = system_info(system)
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float64).reshape(dim, -1))
    n = y0.shape[1]
    per_traj = 0
    pp = None
    if n_params:
        params = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
        per_traj = int(params.ndim == 2)
        assert params.shape[0] == n_params and (not per_traj or params.shape[1] == n)
        pp = params.ctypes.data
    m = 5 if cfg.method == _abi.DOPRI5 else 8
    out = _abi.HostOutputs(n, dim, out_cap, com_cap, m)
    s = out.struct()
    rc = lib().ode_oracle_integrate_batch(system.encode(), C.byref(cfg), n, y0.ctypes.data, pp, per_traj,
                                          C.byref(s), n_threads)
    assert rc == 0, rc
    return out


def config_new_f32(method, x, x_end, dx, rtol, atol, out_type=None):
    c = _abi.ConfigF32()
    assert lib().ode_oracle_config_new_f32(method, x, x_end, dx, rtol, atol, C.byref(c)) == 0
    if out_type is not None:
        c.out_type = out_type
    return c


def config_from_param_f32(method, x, x_end, dx, rtol, atol, safety, beta, fac_min, fac_max, h_max, h, n_max,
                          n_stiff, out_type):
    c = _abi.ConfigF32()
    assert lib().ode_oracle_config_from_param_f32(method, x, x_end, dx, rtol, atol, safety, beta, fac_min, fac_max,
                                                  h_max, h, n_max, n_stiff, out_type, C.byref(c)) == 0
    return c


def config_rk4_f32(x, x_end, step):
    c = _abi.ConfigF32()
    assert lib().ode_oracle_config_rk4_f32(x, x_end, step, C.byref(c)) == 0
    return c


def integrate_batch_f32(system, cfg, y0, params=None, out_cap=0, com_cap=0, n_threads=0):
    """the three steppers with T = f32; y0: [dim][n] float32"""
    dim, n_params = system_info(system)
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float32).reshape(dim, -1))
    n = y0.shape[1]
    per_traj, pp = 0, None
    if n_params:
        params = np.ascontiguousarray(np.asarray(params, dtype=np.float32))
        per_traj = int(params.ndim == 2)
        pp = params.ctypes.data
    m = 5 if cfg.method == _abi.DOPRI5 else 8
    out = _abi.HostOutputsF32(n, dim, out_cap, com_cap, m)
    s = out.struct()
    f = lib().ode_oracle_integrate_batch_f32
    f.argtypes = [C.c_char_p, C.POINTER(_abi.ConfigF32), C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                  C.POINTER(_abi.OutputsF32), C.c_int]
    f.restype = C.c_int
    rc = f(system.encode(), C.byref(cfg), n, y0.ctypes.data, pp, per_traj, C.byref(s), n_threads)
    assert rc == 0, rc
    return out


def rk4_f32_integrate_batch(system, x, x_end, step, y0, params=None, out_cap=0):
    dim, n_params = system_info(system)
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float32).reshape(dim, -1))
    n = y0.shape[1]
    per_traj, pp = 0, None
    if n_params:
        params = np.ascontiguousarray(np.asarray(params, dtype=np.float32))
        per_traj = int(params.ndim == 2)
        pp = params.ctypes.data
    out = _abi.HostOutputsF32(n, dim, out_cap)
    s = out.struct()
    f = lib().ode_oracle_rk4_f32_integrate_batch
    f.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_float, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                  C.POINTER(_abi.OutputsF32), C.c_int]
    f.restype = C.c_int
    assert f(system.encode(), x, x_end, step, n, y0.ctypes.data, pp, per_traj, C.byref(s), 0) == 0
    return out


def com_evaluate(out, i, x):
    """ContinuousOutputModel::evaluate for trajectory i of a batch result; None or ndarray."""
    cnt = int(min(out.com_count[i], out.com_cap))
    y = np.zeros(out.dim)
    ok = lib().ode_oracle_com_evaluate(out.dim, out.m, cnt, float(out.com_lower[i]),
                                       out.com_break[i].ctypes.data, out.com_step[i].ctypes.data,
                                       out.com_coef[i].ctypes.data, float(x), y.ctypes.data)
    return y if ok else None
