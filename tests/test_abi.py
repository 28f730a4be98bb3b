"""The C-ABI library loads on a machine WITHOUT a GPU, exports every symbol include/ode_b200.h
declares, reproduces the constructors' defaults, and fails loudly (no CPU fallback) when asked
to integrate without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle
import ode_solvers_b200 as ob
from ode_solvers_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "ode_b200.h")).read()
    declared = set(re.findall(r"\b(ode_b200_[a-z0-9_]+)\s*\(", hdr))
    lib = _abi.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(_abi.ABI_SYMBOLS), declared ^ set(_abi.ABI_SYMBOLS)
    assert lib.ode_b200_abi_version() == 2  # 2: the f32 mirror structs + entries, ode_b200_set_lane_split


def test_builtin_systems_match_the_oracle_registry():
    lib = _abi.load()
    for name in oracle.systems():
        sid = lib.ode_b200_system_builtin(name.encode())
        assert sid >= 0, name
        assert (lib.ode_b200_system_dim(sid), lib.ode_b200_system_n_params(sid)) == oracle.system_info(name)
        assert lib.ode_b200_system_name(sid).decode() == name
    assert lib.ode_b200_system_builtin(b"no_such_system") < 0
    assert "no_such_system" in _abi.last_error()


def _fields(c):
    return [getattr(c, f) for f, _ in _abi.Config._fields_]


def test_constructors_match_the_reference_defaults_bit_for_bit():
    for m in (ob.DOPRI5, ob.DOP853):
        a, b = ob.config_new(m, 0.0, 20.0, 0.1, 1e-6, 1e-7), oracle.config_new(m, 0.0, 20.0, 0.1, 1e-6, 1e-7)
        assert _fields(a) == _fields(b)
        args = (m, 1.0, -3.0, 0.25, 1e-8, 1e-9, 0.85, 0.08, 0.25, 8.0, 2.0, 1e-3, 5000, 77, ob.OUT_SPARSE)
        assert _fields(ob.config_from_param(*args)) == _fields(oracle.config_from_param(*args))
    c = ob.config_new(ob.DOPRI5, 0.0, 1.0, 0.1, 1e-6, 1e-6)
    # dopri5.rs:14-27,76-105
    assert (c.alpha, c.beta, c.fac_max, c.fac_min, c.safety_factor) == (0.2 - 0.04 * 0.75, 0.04, 10.0, 0.2, 0.9)
    assert (c.n_max, c.n_stiff, c.out_type, c.h0, c.uround) == (100000, 1000, ob.OUT_DENSE, 0.0, 2.0 ** -52)
    c = ob.config_new(ob.DOP853, 0.0, 1.0, 0.1, 1e-6, 1e-6)
    # dop853.rs:14-27,76-109
    assert (c.alpha, c.beta, c.fac_max, c.fac_min, c.safety_factor) == (1.0 / 8.0, 0.0, 6.0, 0.333, 0.9)
    assert _fields(ob.config_rk4(0.0, 2.0, 0.01)) == _fields(oracle.config_rk4(0.0, 2.0, 0.01))


def test_no_cpu_fallback_without_a_device():
    if ob.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(ob.Ode200Error) as e:
        ob.Context(0)
    assert "no CPU fallback" in str(e.value)
    lib = _abi.load()
    cfg = ob.config_new(ob.DOPRI5, 0.0, 1.0, 0.1, 1e-6, 1e-6)
    y0 = np.ones((3, 4))
    p = np.array([10.0, 8 / 3, 28.0])
    out = _abi.HostOutputs(4, 3)
    s = out.struct()
    rc = lib.ode_b200_integrate_batch_multi(0, C.byref(cfg), 4, y0.ctypes.data, p.ctypes.data, 0, 1, C.byref(s))
    assert rc == -2 and "no CPU fallback" in _abi.last_error()
    assert lib.ode_b200_integrate_batch(None, 0, C.byref(cfg), 4, y0.ctypes.data, p.ctypes.data, 0, C.byref(s)) < 0
    assert (out.status == -1).all()  # nothing was computed


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ode_solvers_b200")
    for dp, _dn, fn in os.walk(pkg):
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) and "build" not in dp:
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "libode_oracle" not in txt and "import oracle" not in txt, os.path.join(dp, f)
                assert "ode_oracle_" not in txt or f == "ode_systems.cuh", os.path.join(dp, f)


def test_ctypes_prototypes_match_the_header():
    """every argtypes / restype the Python wrapper (and tests/oracle.py through it) declares agrees with the C
    prototype in include/ode_b200.h: argument count, integer width, float vs double, pointer vs scalar. A wrong width
    would silently truncate an argument on both sides of a parity test."""
    import ctypes as C
    from test_rust_ffi import c_protos
    lib = _abi.load()
    scalar = {"c_int": (C.c_int,), "i32": (C.c_int32, C.c_int), "u32": (C.c_uint32, C.c_uint), "i64": (C.c_int64, C.c_longlong, C.c_long),
              "u64": (C.c_uint64, C.c_ulonglong, C.c_ulong), "c_double": (C.c_double,), "c_float": (C.c_float,), "u8": (C.c_uint8,)}

    def is_pointer(t):
        return t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer))

    checked = 0
    for name, (ret, args) in c_protos().items():
        fn = getattr(lib, name)
        if fn.argtypes is None:
            assert not args, "%s takes %d arguments but declares no argtypes" % (name, len(args))
        else:
            assert len(fn.argtypes) == len(args), (name, len(fn.argtypes), args)
            for i, (ct, want) in enumerate(zip(fn.argtypes, args)):
                if want.startswith("*"):
                    assert is_pointer(ct), (name, i, ct, want)
                else:
                    assert ct in scalar[want], (name, i, ct, want)
        if ret is None:
            assert fn.restype is None, name
        elif ret.startswith("*"):
            assert is_pointer(fn.restype), (name, fn.restype, ret)
        else:
            assert fn.restype in scalar[ret], (name, fn.restype, ret)
        checked += 1
    assert checked == len(_abi.ABI_SYMBOLS)
