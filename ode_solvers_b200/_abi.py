"""ctypes mirror of include/ode_b200.h (PODs, enums, library loader).

The product library is `ode_solvers_b200/libode_b200.so` (built in-tree by
`__graft_entry__.build()` / `make -C ode_solvers_b200/csrc`). There is NO fallback: if the
library is missing, loading raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# ODE_B200_LIB selects another build of the SAME library (kernel-tuning experiments only)
LIB_PATH = os.environ.get("ODE_B200_LIB") or os.path.join(HERE, "libode_b200.so")

# method ids / OutputType (src/dop_shared.rs:112-117) / status codes (src/dop_shared.rs:120-128)
RK4, DOPRI5, DOP853 = 0, 1, 2
OUT_DENSE, OUT_SPARSE, OUT_CONTINUOUS, OUT_CONTINUOUS8 = 0, 1, 2, 3
(OK, MAX_NUM_STEP_REACHED, STEP_SIZE_UNDERFLOW, STIFFNESS_DETECTED,
 OUTPUT_CAPACITY_EXCEEDED, UNSUPPORTED_DOMAIN) = range(6)
MAX_DIM, MAX_PARAMS = 32, 16

c_double_p = C.POINTER(C.c_double)
c_u32_p = C.POINTER(C.c_uint32)
c_i32_p = C.POINTER(C.c_int32)


class Config(C.Structure):
    """ode_b200_config: the from_param argument list (dopri5.rs:129-146, dop853.rs:133-150)."""
    _fields_ = [
        ("method", C.c_int32), ("out_type", C.c_int32),
        ("x0", C.c_double), ("x_end", C.c_double), ("dx", C.c_double),
        ("rtol", C.c_double), ("atol", C.c_double),
        ("safety_factor", C.c_double), ("alpha", C.c_double), ("beta", C.c_double),
        ("fac_min", C.c_double), ("fac_max", C.c_double), ("h_max", C.c_double),
        ("h0", C.c_double), ("uround", C.c_double), ("step_size", C.c_double),
        ("n_max", C.c_uint32), ("n_stiff", C.c_uint32),
    ]

    def copy(self):
        c = Config()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(Config))
        return c


class Outputs(C.Structure):
    """ode_b200_outputs."""
    _fields_ = [
        ("y_final", c_double_p), ("x_final", c_double_p), ("h_next", c_double_p),
        ("num_eval", c_u32_p), ("accepted", c_u32_p), ("rejected", c_u32_p), ("n_step", c_u32_p),
        ("status", c_i32_p),
        ("out_cap", C.c_int64), ("out_x", c_double_p), ("out_y", c_double_p), ("out_count", c_u32_p),
        ("com_cap", C.c_int64), ("com_lower", c_double_p), ("com_break", c_double_p),
        ("com_step", c_double_p), ("com_coef", c_double_p), ("com_count", c_u32_p),
    ]


c_float_p = C.POINTER(C.c_float)


class ConfigF32(C.Structure):
    """ode_b200_config_f32: ode_b200_config with T = f32, field for field."""
    _fields_ = [
        ("method", C.c_int32), ("out_type", C.c_int32),
        ("x0", C.c_float), ("x_end", C.c_float), ("dx", C.c_float),
        ("rtol", C.c_float), ("atol", C.c_float),
        ("safety_factor", C.c_float), ("alpha", C.c_float), ("beta", C.c_float),
        ("fac_min", C.c_float), ("fac_max", C.c_float), ("h_max", C.c_float),
        ("h0", C.c_float), ("uround", C.c_float), ("step_size", C.c_float),
        ("n_max", C.c_uint32), ("n_stiff", C.c_uint32),
    ]


class OutputsF32(C.Structure):
    """ode_b200_outputs_f32: ode_b200_outputs with T = f32, field for field."""
    _fields_ = [
        ("y_final", c_float_p), ("x_final", c_float_p), ("h_next", c_float_p),
        ("num_eval", c_u32_p), ("accepted", c_u32_p), ("rejected", c_u32_p), ("n_step", c_u32_p),
        ("status", c_i32_p),
        ("out_cap", C.c_int64), ("out_x", c_float_p), ("out_y", c_float_p), ("out_count", c_u32_p),
        ("com_cap", C.c_int64), ("com_lower", c_float_p), ("com_break", c_float_p),
        ("com_step", c_float_p), ("com_coef", c_float_p), ("com_count", c_u32_p),
    ]


class HostOutputsF32:
    """numpy-backed ode_b200_outputs_f32."""

    def __init__(self, n, dim, out_cap=0, com_cap=0, m=5):
        self.n, self.dim, self.out_cap, self.com_cap, self.m = n, dim, int(out_cap), int(com_cap), m
        self.y_final = np.zeros((dim, n), dtype=np.float32)
        self.x_final = np.zeros(n, dtype=np.float32)
        self.h_next = np.zeros(n, dtype=np.float32)
        self.num_eval = np.zeros(n, dtype=np.uint32)
        self.accepted = np.zeros(n, dtype=np.uint32)
        self.rejected = np.zeros(n, dtype=np.uint32)
        self.n_step = np.zeros(n, dtype=np.uint32)
        self.status = np.full(n, -1, dtype=np.int32)
        self.out_count = np.zeros(n, dtype=np.uint32)
        self.out_x = np.zeros((n, out_cap), dtype=np.float32) if out_cap else None
        self.out_y = np.zeros((n, out_cap, dim), dtype=np.float32) if out_cap else None
        self.com_lower = np.zeros(n, dtype=np.float32)
        self.com_count = np.zeros(n, dtype=np.uint32)
        self.com_break = np.zeros((n, com_cap), dtype=np.float32) if com_cap else None
        self.com_step = np.zeros((n, com_cap), dtype=np.float32) if com_cap else None
        self.com_coef = np.zeros((n, com_cap, m, dim), dtype=np.float32) if com_cap else None

    def struct(self):
        o = OutputsF32()
        for k in ("y_final", "x_final", "h_next", "out_x", "out_y", "com_lower", "com_break", "com_step", "com_coef"):
            setattr(o, k, _ptr(getattr(self, k), c_float_p))
        for k in ("num_eval", "accepted", "rejected", "n_step", "out_count", "com_count"):
            setattr(o, k, _ptr(getattr(self, k), c_u32_p))
        o.status = _ptr(self.status, c_i32_p)
        o.out_cap = self.out_cap
        o.com_cap = self.com_cap
        return o


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else typ()


class HostOutputs:
    """numpy-backed ode_b200_outputs (host memory)."""

    def __init__(self, n, dim, out_cap=0, com_cap=0, m=5):
        self.n, self.dim, self.out_cap, self.com_cap, self.m = n, dim, int(out_cap), int(com_cap), m
        self.y_final = np.zeros((dim, n), dtype=np.float64)
        self.x_final = np.zeros(n, dtype=np.float64)
        self.h_next = np.zeros(n, dtype=np.float64)
        self.num_eval = np.zeros(n, dtype=np.uint32)
        self.accepted = np.zeros(n, dtype=np.uint32)
        self.rejected = np.zeros(n, dtype=np.uint32)
        self.n_step = np.zeros(n, dtype=np.uint32)
        self.status = np.full(n, -1, dtype=np.int32)
        self.out_x = self.out_y = self.out_count = None
        if out_cap > 0:
            self.out_x = np.zeros((n, out_cap), dtype=np.float64)
            self.out_y = np.zeros((n, out_cap, dim), dtype=np.float64)
        self.out_count = np.zeros(n, dtype=np.uint32)
        self.com_lower = self.com_break = self.com_step = self.com_coef = None
        if com_cap > 0:
            self.com_break = np.zeros((n, com_cap), dtype=np.float64)
            self.com_step = np.zeros((n, com_cap), dtype=np.float64)
            self.com_coef = np.zeros((n, com_cap, m, dim), dtype=np.float64)
        self.com_lower = np.zeros(n, dtype=np.float64)
        self.com_count = np.zeros(n, dtype=np.uint32)

    def struct(self):
        o = Outputs()
        o.y_final = _ptr(self.y_final, c_double_p)
        o.x_final = _ptr(self.x_final, c_double_p)
        o.h_next = _ptr(self.h_next, c_double_p)
        o.num_eval = _ptr(self.num_eval, c_u32_p)
        o.accepted = _ptr(self.accepted, c_u32_p)
        o.rejected = _ptr(self.rejected, c_u32_p)
        o.n_step = _ptr(self.n_step, c_u32_p)
        o.status = _ptr(self.status, c_i32_p)
        o.out_cap = self.out_cap
        o.out_x = _ptr(self.out_x, c_double_p)
        o.out_y = _ptr(self.out_y, c_double_p)
        o.out_count = _ptr(self.out_count, c_u32_p)
        o.com_cap = self.com_cap
        o.com_lower = _ptr(self.com_lower, c_double_p)
        o.com_break = _ptr(self.com_break, c_double_p)
        o.com_step = _ptr(self.com_step, c_double_p)
        o.com_coef = _ptr(self.com_coef, c_double_p)
        o.com_count = _ptr(self.com_count, c_u32_p)
        return o


# every symbol include/ode_b200.h declares (tests/test_abi.py checks the .so exports all)
ABI_SYMBOLS = [
    "ode_b200_abi_version", "ode_b200_device_count", "ode_b200_last_error",
    "ode_b200_measure_fp64_peak",
    "ode_b200_system_builtin", "ode_b200_system_dim", "ode_b200_system_n_params",
    "ode_b200_system_name", "ode_b200_system_from_source", "ode_b200_system_compile",
    "ode_b200_config_new", "ode_b200_config_from_param", "ode_b200_config_rk4",
    "ode_b200_create", "ode_b200_destroy", "ode_b200_last_kernel_ms", "ode_b200_launch_count",
    "ode_b200_set_zero_copy", "ode_b200_set_lane_split", "ode_b200_set_pipeline",
    "ode_b200_stream",
    "ode_b200_integrate_batch", "ode_b200_integrate_batch_device", "ode_b200_integrate_batch_multi",
    "ode_b200_rk4_f32_integrate_batch", "ode_b200_integrate_batch_f32",
    "ode_b200_config_new_f32", "ode_b200_config_from_param_f32", "ode_b200_config_rk4_f32",
    "ode_b200_set_work_queue", "ode_b200_com_evaluate", "ode_b200_com_evaluate_device",
    "ode_b200_tableau", "ode_b200_debug_pow", "ode_b200_debug_exp", "ode_b200_debug_div", "ode_b200_debug_powf",
]

_lib = None


def declare_config_fns(lib, prefix):
    """Attach argtypes for the three constructor helpers (shared by product and oracle)."""
    f = getattr(lib, prefix + "config_new")
    f.argtypes = [C.c_int] + [C.c_double] * 5 + [C.POINTER(Config)]
    f.restype = C.c_int
    f = getattr(lib, prefix + "config_from_param")
    f.argtypes = ([C.c_int] + [C.c_double] * 5 + [C.c_double] * 6 + [C.c_uint32, C.c_uint32, C.c_int]
                  + [C.POINTER(Config)])
    f.restype = C.c_int
    f = getattr(lib, prefix + "config_rk4")
    f.argtypes = [C.c_double] * 3 + [C.POINTER(Config)]
    f.restype = C.c_int


def declare_config_fns_f32(lib, prefix):
    """the f32 constructors (shared by product and oracle)"""
    f = getattr(lib, prefix + "config_new_f32")
    f.argtypes = [C.c_int] + [C.c_float] * 5 + [C.POINTER(ConfigF32)]
    f.restype = C.c_int
    f = getattr(lib, prefix + "config_from_param_f32")
    f.argtypes = [C.c_int] + [C.c_float] * 11 + [C.c_uint32, C.c_uint32, C.c_int] + [C.POINTER(ConfigF32)]
    f.restype = C.c_int
    f = getattr(lib, prefix + "config_rk4_f32")
    f.argtypes = [C.c_float] * 3 + [C.POINTER(ConfigF32)]
    f.restype = C.c_int


def load():
    """Load libode_b200.so (raises if it has not been built -- there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "ode_solvers_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C ode_solvers_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.ode_b200_abi_version.restype = C.c_int
    lib.ode_b200_device_count.restype = C.c_int
    lib.ode_b200_last_error.restype = C.c_char_p
    lib.ode_b200_measure_fp64_peak.argtypes = [C.c_int, c_double_p, c_double_p]
    lib.ode_b200_measure_fp64_peak.restype = C.c_int
    lib.ode_b200_system_builtin.argtypes = [C.c_char_p]
    lib.ode_b200_system_builtin.restype = C.c_int
    lib.ode_b200_system_dim.argtypes = [C.c_int]
    lib.ode_b200_system_n_params.argtypes = [C.c_int]
    lib.ode_b200_system_name.argtypes = [C.c_int]
    lib.ode_b200_system_name.restype = C.c_char_p
    lib.ode_b200_system_from_source.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_char_p]
    lib.ode_b200_system_from_source.restype = C.c_int
    lib.ode_b200_system_compile.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.ode_b200_system_compile.restype = C.c_int
    declare_config_fns(lib, "ode_b200_")
    lib.ode_b200_create.argtypes = [C.c_int]
    lib.ode_b200_create.restype = C.c_void_p
    lib.ode_b200_destroy.argtypes = [C.c_void_p]
    lib.ode_b200_destroy.restype = None
    lib.ode_b200_last_kernel_ms.argtypes = [C.c_void_p]
    lib.ode_b200_last_kernel_ms.restype = C.c_double
    lib.ode_b200_launch_count.argtypes = [C.c_void_p]
    lib.ode_b200_launch_count.restype = C.c_uint64
    lib.ode_b200_stream.argtypes = [C.c_void_p]
    lib.ode_b200_stream.restype = C.c_void_p
    lib.ode_b200_integrate_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(Config), C.c_int64,
                                             C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Outputs)]
    lib.ode_b200_integrate_batch.restype = C.c_int
    lib.ode_b200_integrate_batch_device.argtypes = [C.c_void_p, C.c_int, C.POINTER(Config), C.c_int64,
                                                    C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Outputs),
                                                    C.c_void_p]
    lib.ode_b200_integrate_batch_device.restype = C.c_int
    lib.ode_b200_integrate_batch_multi.argtypes = [C.c_int, C.POINTER(Config), C.c_int64, C.c_void_p,
                                                   C.c_void_p, C.c_int, C.c_int, C.POINTER(Outputs)]
    lib.ode_b200_integrate_batch_multi.restype = C.c_int
    lib.ode_b200_rk4_f32_integrate_batch.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int64,
                                                     C.c_void_p, C.c_void_p, C.c_int, C.POINTER(OutputsF32)]
    lib.ode_b200_rk4_f32_integrate_batch.restype = C.c_int
    declare_config_fns_f32(lib, "ode_b200_")
    lib.ode_b200_integrate_batch_f32.argtypes = [C.c_void_p, C.c_int, C.POINTER(ConfigF32), C.c_int64, C.c_void_p,
                                                 C.c_void_p, C.c_int, C.POINTER(OutputsF32)]
    lib.ode_b200_integrate_batch_f32.restype = C.c_int
    lib.ode_b200_set_work_queue.argtypes = [C.c_void_p, C.c_int]
    lib.ode_b200_set_work_queue.restype = C.c_int
    lib.ode_b200_set_zero_copy.argtypes = [C.c_void_p, C.c_int]
    lib.ode_b200_set_zero_copy.restype = C.c_int
    lib.ode_b200_set_lane_split.argtypes = [C.c_void_p, C.c_int]
    lib.ode_b200_set_lane_split.restype = C.c_int
    lib.ode_b200_set_pipeline.argtypes = [C.c_void_p, C.c_int]
    lib.ode_b200_set_pipeline.restype = C.c_int
    lib.ode_b200_com_evaluate.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ode_b200_com_evaluate.restype = C.c_int
    lib.ode_b200_com_evaluate_device.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ode_b200_com_evaluate_device.restype = C.c_int
    lib.ode_b200_tableau.argtypes = [C.c_char_p, C.c_int, C.c_int]
    lib.ode_b200_tableau.restype = C.c_double
    lib.ode_b200_debug_pow.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ode_b200_debug_pow.restype = C.c_int
    lib.ode_b200_debug_powf.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ode_b200_debug_powf.restype = C.c_int
    lib.ode_b200_debug_exp.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    lib.ode_b200_debug_exp.restype = C.c_int
    lib.ode_b200_debug_div.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.ode_b200_debug_div.restype = C.c_int
    _lib = lib
    return lib


def last_error():
    return load().ode_b200_last_error().decode("utf-8", "replace")
