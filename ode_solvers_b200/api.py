"""Host-side mirror of the reference crate's surface over the C ABI (include/ode_b200.h).

The reference is a Rust crate and this image has no Rust toolchain, so the host layer the
parity tests drive is Python (ctypes); `rust/` holds the same surface as Rust sources and
`ode_solvers_b200/cpp/ode_b200.hpp` as C++. Names, argument order and error behaviour follow
the crate:

    System                     trait System                     src/dop_shared.rs:31-42
    Dopri5.new / from_param    Dopri5::new / from_param         src/dopri5.rs:76, :129
    Dop853.new / from_param    Dop853::new / from_param         src/dop853.rs:76, :133
    Rk4.new                    Rk4::new(f, x, y, x_end, step)   src/rk4.rs:42
    .integrate()               -> Result<Stats, IntegrationError>  (raises IntegrationError)
    .integrate_with_continuous_output_model(com)                src/dopri5.rs:246
    .x_out() / .y_out() / .results()
    Stats, OutputType, IntegrationError                         src/dop_shared.rs:112-154
    ContinuousOutputModel.evaluate / bounds                     src/continuous_output_model.rs:31-64
plus the batched entry point `integrate_batch` (many trajectories, one config).

All numerics run in the CUDA library; nothing here computes a step. There is no CPU fallback.
"""
import ctypes as C
import enum
import math

import numpy as np

from . import _abi


class OutputType(enum.IntEnum):
    """src/dop_shared.rs:112-117"""
    Dense = _abi.OUT_DENSE
    Sparse = _abi.OUT_SPARSE
    Continuous = _abi.OUT_CONTINUOUS


class Stats:
    """src/dop_shared.rs:131-154"""

    def __init__(self, num_eval=0, accepted_steps=0, rejected_steps=0):
        self.num_eval, self.accepted_steps, self.rejected_steps = int(num_eval), int(accepted_steps), int(rejected_steps)

    def __str__(self):
        return ("Number of function evaluations: %d\nNumber of accepted steps: %d\nNumber of rejected steps: %d"
                % (self.num_eval, self.accepted_steps, self.rejected_steps))

    def __repr__(self):
        return "Stats(num_eval=%d, accepted_steps=%d, rejected_steps=%d)" % (
            self.num_eval, self.accepted_steps, self.rejected_steps)

    def __eq__(self, o):
        return (self.num_eval, self.accepted_steps, self.rejected_steps) == (
            o.num_eval, o.accepted_steps, o.rejected_steps)


class IntegrationError(Exception):
    """src/dop_shared.rs:120-128"""


class MaxNumStepReached(IntegrationError):
    def __init__(self, x, n_step):
        super().__init__("Stopped at x = %s. Need more than %d steps." % (x, n_step))
        self.x, self.n_step = x, n_step


class StepSizeUnderflow(IntegrationError):
    def __init__(self, x):
        super().__init__("Stopped at x = %s. Step size underflow." % x)
        self.x = x


class StiffnessDetected(IntegrationError):
    def __init__(self, x):
        super().__init__("The problem seems to become stiff at x = %s." % x)
        self.x = x


class UnsupportedDomain(IntegrationError):
    """Inputs on which the reference panics or never returns."""


class Ode200Error(RuntimeError):
    """A failing ABI call (CUDA error, bad argument, missing device)."""


def _check(rc, what):
    if rc != 0:
        raise Ode200Error("%s failed (%d): %s" % (what, rc, _abi.last_error()))


# ------------------------------------------------------------------ contexts
class Context:
    """ode_b200_ctx: one per (host thread, GPU)."""

    def __init__(self, device=0):
        lib = _abi.load()
        self._lib = lib
        self.handle = lib.ode_b200_create(int(device))
        if not self.handle:
            raise Ode200Error("ode_b200_create(%d): %s" % (device, _abi.last_error()))
        self.device = device

    def close(self):
        if self.handle:
            self._lib.ode_b200_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_work_queue(self, enabled):
        _check(self._lib.ode_b200_set_work_queue(self.handle, int(bool(enabled))), "set_work_queue")

    def set_zero_copy(self, enabled):
        """host-buffer entry: store per-trajectory results straight into pinned host arrays (default)
        or go through device buffers and DMA copies (better when several GPUs share the host's PCIe)"""
        _check(self._lib.ode_b200_set_zero_copy(self.handle, int(bool(enabled))), "set_zero_copy")

    def set_pipeline(self, chunk_log2):
        """host-buffer entry: chunked upload / integrate / copy-back pipeline (default 16 = 65536 trajectories per
        chunk; 0 = off)"""
        _check(self._lib.ode_b200_set_pipeline(self.handle, int(chunk_log2)), "set_pipeline")

    def set_lane_split(self, enabled):
        """large states (threebody18, Dopri5 / Dop853): one trajectory per group of three lanes (default) or per thread"""
        _check(self._lib.ode_b200_set_lane_split(self.handle, int(bool(enabled))), "set_lane_split")

    @property
    def last_kernel_ms(self):
        return float(self._lib.ode_b200_last_kernel_ms(self.handle))

    @property
    def launch_count(self):
        return int(self._lib.ode_b200_launch_count(self.handle))


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def device_count():
    return int(_abi.load().ode_b200_device_count())


def measure_fp64_peak(device=0):
    """(FP64 thread-instructions/s, implied SM MHz) from the DFMA-chain probe."""
    a, b = C.c_double(), C.c_double()
    _check(_abi.load().ode_b200_measure_fp64_peak(device, C.byref(a), C.byref(b)), "measure_fp64_peak")
    return a.value, b.value


# ------------------------------------------------------------------ systems
class System:
    """`trait System` (src/dop_shared.rs:31-42): a device RHS (+ optional solout) and the
    parameter row that replaces the implementing struct's fields."""

    def __init__(self, sys_id, params=None):
        lib = _abi.load()
        self.sys_id = int(sys_id)
        self.dim = lib.ode_b200_system_dim(self.sys_id)
        self.n_params = lib.ode_b200_system_n_params(self.sys_id)
        if self.dim < 0:
            raise Ode200Error(_abi.last_error())
        self.name = lib.ode_b200_system_name(self.sys_id).decode()
        self.params = None if params is None else np.asarray(params, dtype=np.float64)

    @classmethod
    def builtin(cls, name, params=None):
        sid = _abi.load().ode_b200_system_builtin(name.encode())
        if sid < 0:
            raise Ode200Error(_abi.last_error())
        return cls(sid, params)

    @classmethod
    def from_source(cls, name, dim, n_params, rhs_body, solout_body=None, params=None):
        """Register CUDA source for `system` / `solout` (compiled by NVRTC on first use)."""
        sid = _abi.load().ode_b200_system_from_source(
            name.encode(), dim, n_params, rhs_body.encode(), solout_body.encode() if solout_body else None)
        if sid < 0:
            raise Ode200Error(_abi.last_error())
        return cls(sid, params)

    def compile(self, method, out_type=_abi.OUT_SPARSE):
        """NVRTC-compile this system's kernel for (method, out_type) now (no GPU needed)."""
        _check(_abi.load().ode_b200_system_compile(self.sys_id, int(method), int(out_type)), "system_compile")

    def with_params(self, params):
        return System(self.sys_id, params)


class DynSystem:
    """A system whose state dimension is only known at run time: the `DVector` instantiation of the
    crate (`System<f64, DVector<f64>>`, examples/kepler_orbit_dvector.rs:7,26,
    examples/three_body_system_dvector.rs:5,22; src/rk4.rs:272-306). The kernels keep the state in
    registers, so the dimension has to be a compile-time constant on the device: the first use with
    a state of length n compiles (NVRTC) and caches the kernels specialised for DIM = n; the source
    may use `DIM` (e.g. `for (int i = 0; i < DIM; ++i) ...`)."""

    def __init__(self, name, n_params, rhs_body, solout_body=None, params=None):
        self.name, self.n_params = name, int(n_params)
        self.rhs_body, self.solout_body = rhs_body, solout_body
        self.params = None if params is None else np.asarray(params, dtype=np.float64)
        self._by_dim = {}

    def resolve(self, dim):
        dim = int(dim)
        if dim < 1:
            raise ValueError("empty state")
        if dim not in self._by_dim:
            self._by_dim[dim] = System.from_source("%s_dim%d" % (self.name, dim), dim, self.n_params, self.rhs_body,
                                                   self.solout_body, self.params)
        return self._by_dim[dim]

    def with_params(self, params):
        d = DynSystem(self.name, self.n_params, self.rhs_body, self.solout_body, params)
        d._by_dim = {k: v.with_params(params) for k, v in self._by_dim.items()}
        return d


def _rust_display_f64(v):
    """`format!("{}", v)` of an f64: shortest digits that round-trip, never an exponent, no ".0"."""
    v = float(v)
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "inf" if v > 0 else "-inf"
    import decimal
    s = format(decimal.Decimal(repr(v)), "f")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    return s


def save(times, states, filename):
    """The `save()` of the crate's examples (examples/lorenz.rs:56-84 and the others): one line per
    stored point, `x, y0, y1, ...` with Rust's `{}` formatting of f64, parent directory created."""
    import os
    d = os.path.dirname(filename)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(filename, "w") as f:
        for t, st in zip(times, states):
            f.write(_rust_display_f64(t))
            for val in np.asarray(st, dtype=np.float64).reshape(-1):
                f.write(", " + _rust_display_f64(val))
            f.write("\n")


# ------------------------------------------------------------------ configs
def config_new(method, x, x_end, dx, rtol, atol, out_type=None):
    c = _abi.Config()
    _check(_abi.load().ode_b200_config_new(method, x, x_end, dx, rtol, atol, C.byref(c)), "config_new")
    if out_type is not None:
        c.out_type = int(out_type)
    return c


def config_from_param(method, x, x_end, dx, rtol, atol, safety_factor, beta, fac_min, fac_max, h_max, h, n_max,
                      n_stiff, out_type):
    c = _abi.Config()
    _check(_abi.load().ode_b200_config_from_param(method, x, x_end, dx, rtol, atol, safety_factor, beta, fac_min,
                                                  fac_max, h_max, h, n_max, n_stiff, int(out_type), C.byref(c)),
           "config_from_param")
    return c


def config_rk4(x, x_end, step_size):
    c = _abi.Config()
    _check(_abi.load().ode_b200_config_rk4(x, x_end, step_size, C.byref(c)), "config_rk4")
    return c


# ------------------------------------------------------------------ batched entry point
def integrate_batch(system, cfg, y0, params=None, out_cap=0, com_cap=0, ctx=None, n_gpus=1):
    """Integrate y0.shape[1] trajectories of `system` with one config.

    y0: [dim, n] (SoA); params: [n_params] shared or [n_params, n] per trajectory (defaults to
    system.params). Returns an _abi.HostOutputs (numpy arrays; layouts in include/ode_b200.h).
    """
    lib = _abi.load()
    if isinstance(system, DynSystem):  # runtime dimension: [dim, n] decides
        y0 = np.asarray(y0, dtype=np.float64)
        if y0.ndim != 2:
            raise ValueError("a DynSystem needs y0 of shape [dim, n]")
        system = system.resolve(y0.shape[0])
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float64).reshape(system.dim, -1))
    n = y0.shape[1]
    if params is None:
        params = system.params
    per_traj, pp = 0, None
    if system.n_params:
        if params is None:
            raise ValueError("system %s needs %d parameters" % (system.name, system.n_params))
        params = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
        per_traj = int(params.ndim == 2)
        if params.shape[0] != system.n_params or (per_traj and params.shape[1] != n):
            raise ValueError("bad params shape %r" % (params.shape,))
        pp = params.ctypes.data
    m = 5 if cfg.method == _abi.DOPRI5 else 8
    out = _abi.HostOutputs(n, system.dim, out_cap, com_cap, m)
    s = out.struct()
    if n_gpus == 1:
        ctx = ctx or default_context(0)
        rc = lib.ode_b200_integrate_batch(ctx.handle, system.sys_id, C.byref(cfg), n, y0.ctypes.data, pp, per_traj,
                                          C.byref(s))
        out.kernel_ms = ctx.last_kernel_ms
    else:
        rc = lib.ode_b200_integrate_batch_multi(system.sys_id, C.byref(cfg), n, y0.ctypes.data, pp, per_traj,
                                                int(n_gpus), C.byref(s))
    _check(rc, "integrate_batch")
    return out


def config_new_f32(method, x, x_end, dx, rtol, atol, out_type=None):
    """Dopri5::<f32>::new / Dop853::<f32>::new (dopri5.rs:76, dop853.rs:76 with T = f32)"""
    c = _abi.ConfigF32()
    _check(_abi.load().ode_b200_config_new_f32(method, x, x_end, dx, rtol, atol, C.byref(c)), "config_new_f32")
    if out_type is not None:
        c.out_type = out_type
    return c


def config_from_param_f32(method, x, x_end, dx, rtol, atol, safety_factor, beta, fac_min, fac_max, h_max, h, n_max,
                          n_stiff, out_type):
    c = _abi.ConfigF32()
    _check(_abi.load().ode_b200_config_from_param_f32(method, x, x_end, dx, rtol, atol, safety_factor, beta, fac_min,
                                                      fac_max, h_max, h, n_max, n_stiff, int(out_type), C.byref(c)),
           "config_from_param_f32")
    return c


def config_rk4_f32(x, x_end, step_size):
    c = _abi.ConfigF32()
    _check(_abi.load().ode_b200_config_rk4_f32(x, x_end, step_size, C.byref(c)), "config_rk4_f32")
    return c


def integrate_batch_f32(system, cfg, y0, params=None, out_cap=0, com_cap=0, ctx=None):
    """The three steppers with T = f32 (`impl FloatNumber for f32`, src/dop_shared.rs:74) for a batch.
    cfg: _abi.ConfigF32; y0: [dim, n] float32. Returns an _abi.HostOutputsF32."""
    lib = _abi.load()
    ctx = ctx or default_context(0)
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float32).reshape(system.dim, -1))
    n = y0.shape[1]
    if params is None:
        params = system.params
    per_traj, pp = 0, None
    if system.n_params:
        if params is None:
            raise ValueError("system %s needs %d parameters" % (system.name, system.n_params))
        params = np.ascontiguousarray(np.asarray(params, dtype=np.float32))
        per_traj = int(params.ndim == 2)
        pp = params.ctypes.data
    out = _abi.HostOutputsF32(n, system.dim, out_cap, com_cap, 5)
    s = out.struct()
    _check(lib.ode_b200_integrate_batch_f32(ctx.handle, system.sys_id, C.byref(cfg), n, y0.ctypes.data, pp, per_traj,
                                            C.byref(s)), "integrate_batch_f32")
    out.kernel_ms = ctx.last_kernel_ms
    return out


def rk4_f32_integrate_batch(system, x, x_end, step_size, y0, params=None, out_cap=0, ctx=None):
    """Rk4::<f32>::new(f, x, y, x_end, step_size) + integrate() for a batch (the ABI's convenience entry)."""
    lib = _abi.load()
    ctx = ctx or default_context(0)
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float32).reshape(system.dim, -1))
    n = y0.shape[1]
    if params is None:
        params = system.params
    per_traj, pp = 0, None
    if system.n_params:
        params = np.ascontiguousarray(np.asarray(params, dtype=np.float32))
        per_traj = int(params.ndim == 2)
        pp = params.ctypes.data
    out = _abi.HostOutputsF32(n, system.dim, out_cap)
    s = out.struct()
    _check(lib.ode_b200_rk4_f32_integrate_batch(ctx.handle, system.sys_id, x, x_end, step_size, n, y0.ctypes.data, pp,
                                                per_traj, C.byref(s)), "rk4_f32_integrate_batch")
    out.kernel_ms = ctx.last_kernel_ms
    return out


def com_evaluate_batch(out, xq, ctx=None):
    """Batched ContinuousOutputModel::evaluate on the device for every trajectory of `out`.
    Returns (y [n, nq, dim], found [n, nq] bool)."""
    lib = _abi.load()
    ctx = ctx or default_context(0)
    xq = np.ascontiguousarray(np.asarray(xq, dtype=np.float64))
    nq = xq.size
    y = np.zeros((out.n, nq, out.dim))
    found = np.zeros((out.n, nq), dtype=np.uint8)
    _check(lib.ode_b200_com_evaluate(ctx.handle, out.n, out.dim, out.m, out.com_cap, out.com_lower.ctypes.data,
                                     out.com_break.ctypes.data, out.com_step.ctypes.data,
                                     out.com_coef.ctypes.data, out.com_count.ctypes.data, nq, xq.ctypes.data,
                                     y.ctypes.data, found.ctypes.data), "com_evaluate")
    return y, found.astype(bool)


# ------------------------------------------------------------------ ContinuousOutputModel
class Interval:
    def __init__(self, coefficients, step_size):
        self.coefficients, self.step_size = coefficients, step_size


class ContinuousOutputModel:
    """src/continuous_output_model.rs:8-104; serde field names kept (lower_bound, breakpoints,
    intervals{coefficients, step_size}). `evaluate` runs on the device (batched kernel, N=1)."""

    def __init__(self):
        self.lower_bound = 0.0
        self.breakpoints = []
        self.intervals = []

    default = classmethod(lambda cls: cls())
    new = classmethod(lambda cls: cls())

    def set_lower_bound(self, lb):
        self.lower_bound = float(lb)

    def add_interval(self, breakpoint, coefficients, step_size):
        self.breakpoints.append(float(breakpoint))
        dt = np.float32 if all(getattr(c, "dtype", None) == np.float32 for c in coefficients) else np.float64
        self.intervals.append(Interval([np.asarray(c, dtype=dt) for c in coefficients], float(step_size)))

    def bounds(self):
        return self.lower_bound, self.breakpoints[-1]

    def evaluate(self, x, ctx=None):
        if not self.intervals:
            return None
        m = len(self.intervals[0].coefficients)
        if m == 0:
            raise IndexError("interval without coefficients")
        if self.intervals[0].coefficients[0].dtype == np.float32:
            return self._evaluate_f32(np.float32(x))
        dim = self.intervals[0].coefficients[0].size
        k = len(self.intervals)
        out = _abi.HostOutputs(1, dim, 0, k, m)
        out.com_lower[0] = self.lower_bound
        out.com_break[0, :] = self.breakpoints
        out.com_step[0, :] = [iv.step_size for iv in self.intervals]
        for q, iv in enumerate(self.intervals):
            out.com_coef[0, q] = np.stack(iv.coefficients)
        out.com_count[0] = k
        y, found = com_evaluate_batch(out, [x], ctx)
        return y[0, 0].copy() if found[0, 0] else None

    def _evaluate_f32(self, x):
        """continuous_output_model.rs:31-56,86-104 with T = f32 on the host (the device kernel is f64): every
        operation on np.float32 values, one rounding each"""
        import bisect
        f = np.float32
        if x < f(self.lower_bound) or x != x:
            return None
        idx = bisect.bisect_left([f(b) for b in self.breakpoints], x)
        if idx >= len(self.intervals):
            return None
        lower = f(self.lower_bound) if idx == 0 else f(self.breakpoints[idx - 1])
        iv = self.intervals[idx]
        theta = (x - lower) / f(iv.step_size)
        theta1 = f(1.0) - theta
        c = iv.coefficients
        r = c[-1].copy()
        for i in range(len(c) - 2, -1, -1):
            r = c[i] + r * (theta if i % 2 == 0 else theta1)
        return r

    def to_dict(self):
        return {"lower_bound": self.lower_bound, "breakpoints": list(self.breakpoints),
                "intervals": [{"coefficients": [c.tolist() for c in iv.coefficients], "step_size": iv.step_size}
                              for iv in self.intervals]}

    @classmethod
    def from_dict(cls, d):
        m = cls()
        m.lower_bound = float(d["lower_bound"])
        for b, iv in zip(d["breakpoints"], d["intervals"]):
            m.add_interval(b, iv["coefficients"], iv["step_size"])
        return m


# ------------------------------------------------------------------ single-trajectory steppers
class SolverResult:
    """src/dop_shared.rs:46,79-109"""

    def __init__(self, x=None, y=None):
        self.x = list(x) if x is not None else []
        self.y = list(y) if y is not None else []

    def get(self):
        return self.x, self.y

    def append(self, other):
        self.x += other.x
        self.y += other.y


def _raise_for_status(status, x, n_step):
    if status == _abi.MAX_NUM_STEP_REACHED:
        raise MaxNumStepReached(x, n_step)
    if status == _abi.STEP_SIZE_UNDERFLOW:
        raise StepSizeUnderflow(x)
    if status == _abi.STIFFNESS_DETECTED:
        raise StiffnessDetected(x)
    if status == _abi.UNSUPPORTED_DOMAIN:
        raise UnsupportedDomain("input outside the domain the reference handles (it panics or never returns)")


class _Stepper:
    method = None

    def __init__(self, f, cfg, y):
        # T = f32 (`impl FloatNumber for f32`, dop_shared.rs:74): the constructors' dtype=np.float32
        self.f32 = isinstance(cfg, _abi.ConfigF32)
        self.y0 = np.asarray(y, dtype=np.float32 if self.f32 else np.float64).reshape(-1)
        if isinstance(f, DynSystem):  # DVector state: its length is the dimension
            f = f.resolve(self.y0.size)
        self.f, self.cfg = f, cfg
        if self.y0.size != f.dim:
            raise ValueError("state has %d components, system %s needs %d" % (self.y0.size, f.name, f.dim))
        self._results = SolverResult()
        self.stats = Stats()
        self.h_next = 0.0

    def _guess_cap(self):
        c = self.cfg
        if c.method == _abi.RK4:
            return int(max(0.0, math.ceil((c.x_end - c.x0) / c.step_size))) + 2 if c.step_size else 2
        if c.out_type == _abi.OUT_DENSE and c.dx > 0:
            return int(min(max(0.0, (c.x_end - c.x0) / c.dx), 5e7)) + 3
        return 4096

    def _run(self, com=None, ctx=None):
        cap = self._guess_cap()
        want_com = com is not None
        while True:
            out = (integrate_batch_f32 if self.f32 else integrate_batch)(
                self.f, self.cfg, self.y0.reshape(-1, 1), out_cap=cap, com_cap=cap if want_com else 0, ctx=ctx)
            need = int(out.out_count[0])
            need_com = int(out.com_count[0]) if want_com else 0
            if need <= cap and need_com <= cap:
                break
            cap = max(need, need_com)
        n = int(out.out_count[0])
        self._results = SolverResult(out.out_x[0, :n].tolist(), [out.out_y[0, i].copy() for i in range(n)])
        self.stats = Stats(out.num_eval[0], out.accepted[0], out.rejected[0])
        self.h_next = float(out.h_next[0])
        self.x = float(out.x_final[0])
        self.y = out.y_final[:, 0].copy()
        if want_com:
            com.set_lower_bound(float(out.com_lower[0]))
            for q in range(int(out.com_count[0])):
                com.add_interval(out.com_break[0, q], [out.com_coef[0, q, r].copy() for r in range(out.m)],
                                 out.com_step[0, q])
        _raise_for_status(int(out.status[0]), self.x, int(out.n_step[0]))
        return self.stats

    def x_out(self):
        return self._results.x

    def y_out(self):
        return self._results.y

    def results(self):
        return self._results

    def into(self):
        """`From<stepper> for SolverResult` (dopri5.rs:513-522)"""
        return self._results


class Dopri5(_Stepper):
    """src/dopri5.rs"""

    @classmethod
    def new(cls, f, x, x_end, dx, y, rtol, atol, dtype=np.float64):
        mk = config_new_f32 if np.dtype(dtype) == np.float32 else config_new
        return cls(f, mk(_abi.DOPRI5, x, x_end, dx, rtol, atol), y)

    @classmethod
    def from_param(cls, f, x, x_end, dx, y, rtol, atol, safety_factor, beta, fac_min, fac_max, h_max, h, n_max,
                   n_stiff, out_type, dtype=np.float64):
        mk = config_from_param_f32 if np.dtype(dtype) == np.float32 else config_from_param
        return cls(f, mk(_abi.DOPRI5, x, x_end, dx, rtol, atol, safety_factor, beta, fac_min,
                         fac_max, h_max, h, n_max, n_stiff, out_type), y)

    def integrate(self, ctx=None):
        if self.cfg.out_type == _abi.OUT_CONTINUOUS:  # dopri5.rs:256-258 panics
            raise RuntimeError("Please use `integrate_with_continuous_output_model` to compute a continuous "
                               "output model.")
        return self._run(ctx=ctx)

    def integrate_with_continuous_output_model(self, continuous_output, ctx=None):
        self.cfg.out_type = _abi.OUT_CONTINUOUS  # dopri5.rs:250
        return self._run(com=continuous_output, ctx=ctx)


class Dop853(_Stepper):
    """src/dop853.rs (no ContinuousOutputModel support, as in the reference)"""

    @classmethod
    def new(cls, f, x, x_end, dx, y, rtol, atol, dtype=np.float64):
        mk = config_new_f32 if np.dtype(dtype) == np.float32 else config_new
        return cls(f, mk(_abi.DOP853, x, x_end, dx, rtol, atol), y)

    @classmethod
    def from_param(cls, f, x, x_end, dx, y, rtol, atol, safety_factor, beta, fac_min, fac_max, h_max, h, n_max,
                   n_stiff, out_type, dtype=np.float64):
        mk = config_from_param_f32 if np.dtype(dtype) == np.float32 else config_from_param
        return cls(f, mk(_abi.DOP853, x, x_end, dx, rtol, atol, safety_factor, beta, fac_min,
                         fac_max, h_max, h, n_max, n_stiff, out_type), y)

    def integrate(self, ctx=None):
        return self._run(ctx=ctx)

    def integrate_with_continuous_output_model(self, continuous_output, ctx=None):
        """EXTENSION (the reference has this for Dopri5 only): record Dop853's 8 dense-output
        vectors per accepted step as model intervals (ODE_B200_OUT_CONTINUOUS8)."""
        if self.f32:
            raise ValueError("the Continuous8 extension exists for f64 only")
        self.cfg.out_type = _abi.OUT_CONTINUOUS8
        return self._run(com=continuous_output, ctx=ctx)


class Rk4(_Stepper):
    """src/rk4.rs -- note the argument order (f, x, y, x_end, step_size)"""

    @classmethod
    def new(cls, f, x, y, x_end, step_size, dtype=np.float64):
        return cls(f, (config_rk4_f32 if np.dtype(dtype) == np.float32 else config_rk4)(x, x_end, step_size), y)

    def integrate(self, ctx=None):
        return self._run(ctx=ctx)
