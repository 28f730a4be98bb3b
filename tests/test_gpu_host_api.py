"""The crate-shaped host API (ode_solvers_b200.api) on the GPU, written like the reference's
own tests (src/rk4.rs:223-306, src/dopri5.rs:545-581, src/dop853.rs:609-620,
src/continuous_output_model.rs:121-252, tests/dopri5.rs) plus the batched extras: NVRTC-registered
user systems, device ContinuousOutputModel::evaluate, error statuses, the multi-GPU entry."""
import math

import numpy as np
import pytest

import ode_solvers_b200 as ob
import oracle
from ode_solvers_b200 import workloads as W
from parity_util import assert_same_outputs, same_f64

pytestmark = pytest.mark.gpu


def analytic(t):
    return -math.log(t ** 3 + 1.0)


# ---- src/rk4.rs tests
def test_rk4_test1_to_test4():
    s = ob.Rk4.new(ob.System.builtin("test_rk4_1"), 0.0, [1.0], 0.2, 0.1)
    s.integrate()
    assert abs(s.x_out()[-1] - 0.2) < 1e-8
    assert abs(s.y_out()[1][0] - 0.95369) < 1e-5 and abs(s.y_out()[2][0] - 0.91451) < 1e-5
    s = ob.Rk4.new(ob.System.builtin("test_rk4_2"), 0.0, [-1.0], 0.5, 0.1)
    s.integrate()
    assert abs(s.x_out()[-1] - 0.5) < 1e-8
    assert abs(s.y_out()[3][0] + 0.82246) < 1e-5 and abs(s.y_out()[5][0] + 0.81959) < 1e-5
    s = ob.Rk4.new(ob.System.builtin("test_exp"), 0.0, [1.0], 1.0, 0.1)
    st = s.integrate()
    out = s.y_out()
    assert abs(out[5][0] - 0.913059839) < 1e-9 and abs(out[8][0] - 0.9838057659) < 1e-9
    assert abs(out[10][0] - 1.0715783953) < 1e-9 and len(out) == 11
    assert (st.num_eval, st.accepted_steps, st.rejected_steps) == (40, 10, 0)
    s = ob.Rk4.new(ob.System.builtin("test_exp_stop", [0.5]), 0.0, [1.0], 1.0, 0.1)
    s.integrate()
    assert abs(s.x_out()[-1] - 0.5) < 1e-9 and abs(s.y_out()[5][0] - 0.913059839) < 1e-9 and len(s.y_out()) == 6


# ---- src/dopri5.rs tests
def test_dopri5_test1_and_continuous_panic():
    s = ob.Dopri5.new(ob.System.builtin("test_exp_stop", [0.5]), 0.0, 1.0, 0.1, [1.0], 1e-12, 1e-6)
    s.integrate()
    assert abs(s.x_out()[-1] - 0.5) < 1e-9
    assert abs(s.y_out()[5][0] - 0.9130611474392001) < 1e-9
    assert s.y_out()[5][0] == 0.9130611474392001
    s = ob.Dopri5.from_param(ob.System.builtin("test_exp_stop", [0.5]), 0.0, 1.0, 0.1, [1.0], 1e-12, 1e-6, 0.1, 0.2,
                             0.3, 2.0, 5.0, 0.1, 10000, 1000, ob.OutputType.Continuous)
    with pytest.raises(RuntimeError):  # #[should_panic], dopri5.rs:558-581
        s.integrate()


# ---- src/dop853.rs test
def test_dop853_test1():
    s = ob.Dop853.new(ob.System.builtin("test_exp_stop", [0.5]), 0.0, 1.0, 0.1, [1.0], 1e-12, 1e-6)
    s.integrate()
    assert abs(s.x_out()[-1] - 0.5) < 1e-9
    assert abs(s.y_out()[5][0] - 0.912968195) < 1e-9


# ---- tests/dopri5.rs
def test_integration_dense():
    s = ob.Dopri5.new(ob.System.builtin("test_cubic"), 0.0, 6.0, 0.1, [0.0], 1e-10, 1e-10)
    s.integrate()
    ts, ys = s.results().get()
    assert len(ts) == 61 and len(ys) == 61
    for t, y in zip(ts, ys):
        assert abs(y[0] - analytic(t)) < 1e-8


def test_integration_continuous_output_model():
    s = ob.Dopri5.new(ob.System.builtin("test_cubic"), 0.0, 6.0, 0.1, [0.0], 1e-10, 1e-10)
    com = ob.ContinuousOutputModel.default()
    s.integrate_with_continuous_output_model(com)
    assert com.evaluate(-0.0001) is None and com.evaluate(0.0) is not None
    assert com.evaluate(6.0) is not None and com.evaluate(6.0001) is None
    for i in range(100):
        t = i * (6.0 / 100)
        assert abs(com.evaluate(t)[0] - analytic(t)) < 1e-9
    # serde-shaped round trip keeps every bit
    back = ob.ContinuousOutputModel.from_dict(com.to_dict())
    assert back.evaluate(3.21)[0] == com.evaluate(3.21)[0]


def test_integration_continuous_output_model_and_stopping_condition():
    s = ob.Dopri5.new(ob.System.builtin("test_cubic_stop", [4.1]), 0.0, 6.0, 0.1, [0.0], 1e-10, 1e-10)
    com = ob.ContinuousOutputModel.default()
    s.integrate_with_continuous_output_model(com)
    assert com.evaluate(-0.0001) is None and com.evaluate(0.0) is not None
    assert com.evaluate(4.1) is not None and com.evaluate(4.1 + 0.1) is None
    assert 4.1 <= com.bounds()[1] < 4.2
    for i in range(100):
        t = i * ((6.0 - 4.1) / 100)
        assert abs(com.evaluate(t)[0] - analytic(t)) < 1e-9


# ---- src/continuous_output_model.rs tests, evaluated by the device kernel
def test_com_unit_tests_on_device():
    m = ob.ContinuousOutputModel.default()
    m.add_interval(2.0, [[0.2], [1.0], [-2.5]], 0.1)
    m.add_interval(5.0, [[-1.5], [0.4], [1.2]], 0.2)
    assert m.evaluate(-0.1) is None and m.evaluate(5.1) is None
    for x, want in ((0.0, 0.2), (1.2, 342.2), (2.0, 970.2), (2.5, -5.0), (5.0, -247.5)):
        assert abs(m.evaluate(x)[0] - want) < 1e-9 * max(1.0, abs(want))
    m = ob.ContinuousOutputModel.default()
    m.add_interval(2.0, [[0.2], [1.0], [-2.5], [0.3]], 0.1)
    m.add_interval(5.0, [[-1.5], [0.4], [1.2], [2.7]], 0.2)
    for x, want in ((0.0, 0.2), (1.2, -133.0), (2.0, -1309.8), (2.5, -30.3125), (5.0, -8752.5)):
        assert abs(m.evaluate(x)[0] - want) < 1e-9 * max(1.0, abs(want))
    assert ob.ContinuousOutputModel.default().evaluate(0.0) is None


def test_batched_com_evaluate_matches_oracle_bitwise():
    n = 200
    y0, p = W.lorenz_ensemble(n, offset=3)
    cfg = ob.config_new(ob.DOPRI5, 0.0, 3.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_CONTINUOUS)
    g = ob.integrate_batch(ob.System.builtin("lorenz"), cfg, y0, p, out_cap=400, com_cap=400)
    assert (g.com_count <= 400).all()
    xq = np.linspace(-0.5, 3.5, 97)
    y, found = ob.com_evaluate_batch(g, xq)
    for i in range(0, n, 7):
        for q, x in enumerate(xq):
            r = oracle.com_evaluate(g, i, x)
            assert (r is not None) == bool(found[i, q])
            if r is not None:
                assert same_f64(y[i, q], r).all()


# ---- user systems through NVRTC (the `System` trait implemented by the user)
def test_user_system_from_source_matches_builtin_bitwise():
    f = ob.System.from_source(
        "user_lorenz", 3, 3,
        "dy[0] = p[0] * (y[1] - y[0]);\n dy[1] = y[0] * (p[2] - y[2]) - y[1];\n dy[2] = y[0] * y[1] - p[1] * y[2];")
    y0, p = W.lorenz_ensemble(300, offset=11)
    for cfg in (ob.config_new(ob.DOPRI5, 0.0, 4.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_SPARSE),
                ob.config_new(ob.DOP853, 0.0, 4.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_SPARSE),
                ob.config_new(ob.DOPRI5, 0.0, 4.0, 0.1, 1e-6, 1e-6), ob.config_rk4(0.0, 1.0, 1e-3)):
        cap = 60 if cfg.out_type == ob.OUT_DENSE and cfg.method != ob.RK4 else 0
        a = ob.integrate_batch(f, cfg, y0, p, out_cap=cap)
        b = ob.integrate_batch(ob.System.builtin("lorenz"), cfg, y0, p, out_cap=cap)
        assert_same_outputs(a, b, "jit vs builtin m%d" % cfg.method)
    # Rk4 with every step recorded: the NVRTC build of the warp-level tensor stores (full warps) next to the per-lane
    # copies (the last, partial warp), capacity no multiple of the group
    cfg = ob.config_rk4(0.0, 0.013, 1e-3)
    a = ob.integrate_batch(f, cfg, y0, p, out_cap=14)
    b = ob.integrate_batch(ob.System.builtin("lorenz"), cfg, y0, p, out_cap=14)
    assert_same_outputs(a, b, "jit vs builtin, Rk4 recorded")
    assert (a.out_count == 14).all() and (a.status == ob.OK).all()


def test_user_system_div_d_sqrt_d_match_plain_operators_and_builtin():
    """div_d / sqrt_d route a user RHS through the kernels' speculative exact arithmetic (one
    straight-line block, lockstep CTA for n = 6 Dop853): same bits as `/` and sqrt(), and as the
    built-in Kepler system"""
    spec = ("const double r = sqrt_d(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);\n const double r3 = r * r * r;\n"
            "dy[0] = y[3]; dy[1] = y[4]; dy[2] = y[5];\n"
            "dy[3] = div_d(-p[0] * y[0], r3); dy[4] = div_d(-p[0] * y[1], r3); dy[5] = div_d(-p[0] * y[2], r3);")
    plain = spec.replace("sqrt_d(", "sqrt(").replace("div_d(-p[0] * y[0], r3)", "-p[0] * y[0] / r3").replace(
        "div_d(-p[0] * y[1], r3)", "-p[0] * y[1] / r3").replace("div_d(-p[0] * y[2], r3)", "-p[0] * y[2] / r3")
    assert "div_d" not in plain and "sqrt_d" not in plain
    fs = ob.System.from_source("user_kepler_spec", 6, 1, spec)
    fp = ob.System.from_source("user_kepler_plain", 6, 1, plain)
    y0, p = W.kepler_sweep(700, offset=21)
    for cfg in (ob.config_new(ob.DOP853, 0.0, W.kepler_period(), 60.0, 1e-10, 1e-10, out_type=ob.OUT_SPARSE),
                ob.config_new(ob.DOPRI5, 0.0, W.kepler_period(), 600.0, 1e-8, 1e-8),
                ob.config_rk4(0.0, 2000.0, 5.0)):
        cap = 30 if cfg.out_type == ob.OUT_DENSE and cfg.method != ob.RK4 else 0
        b = ob.integrate_batch(ob.System.builtin("kepler"), cfg, y0, p, out_cap=cap)
        for f in (fs, fp):
            a = ob.integrate_batch(f, cfg, y0, p, out_cap=cap)
            assert_same_outputs(a, b, "jit kepler m%d" % cfg.method)


def test_out_of_window_operands_take_the_exact_path_with_identical_bits():
    """Kepler in units that push numerators below 2^-969 and reciprocals out of the normal range:
    the speculative block flags every such attempt and the out-of-line exact recomputation must
    reproduce the oracle (plain IEEE `/` and sqrt) bit for bit; mixed with ordinary trajectories in the
    same warps"""
    n = 512
    y0, p = W.kepler_sweep(n, offset=3)
    y0 = y0.copy()
    p = np.array(p, dtype=np.float64).reshape(1, -1) * np.ones((1, n))
    # rescale lengths by s and mu by s^3 (same orbits, same periods). s = 1e-76: numerators
    # mu * y ~ 3e-295 < 2^-969 (normal numbers, but below the fast division's window) for every third
    # trajectory; s = 1e70 (large but inside the window) for every fifth
    for sel, s in ((slice(0, n, 3), 1e-76), (slice(1, n, 5), 1e70)):
        y0[:, sel] *= s
        p[:, sel] *= s ** 3
    for cfg in (ob.config_new(ob.DOP853, 0.0, 0.5 * W.kepler_period(), 60.0, 1e-10, 1e-10, out_type=ob.OUT_SPARSE),
                ob.config_new(ob.DOPRI5, 0.0, 0.25 * W.kepler_period(), 600.0, 1e-8, 1e-10),
                ob.config_rk4(0.0, 500.0, 5.0)):
        cap = 30 if cfg.out_type == ob.OUT_DENSE and cfg.method != ob.RK4 else 0
        g = ob.integrate_batch(ob.System.builtin("kepler"), cfg, y0, p, out_cap=cap)
        o = oracle.integrate_batch("kepler", cfg, y0, p, out_cap=cap)
        assert_same_outputs(g, o, "kepler rescaled m%d" % cfg.method)
        assert int(g.accepted[0::3].min()) > 3


def test_user_system_recording_every_step_uses_the_staged_kernels():
    """Sparse / Rk4 recording goes through the STAGE kernel variants (shared-memory runs); the NVRTC
    build of those must match the ahead-of-time one bit for bit"""
    f = ob.System.from_source(
        "user_lorenz_staged", 3, 3,
        "dy[0] = p[0] * (y[1] - y[0]);\n dy[1] = y[0] * (p[2] - y[2]) - y[1];\n dy[2] = y[0] * y[1] - p[1] * y[2];")
    y0, p = W.lorenz_ensemble(200, offset=4)
    for cfg, cap in ((ob.config_new(ob.DOPRI5, 0.0, 2.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_SPARSE), 200),
                     (ob.config_new(ob.DOP853, 0.0, 2.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_CONTINUOUS8), 120),
                     (ob.config_rk4(0.0, 0.3, 1e-3), 310)):
        a = ob.integrate_batch(f, cfg, y0, p, out_cap=cap, com_cap=cap if cfg.out_type == ob.OUT_CONTINUOUS8 else 0)
        b = ob.integrate_batch(ob.System.builtin("lorenz"), cfg, y0, p, out_cap=cap,
                               com_cap=cap if cfg.out_type == ob.OUT_CONTINUOUS8 else 0)
        assert_same_outputs(a, b, "jit staged m%d" % cfg.method)
        assert (a.out_count > 10).all()


def test_user_system_with_solout_and_exp():
    f = ob.System.from_source("user_exp_stop", 1, 1, "dy[0] = (5.0 * x * x - y[0]) / exp_d(x + y[0]);",
                              "return x >= p[0];")
    s = ob.Dopri5.new(f.with_params([0.5]), 0.0, 1.0, 0.1, [1.0], 1e-12, 1e-6)
    s.integrate()
    assert s.y_out()[5][0] == 0.9130611474392001 and abs(s.x_out()[-1] - 0.5) < 1e-9


def test_user_system_compile_error_is_reported():
    f = ob.System.from_source("broken", 1, 0, "dy[0] = undefined_symbol;")
    with pytest.raises(ob.Ode200Error) as e:
        ob.integrate_batch(f, ob.config_rk4(0.0, 1.0, 0.1), np.ones((1, 4)))
    assert "undefined_symbol" in str(e.value)


# ---- IntegrationError variants
def test_integration_errors():
    lor = ob.System.builtin("lorenz", W.LORENZ_PARAMS)
    s = ob.Dopri5.from_param(lor, 0.0, 20.0, 0.1, [1.0, 1.0, 1.0], 1e-6, 1e-6, 0.9, 0.04, 0.2, 10.0, 20.0, 0.0, 50,
                             1000, ob.OutputType.Sparse)
    with pytest.raises(ob.MaxNumStepReached) as e:
        s.integrate()
    assert e.value.n_step == 51 and len(s.x_out()) > 1  # partial results stay readable
    rob = ob.System.builtin("robertson", W.ROBERTSON_K)
    s = ob.Dopri5.from_param(rob, 0.0, 40.0, 1.0, [1.0, 0.0, 0.0], 1e-6, 1e-8, 0.9, 0.04, 0.2, 10.0, 40.0, 0.0,
                             100000, 10, ob.OutputType.Sparse)
    with pytest.raises(ob.StiffnessDetected):
        s.integrate()
    s = ob.Dopri5.from_param(rob, 0.0, 40.0, 1.0, [1.0, 0.0, 0.0], 1e-6, 1e-8, 0.9, 0.04, 0.2, 10.0, 40.0, 0.0,
                             100000, 0, ob.OutputType.Sparse)
    with pytest.raises(ob.UnsupportedDomain):
        s.integrate()


def test_error_and_nonfinite_paths_match_oracle_bitwise():
    """blow-up ODE y' = y^2 style behaviour via Robertson with huge rate constants and tight n_max:
    rejected-step shrinking, underflow and max-step exits must agree bit for bit."""
    n = 256
    y0, k = W.robertson_sweep(n, offset=77)
    k = k * np.array([[1.0], [1e6], [1e8]])
    for method in (ob.DOPRI5, ob.DOP853):
        cfg = ob.config_from_param(method, 0.0, 1e6, 1.0, 1e-9, 1e-12, 0.9, 0.04 if method == ob.DOPRI5 else 0.0, 0.2,
                                   10.0, 1e6, 0.0, 3000, 20, ob.OUT_SPARSE)
        g = ob.integrate_batch(ob.System.builtin("robertson"), cfg, y0, k)
        o = oracle.integrate_batch("robertson", cfg, y0, k)
        assert_same_outputs(g, o, "robertson stiff m%d" % method)
        assert (g.status != ob.OK).any()
    # non-finite initial state: every attempt is rejected until the step size underflows
    y0 = np.array([[np.inf, np.nan, 1e308, 1.0]] * 3)
    cfg = ob.config_new(ob.DOPRI5, 0.0, 1.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_SPARSE)
    g = ob.integrate_batch(ob.System.builtin("lorenz"), cfg, y0, W.LORENZ_PARAMS)
    o = oracle.integrate_batch("lorenz", cfg, y0, W.LORENZ_PARAMS)
    assert_same_outputs(g, o, "non-finite y0")


def test_multi_gpu_entry_with_one_gpu_matches_single():
    y0, p = W.kepler_sweep(777, offset=9)
    cfg = ob.config_new(ob.DOP853, 0.0, W.kepler_period(), 60.0, 1e-10, 1e-10, out_type=ob.OUT_SPARSE)
    f = ob.System.builtin("kepler")
    a = ob.integrate_batch(f, cfg, y0, p)
    lib = ob._abi.load()
    import ctypes as C
    out = ob._abi.HostOutputs(777, 6)
    s = out.struct()
    rc = lib.ode_b200_integrate_batch_multi(f.sys_id, C.byref(cfg), 777, y0.ctypes.data, p.ctypes.data, 0, 1,
                                            C.byref(s))
    assert rc == 0
    assert_same_outputs(a, out, "multi(1) vs single")
    if ob.device_count() >= 2:
        out2 = ob._abi.HostOutputs(777, 6)
        s2 = out2.struct()
        assert lib.ode_b200_integrate_batch_multi(f.sys_id, C.byref(cfg), 777, y0.ctypes.data, p.ctypes.data, 0, 2,
                                                  C.byref(s2)) == 0
        assert_same_outputs(a, out2, "multi(2) vs single")


def test_pinned_host_outputs_take_the_zero_copy_path_with_identical_results():
    """pinned (device-mapped) output arrays are written directly by the kernel; pageable ones go
    through the staged copy: same bits either way"""
    import ctypes as C
    import torch
    from ode_solvers_b200 import _abi
    n = 5000
    y0, p = W.lorenz_ensemble(n, offset=99)
    cfg = ob.config_new(ob.DOPRI5, 0.0, 3.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_SPARSE)
    f = ob.System.builtin("lorenz")
    ref = ob.integrate_batch(f, cfg, y0, p)  # numpy (pageable) outputs
    pin = {k: torch.zeros(n, dtype=torch.float64).pin_memory() for k in ("x_final", "h_next")}
    pin["y_final"] = torch.zeros((3, n), dtype=torch.float64).pin_memory()
    for k in ("num_eval", "accepted", "rejected", "n_step", "status"):
        pin[k] = torch.full((n,), -7, dtype=torch.int32).pin_memory()
    o = _abi.Outputs()
    for k in ("y_final", "x_final", "h_next"):
        setattr(o, k, C.cast(pin[k].data_ptr(), _abi.c_double_p))
    for k in ("num_eval", "accepted", "rejected", "n_step"):
        setattr(o, k, C.cast(pin[k].data_ptr(), _abi.c_u32_p))
    o.status = C.cast(pin["status"].data_ptr(), _abi.c_i32_p)
    ctx = ob.default_context(0)
    pp = np.ascontiguousarray(p)
    rc = _abi.load().ode_b200_integrate_batch(ctx.handle, f.sys_id, C.byref(cfg), n, y0.ctypes.data, pp.ctypes.data,
                                             0, C.byref(o))
    assert rc == 0

    def check():
        assert same_f64(pin["y_final"].numpy(), ref.y_final).all()
        assert same_f64(pin["x_final"].numpy(), ref.x_final).all()
        assert same_f64(pin["h_next"].numpy(), ref.h_next).all()
        for k in ("num_eval", "accepted", "rejected", "n_step"):
            assert np.array_equal(pin[k].numpy().view(np.uint32), getattr(ref, k))
        assert np.array_equal(pin["status"].numpy(), ref.status)

    check()
    # ode_b200_set_zero_copy(ctx, 0): the same pinned arrays filled by DMA copies instead
    for t in pin.values():
        t.fill_(-7)
    ctx2 = ob.Context(0)
    ctx2.set_zero_copy(False)
    rc = _abi.load().ode_b200_integrate_batch(ctx2.handle, f.sys_id, C.byref(cfg), n, y0.ctypes.data,
                                             pp.ctypes.data, 0, C.byref(o))
    assert rc == 0
    check()
    ctx2.close()


def test_dop853_continuous_output_model_extension():
    """BASELINE config 4: Dop853 'with dense output into ContinuousOutputModel' -- an extension (the
    reference builds the model for Dopri5 only). The model must reproduce Dop853's own dense samples."""
    f = ob.System.builtin("cr3bp", [W.CR3BP_MU])
    dense = ob.Dop853.new(f, 0.0, 5.0, 0.25, W.CR3BP_Y0, 1e-12, 1e-12)
    st_d = dense.integrate()
    s = ob.Dop853.new(f, 0.0, 5.0, 0.25, W.CR3BP_Y0, 1e-12, 1e-12)
    com = ob.ContinuousOutputModel.default()
    st_c = s.integrate_with_continuous_output_model(com)
    assert st_c == st_d  # same numerics as Dense mode (incl. the 3 extra stages per accepted step)
    assert len(com.intervals) == st_c.accepted_steps and len(com.intervals[0].coefficients) == 8
    assert com.bounds() == (0.0, 5.0) and com.evaluate(-1e-3) is None and com.evaluate(5.001) is None
    same = 0
    for x, y in zip(dense.x_out()[1:-1], dense.y_out()[1:-1]):
        v = com.evaluate(x)
        assert np.allclose(v, y, rtol=1e-13, atol=1e-15)
        same += int(np.array_equal(v, y))
    assert same >= len(dense.x_out()) - 4
    # Dopri5 has no such mode
    cfg = ob.config_new(ob.DOPRI5, 0.0, 1.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_CONTINUOUS8)
    with pytest.raises(ob.Ode200Error):
        ob.integrate_batch(f, cfg, np.asarray(W.CR3BP_Y0).reshape(6, 1))


def test_com_evaluate_device_entry_unsorted_nan_large_capacity_and_generic_shape():
    """ode_b200_com_evaluate / _device after the round-2 rewrite (one warp per model, breakpoints staged in shared
    memory): unsorted queries, queries outside the bounds, NaN, a capacity beyond the staging limit (breakpoints read
    from global memory), the (m, dim) shapes without a compile-time instantiation, and the device-pointer entry"""
    import ctypes as C
    import torch
    from ode_solvers_b200 import _abi
    lib = _abi.load()
    rng = np.random.default_rng(3)
    # (a) Lorenz / Dopri5, ragged interval counts, small and huge capacity
    y0, p = W.lorenz_ensemble(150, offset=9)
    cfg = ob.config_new(ob.DOPRI5, 0.0, 3.0, 0.1, 1e-7, 1e-7, out_type=ob.OUT_CONTINUOUS)
    for cap in (400, 2000):
        g = ob.integrate_batch(ob.System.builtin("lorenz"), cfg, y0, p, out_cap=cap, com_cap=cap)
        xq = np.concatenate([rng.uniform(-0.5, 3.5, 97), [0.0, 3.0, np.nan, -0.0, 1e300], g.com_break[0, :5]])
        y, found = ob.com_evaluate_batch(g, xq)
        for i in range(0, g.n, 7):
            for q, x in enumerate(xq):
                r = oracle.com_evaluate(g, i, x) if x == x else None
                assert (r is not None) == bool(found[i, q]), (cap, i, x)
                if r is not None:
                    assert same_f64(y[i, q], r).all(), (cap, i, x)
    # (b) a 5-dimensional user system (no compile-time (5, 5) instantiation) and Dop853's 8-row intervals of dim 1
    f5 = ob.System.from_source("decay5", 5, 1, "for (int i = 0; i < 5; ++i) dy[i] = -p[0] * (i + 1) * y[i];", params=[0.7])
    g5 = ob.integrate_batch(f5, ob.config_new(ob.DOPRI5, 0.0, 2.0, 0.1, 1e-8, 1e-8, out_type=ob.OUT_CONTINUOUS),
                            rng.uniform(0.5, 2.0, (5, 40)), out_cap=300, com_cap=300)
    g1 = ob.integrate_batch(ob.System.builtin("test_exp"), ob.config_new(ob.DOP853, 0.0, 1.0, 0.1, 1e-4, 1e-4,
                                                                        out_type=ob.OUT_CONTINUOUS8),
                            np.linspace(0.5, 1.5, 40).reshape(1, -1), out_cap=300, com_cap=300)
    for g in (g5, g1):
        xq = rng.uniform(0.0, g.x_final[0], 33)
        y, found = ob.com_evaluate_batch(g, xq)
        assert found.any()
        for i in range(0, g.n, 3):
            for q, x in enumerate(xq):
                r = oracle.com_evaluate(g, i, x)  # (a model cut off at com_cap answers None beyond its last interval)
                assert (r is not None) == bool(found[i, q])
                if r is not None:
                    assert same_f64(y[i, q], r).all()
    # (c) the device-pointer entry on the model of (a)
    dev = torch.device("cuda", 0)
    ctx = ob.default_context(0)
    xq = np.sort(rng.uniform(0.0, 3.0, 500))
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    g = ob.integrate_batch(ob.System.builtin("lorenz"), cfg, y0, p, out_cap=400, com_cap=400)
    d = {k: up(getattr(g, k)) for k in ("com_lower", "com_break", "com_step", "com_coef", "com_count")}
    d_xq, d_y = up(xq), torch.zeros((g.n, 500, 3), dtype=torch.float64, device=dev)
    d_f = torch.zeros((g.n, 500), dtype=torch.uint8, device=dev)
    rc = lib.ode_b200_com_evaluate_device(ctx.handle, g.n, 3, 5, 400, d["com_lower"].data_ptr(), d["com_break"].data_ptr(),
                                          d["com_step"].data_ptr(), d["com_coef"].data_ptr(), d["com_count"].data_ptr(), 500,
                                          d_xq.data_ptr(), d_y.data_ptr(), d_f.data_ptr(), lib.ode_b200_stream(ctx.handle))
    assert rc == 0, _abi.last_error()
    torch.cuda.synchronize()
    yh, fh = ob.com_evaluate_batch(g, xq)
    assert np.array_equal(d_f.cpu().numpy().astype(bool), fh) and same_f64(d_y.cpu().numpy()[fh], yh[fh]).all()
    assert ctx.last_kernel_ms > 0.0
