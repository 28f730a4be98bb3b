"""Rk4 with T = f32 (`impl FloatNumber for f32`, src/dop_shared.rs:74; the type
examples/bouncing_ball.rs:15 runs in): CUDA kernel vs the f32 oracle, bit for bit."""
import numpy as np
import pytest

import ode_solvers_b200 as ob
import oracle
from ode_solvers_b200 import workloads as W

pytestmark = pytest.mark.gpu


def same_f32(a, b):
    a, b = np.ascontiguousarray(a, dtype=np.float32), np.ascontiguousarray(b, dtype=np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def check(system, x, x_end, step, y0, params, cap):
    g = ob.rk4_f32_integrate_batch(ob.System.builtin(system), x, x_end, step, y0, params, out_cap=cap)
    o = oracle.rk4_f32_integrate_batch(system, x, x_end, step, y0, params, out_cap=cap)
    for k in ("status", "accepted", "num_eval", "out_count"):
        assert np.array_equal(getattr(g, k), getattr(o, k)), k
    assert same_f32(g.y_final, o.y_final).all() and same_f32(g.x_final, o.x_final).all()
    if cap:
        for i in range(g.n):
            c = int(min(g.out_count[i], cap))
            assert same_f32(g.out_x[i, :c], o.out_x[i, :c]).all() and same_f32(g.out_y[i, :c], o.out_y[i, :c]).all()
    return g


def test_bouncing_ball_example_in_f32():
    """examples/bouncing_ball.rs:24-37 literal: State = Vector3<f32>, Rk4::new(system, 0f32, y0, 10f32, 0.01f32)"""
    g = check("bouncing_ball", 0.0, 10.0, 0.01, np.array([[10.0], [0.0], [0.0]]), [9.81], 1100)
    assert g.accepted[0] == 143 and g.out_count[0] == 144 and g.y_final[0, 0] < 0.0
    y0, p = W.bouncing_ball_ensemble(3000, offset=1)
    g = check("bouncing_ball", 0.0, 10.0, 0.01, y0, p, 1010)
    assert len(np.unique(g.accepted)) > 100


def test_f32_systems_bitwise():
    y0, p = W.lorenz_ensemble(700, offset=2)
    check("lorenz", 0.0, 2.0, 1e-3, y0, p, 0)
    check("lorenz", 0.0, 0.05, 1e-3, y0, p, 60)
    y0, k = W.robertson_sweep(300)
    check("robertson", 0.0, 0.01, 1e-5, y0, k, 0)
    g = check("test_rk4_1", 0.0, 0.2, 0.1, np.ones((1, 40)), None, 8)
    assert abs(g.out_y[0, 1, 0] - 0.95369) < 1e-5 and abs(g.out_y[0, 2, 0] - 0.91451) < 1e-5  # rk4.rs:230-232
    g = check("test_rk4_2", 0.0, 0.5, 0.1, -np.ones((1, 40)), None, 8)
    assert abs(g.out_y[0, 3, 0] + 0.82246) < 1e-5 and abs(g.out_y[0, 5, 0] + 0.81959) < 1e-5  # rk4.rs:242-244


def test_f32_only_where_the_rhs_has_an_f32_form():
    with pytest.raises(ob.Ode200Error) as e:
        ob.rk4_f32_integrate_batch(ob.System.builtin("kepler"), 0.0, 1.0, 0.1, np.ones((6, 4)), [1.0])
    assert "no f32 instantiation" in str(e.value)


@pytest.mark.parametrize("seed", range(40))
def test_f32_random_configurations_bitwise(seed):
    """random systems, spans, step sizes (incl. negative and non-dividing ones), capacities and
    non-finite states for the f32 fixed-step path"""
    rng = np.random.default_rng(77 + seed)
    system = str(rng.choice(["lorenz", "robertson", "bouncing_ball", "test_rk4_1", "test_rk4_2"]))
    n = int(rng.choice([1, 33, 200]))
    if system == "lorenz":
        y0, p = W.lorenz_ensemble(n, offset=seed)
        span = 1.0
    elif system == "robertson":
        y0, p = W.robertson_sweep(n, offset=seed)
        span = 0.01
    elif system == "bouncing_ball":
        y0, p = W.bouncing_ball_ensemble(n, offset=seed)
        span = 5.0
    else:
        y0, p = rng.uniform(-2, 2, (1, n)), None
        span = 1.0
    steps = int(rng.choice([3, 10, 250, 1000]))
    step = span / steps * float(rng.choice([1.0, 1.0, 1.0001, 0.37])) * (1.0 if rng.random() < 0.9 else -1.0)
    x0 = float(rng.choice([0.0, 0.25 * span]))
    if rng.random() < 0.1:
        y0 = y0.copy()
        y0[0, ::5] = rng.choice([np.inf, np.nan, 3e38, -3e38, 1e-45])
    cap = int(rng.choice([0, 4, steps + 2, 2 * steps + 5]))
    check(system, x0, x0 + span, step, y0, p, cap)


# ------------------------------------------------------------------ Dopri5 / Dop853 with T = f32
def check_adaptive(system, cfg, y0, params, cap=0, com_cap=0, label=""):
    g = ob.integrate_batch_f32(ob.System.builtin(system), cfg, y0, params, out_cap=cap, com_cap=com_cap)
    o = oracle.integrate_batch_f32(system, cfg, y0, params, out_cap=cap, com_cap=com_cap)
    for k in ("status", "accepted", "rejected", "num_eval", "n_step", "out_count", "com_count"):
        ga, oa = getattr(g, k), getattr(o, k)
        bad = np.nonzero(ga != oa)[0]
        assert bad.size == 0, "%s: %s differs for %d trajectories, first %d: gpu %s oracle %s" % (
            label, k, bad.size, bad[0], ga[bad[0]], oa[bad[0]])
    for k in ("y_final", "x_final", "h_next", "com_lower"):
        assert same_f32(getattr(g, k), getattr(o, k)).all(), "%s: %s" % (label, k)
    if cap:
        for i in range(g.n):
            c = int(min(g.out_count[i], cap))
            assert same_f32(g.out_x[i, :c], o.out_x[i, :c]).all() and same_f32(g.out_y[i, :c], o.out_y[i, :c]).all(), label
    if com_cap:
        for i in range(g.n):
            c = int(min(g.com_count[i], com_cap))
            for k in ("com_break", "com_step", "com_coef"):
                assert same_f32(getattr(g, k)[i, :c], getattr(o, k)[i, :c]).all(), "%s: %s" % (label, k)
    return g


@pytest.mark.parametrize("method", [ob.DOPRI5, ob.DOP853])
@pytest.mark.parametrize("out_type", [ob.OUT_DENSE, ob.OUT_SPARSE, ob.OUT_CONTINUOUS])
def test_f32_adaptive_steppers_bitwise(method, out_type):
    """every f32-capable system x Dopri5 / Dop853 x every output mode against the f32 oracle, bit for bit
    (identical accepted / rejected step counts included): the controller's two powf per attempt go
    through the libm-exact clone"""
    com = 400 if (out_type == ob.OUT_CONTINUOUS and method == ob.DOPRI5) else 0
    y0, p = W.lorenz_ensemble(300, offset=11)
    cfg = ob.config_new_f32(method, 0.0, 2.0, 0.05, 1e-4, 1e-4, out_type=out_type)
    g = check_adaptive("lorenz", cfg, y0, p, 400, com, "lorenz m%d o%d" % (method, out_type))
    assert (g.status == ob.OK).all() and g.accepted.min() > 10
    y0, k = W.robertson_sweep(200, offset=3)
    cfg = ob.config_new_f32(method, 0.0, 0.3, 0.03, 1e-2, 1e-6, out_type=out_type)  # chemical_reaction.rs:18
    check_adaptive("robertson", cfg, y0, k, 300, com, "robertson m%d o%d" % (method, out_type))
    y0, p = W.bouncing_ball_ensemble(400, offset=8)
    cfg = ob.config_new_f32(method, 0.0, 10.0, 0.01, 1e-2, 1e-6, out_type=out_type)  # bouncing_ball.rs:35
    g = check_adaptive("bouncing_ball", cfg, y0, p, 1010, com, "ball m%d o%d" % (method, out_type))
    assert (g.y_final[0] < 0).all()
    for system, y in (("test_rk4_1", np.linspace(0.5, 1.5, 50)), ("test_rk4_2", -np.linspace(0.5, 1.5, 50))):
        cfg = ob.config_new_f32(method, 0.0, 1.0, 0.1, 1e-5, 1e-5, out_type=out_type)
        check_adaptive(system, cfg, y.reshape(1, -1), None, 40, 40 if com else 0, "%s m%d o%d" % (system, method, out_type))


def test_f32_dopri5_alternative_of_the_bouncing_ball_example():
    """examples/bouncing_ball.rs:35 literally, State = Vector3<f32>:
    Dopri5::new(system, 0., 10.0, 0.01, y0, 1.0e-2, 1.0e-6) -- and Q3: `new` takes uround = f32::EPSILON,
    `from_param` f64::EPSILON as f32, two different steppers in f32"""
    cfg = ob.config_new_f32(ob.DOPRI5, 0.0, 10.0, 0.01, 1e-2, 1e-6)
    g = check_adaptive("bouncing_ball", cfg, np.array([[10.0], [0.0], [0.0]]), [9.81], 1100, 0, "example")
    n = int(g.out_count[0])
    below = int(np.argmax(g.out_y[0, :n, 0] <= 0.0))  # the example's `y_out.iter().find(|y| y[0] <= 0.)`
    assert g.status[0] == ob.OK and abs(g.out_x[0, below] - 1.43) < 0.011
    y0, p = W.lorenz_ensemble(64, offset=1)
    y0 = y0 * 1e4  # |x| stays 0 at the first test, but states this large make uround * |x| matter later on
    a = ob.config_new_f32(ob.DOPRI5, 0.0, 1e6, 1e5, 1e-3, 1e-3, out_type=ob.OUT_SPARSE)
    b = ob.config_from_param_f32(ob.DOPRI5, 0.0, 1e6, 1e5, 1e-3, 1e-3, 0.9, 0.04, 0.2, 10.0, 1e6, 0.0, 100000, 1000,
                                 ob.OUT_SPARSE)
    ga = check_adaptive("lorenz", a, y0, p, 0, 0, "Q3 new")
    gb = check_adaptive("lorenz", b, y0, p, 0, 0, "Q3 from_param")
    assert a.uround != b.uround and ga.n == gb.n


@pytest.mark.parametrize("seed", range(60))
def test_f32_adaptive_random_configurations_bitwise(seed):
    """the from_param space of the f32 adaptive steppers: both directions, explicit first steps, small n_max /
    n_stiff, capacities that overflow, non-finite states, every status code"""
    rng = np.random.default_rng(9000 + seed)
    system = str(rng.choice(["lorenz", "robertson", "bouncing_ball", "test_rk4_1", "test_rk4_2"]))
    method = int(rng.choice([ob.DOPRI5, ob.DOP853]))
    out_type = int(rng.choice([ob.OUT_DENSE, ob.OUT_SPARSE, ob.OUT_CONTINUOUS]))
    n = int(rng.choice([1, 33, 130]))
    if system == "lorenz":
        y0, p = W.lorenz_ensemble(n, offset=seed)
        span = float(rng.choice([0.5, 2.0]))
    elif system == "robertson":
        y0, p = W.robertson_sweep(n, offset=seed)
        span = float(rng.choice([0.01, 0.3]))
    elif system == "bouncing_ball":
        y0, p = W.bouncing_ball_ensemble(n, offset=seed)
        span = 6.0
    else:
        y0, p = rng.uniform(-2, 2, (1, n)), None
        span = 1.0
    x0 = float(rng.choice([0.0, 0.0, 0.3 * span]))
    x1 = x0 + span if rng.random() < 0.85 else x0 - span
    rtol, atol = float(10.0 ** rng.integers(-6, -1)), float(10.0 ** rng.integers(-7, -2))
    dx = span / float(rng.choice([7, 40, 333]))
    beta = float(rng.choice([0.0, 0.04, 0.08, 0.1]))
    h0 = float(rng.choice([0.0, 0.0, 1e-3 * span, -1e-3 * span]))
    n_max = int(rng.choice([100000, 100000, 25, 3]))
    n_stiff = int(rng.choice([1000, 1000, 2, 1, 0]))
    cfg = ob.config_from_param_f32(method, x0, x1, dx, rtol, atol, float(rng.choice([0.9, 0.8])), beta,
                                   float(rng.choice([0.2, 0.333])), float(rng.choice([10.0, 6.0, 2.0])),
                                   float(rng.choice([abs(x1 - x0), 0.1 * span])), h0, n_max, n_stiff, out_type)
    if rng.random() < 0.15:
        y0 = y0.copy()
        y0[0, ::4] = rng.choice([np.inf, np.nan, 3e38, -3e38, 1e-45, -0.0])
    cap = int(rng.choice([0, 5, 60, 400]))
    com = int(rng.choice([0, 7, 400])) if (out_type == ob.OUT_CONTINUOUS and method == ob.DOPRI5) else 0
    check_adaptive(system, cfg, y0, p, cap, com, "f32 fuzz %d" % seed)


def test_f32_stepper_classes_mirror_the_crate():
    """the host mirror's Rk4 / Dopri5 / Dop853 with dtype=np.float32 (Rk4::<f32>, Dopri5::<f32>: bouncing_ball.rs:15,
    35-36) return what the batched f32 entry returns for the same trajectory, x_out()/y_out() in f32"""
    f = ob.System.builtin("bouncing_ball", [9.81])
    r = ob.Rk4.new(f, 0.0, [10.0, 0.0, 0.0], 10.0, 0.01, dtype=np.float32)
    st = r.integrate()
    g = ob.rk4_f32_integrate_batch(f, 0.0, 10.0, 0.01, np.array([[10.0], [0.0], [0.0]]), out_cap=1100)
    n = int(g.out_count[0])
    assert st.accepted_steps == g.accepted[0] == 143 and len(r.x_out()) == n
    assert same_f32(np.array(r.y_out()), g.out_y[0, :n]).all() and r.y_out()[0].dtype == np.float32
    d = ob.Dopri5.new(f, 0.0, 10.0, 0.01, [10.0, 0.0, 0.0], 1e-2, 1e-6, dtype=np.float32)
    st = d.integrate()
    cfg = ob.config_new_f32(ob.DOPRI5, 0.0, 10.0, 0.01, 1e-2, 1e-6)
    g = ob.integrate_batch_f32(f, cfg, np.array([[10.0], [0.0], [0.0]]), out_cap=1100)
    n = int(g.out_count[0])
    assert (st.num_eval, st.accepted_steps, st.rejected_steps) == (g.num_eval[0], g.accepted[0], g.rejected[0])
    assert same_f32(np.array(d.y_out()), g.out_y[0, :n]).all() and same_f32(np.array(d.x_out()), g.out_x[0, :n]).all()
    com = ob.ContinuousOutputModel()
    c = ob.Dopri5.new(ob.System.builtin("lorenz", [10.0, 8.0 / 3.0, 28.0]), 0.0, 1.0, 0.1, [1.0, 1.0, 1.0], 1e-4, 1e-4,
                      dtype=np.float32)
    st = c.integrate_with_continuous_output_model(com)
    assert len(com.intervals) == st.accepted_steps and com.intervals[0].coefficients[0].dtype == np.float32
    y = com.evaluate(0.5)
    assert y is not None and np.all(np.isfinite(y)) and com.evaluate(1.5) is None
    e = ob.Dop853.from_param(ob.System.builtin("lorenz", [10.0, 8.0 / 3.0, 28.0]), 0.0, 1.0, 0.1, [1.0, 1.0, 1.0], 1e-4, 1e-4,
                             0.9, 0.0, 0.333, 6.0, 1.0, 0.0, 100000, 1000, ob.OutputType.Sparse, dtype=np.float32)
    st = e.integrate()
    assert st.accepted_steps > 3 and abs(e.x - 1.0) < 1e-6 and e.x_out()[-1] == 0.0  # Q6: Dop853 Sparse pushes x0
    with pytest.raises(ValueError):
        e.integrate_with_continuous_output_model(ob.ContinuousOutputModel())
