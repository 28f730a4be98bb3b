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
