"""Edge cases of the step loop, GPU vs oracle bit for bit: backward integration, explicit initial
step, from_param variants, inputs on which the reference panics / never returns (Q8), output
capacity overflow, solout-driven early stops with ragged lengths, ragged dense sample counts."""
import numpy as np
import pytest

import ode_solvers_b200 as ob
from ode_solvers_b200 import workloads as W
from parity_util import assert_same_outputs, run_both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("method", [ob.DOPRI5, ob.DOP853])
def test_backward_integration_sparse(method):
    y0, p = W.kepler_sweep(200, offset=21)  # time-reversible (backward Lorenz blows up: volume expands)
    T = W.kepler_period()
    cfg = ob.config_new(method, 2.0 * T, 0.5 * T, 60.0, 1e-9, 1e-9, out_type=ob.OUT_SPARSE)  # x_end < x: posneg = -1
    g, o = run_both("kepler", cfg, y0, p, out_cap=600)
    assert_same_outputs(g, o, "backward m%d" % method)
    # h_old = posneg * h_new (dopri5.rs:433): for backward integration the stored value is |h_new|
    assert (g.status == ob.OK).all() and (g.x_final == 0.5 * T).all() and (g.h_next > 0).all()
    # and the blow-up case: both sides must agree on how it fails
    y0, p = W.lorenz_ensemble(64, offset=21)
    cfg = ob.config_new(method, 3.0, 0.5, 0.1, 1e-7, 1e-7, out_type=ob.OUT_SPARSE)
    g, o = run_both("lorenz", cfg, y0, p)
    assert_same_outputs(g, o, "backward lorenz m%d" % method)
    assert (g.status != ob.OK).all()


def test_backward_rk4_runs_zero_steps_like_the_reference():
    y0, p = W.lorenz_ensemble(64)
    g, o = run_both("lorenz", ob.config_rk4(1.0, 0.0, 0.01), y0, p, out_cap=4)  # ceil(-100) -> to_usize panics
    assert_same_outputs(g, o, "rk4 backward")
    assert (g.status == ob.UNSUPPORTED_DOMAIN).all()


@pytest.mark.parametrize("method", [ob.DOPRI5, ob.DOP853])
def test_from_param_variants(method):
    y0, p = W.lorenz_ensemble(160, offset=5)
    beta = 0.08 if method == ob.DOPRI5 else 0.1
    cfg = ob.config_from_param(method, 0.0, 4.0, 0.2, 1e-5, 1e-8, 0.8, beta, 0.3, 5.0, 0.05, 1e-3, 20000, 7,
                               ob.OUT_DENSE)
    g, o = run_both("lorenz", cfg, y0, p, out_cap=30)
    assert_same_outputs(g, o, "from_param m%d" % method)
    # beta = 0 for Dopri5 (pow(x, -0.0) path) and a step limit that bites
    cfg = ob.config_from_param(method, 0.0, 4.0, 0.2, 1e-5, 1e-8, 0.9, 0.0, 0.2, 10.0, 4.0, 0.0, 60, 1000,
                               ob.OUT_SPARSE)
    g, o = run_both("lorenz", cfg, y0, p)
    assert_same_outputs(g, o, "from_param beta=0 m%d" % method)
    assert (g.status == ob.MAX_NUM_STEP_REACHED).any()


@pytest.mark.parametrize("method", [ob.DOPRI5, ob.DOP853])
def test_dense_outside_the_reference_domain_is_flagged(method):
    y0, p = W.lorenz_ensemble(96)
    for x0, x_end, dx in ((-1.0, 1.0, 0.1), (0.0, 1.0, 0.0), (0.0, 1.0, -0.1), (2.0, 1.0, 0.1)):
        cfg = ob.config_new(method, x0, x_end, dx, 1e-6, 1e-6)
        g, o = run_both("lorenz", cfg, y0, p, out_cap=40)
        assert_same_outputs(g, o, "dense domain %r m%d" % ((x0, x_end, dx), method))


def test_output_capacity_exceeded_counts_every_push():
    y0, p = W.lorenz_ensemble(100)
    cfg = ob.config_new(ob.DOPRI5, 0.0, 2.0, 0.01, 1e-6, 1e-6)
    g, o = run_both("lorenz", cfg, y0, p, out_cap=50)  # the reference pushes 201 samples
    assert_same_outputs(g, o, "cap overflow")
    assert (g.out_count == 201).all() and (g.status == ob.OUTPUT_CAPACITY_EXCEEDED).all()
    cfg.out_type = ob.OUT_CONTINUOUS
    g, o = run_both("lorenz", cfg, y0, p, out_cap=10, com_cap=10)
    assert_same_outputs(g, o, "com cap overflow")
    assert (g.com_count > 10).all() and (g.status == ob.OUTPUT_CAPACITY_EXCEEDED).all()


@pytest.mark.parametrize("method,out_type", [(ob.RK4, ob.OUT_SPARSE), (ob.DOPRI5, ob.OUT_DENSE),
                                             (ob.DOPRI5, ob.OUT_SPARSE), (ob.DOP853, ob.OUT_DENSE),
                                             (ob.DOP853, ob.OUT_SPARSE)])
def test_solout_early_stop_ragged(method, out_type):
    """bouncing ball: every trajectory stops at its own time (examples/bouncing_ball.rs:83-85)"""
    n = 1000
    y0, p = W.bouncing_ball_ensemble(n, offset=3)
    if method == ob.RK4:
        cfg = ob.config_rk4(0.0, 10.0, 0.01)
    else:
        cfg = ob.config_new(method, 0.0, 10.0, 0.01, 1e-2, 1e-6, out_type=out_type)  # bouncing_ball.rs:35
    g, o = run_both("bouncing_ball", cfg, y0, p, out_cap=1010)
    assert_same_outputs(g, o, "ball m%d o%d" % (method, out_type))
    assert (g.y_final[0] < 0).all() and (g.status == ob.OK).all()  # every ball ends below the floor
    if method == ob.RK4:  # fixed step: every trajectory stops at its own step count
        assert len(np.unique(g.out_count)) > 100 and (g.x_final < 10.0).all()
    if method == ob.DOP853:  # the error estimate is exactly zero, steps grow 6x: a handful of steps each
        assert (g.x_final < 10.0).all() and (g.accepted < 20).all()


def test_single_trajectory_and_huge_tolerance_edge():
    y0 = np.array([[1.0], [1.0], [1.0]])
    for tol in (1e-14, 1e-1, 10.0):
        cfg = ob.config_new(ob.DOP853, 0.0, 1.0, 0.25, tol, tol)
        g, o = run_both("lorenz", cfg, y0, W.LORENZ_PARAMS, out_cap=8)
        assert_same_outputs(g, o, "tol %g" % tol)
