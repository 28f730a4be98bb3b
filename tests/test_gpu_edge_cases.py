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


@pytest.mark.parametrize("system,method,out_type", [("lorenz", ob.DOPRI5, ob.OUT_CONTINUOUS),
                                                    ("robertson", ob.DOPRI5, ob.OUT_CONTINUOUS),
                                                    ("test_exp", ob.DOP853, ob.OUT_CONTINUOUS8),
                                                    ("kepler", ob.DOPRI5, ob.OUT_CONTINUOUS)])
def test_continuous_with_samples_but_without_model_capacity(system, method, out_type):
    """OutputType::Continuous with out_cap > 0 and com_cap = 0: the staged kernels keep their interval ring
    in the shared-memory layout whether or not the model is recorded (round-1 advisor finding: the launcher
    sized the dynamic shared memory without it)."""
    if system == "test_exp":
        y0, p = np.linspace(0.5, 1.5, 300).reshape(1, -1), None
        cfg = ob.config_new(method, 0.0, 1.0, 0.1, 1e-4, 1e-4, out_type=out_type)
    elif system == "kepler":
        y0, p = W.kepler_sweep(300)
        cfg = ob.config_new(method, 0.0, W.kepler_period(), 60.0, 1e-8, 1e-8, out_type=out_type)
    elif system == "robertson":
        y0, p = W.robertson_sweep(300)
        cfg = ob.config_new(method, 0.0, 0.3, 0.3, 1e-4, 1e-8, out_type=out_type)
    else:
        y0, p = W.lorenz_ensemble(300)
        cfg = ob.config_new(method, 0.0, 2.0, 0.1, 1e-6, 1e-6, out_type=out_type)
    g, o = run_both(system, cfg, y0, p, out_cap=400, com_cap=0)
    assert_same_outputs(g, o, "%s continuous, com_cap=0" % system)
    ok = g.status == ob.OK  # an IntegrationError leaves the loop before the last step's interval is added
    assert ok.any() and (g.com_count[ok] == g.accepted[ok]).all() and (g.out_count > 1).all()
    g2, o2 = run_both(system, cfg, y0, p, out_cap=400, com_cap=400)  # and the recorded model still agrees
    assert_same_outputs(g2, o2, "%s continuous, com_cap=400" % system)


@pytest.mark.parametrize("method", [ob.RK4, ob.DOPRI5, ob.DOP853])
def test_negative_zero_state_components(method):
    """Signed zeros (round-1 advisor finding on the skipped zero-coefficient terms, see the header note of
    ode_kernels.cuh): planar orbits with z = vz = -0.0 / +-0.0 mixtures, forward and backward, and a
    Lorenz / bouncing-ball ensemble with -0.0 components, all bit for bit including the sign of zero."""
    T = W.kepler_period()
    y0, p = W.kepler_sweep(256)
    signs = [(-0.0, -0.0), (-0.0, 0.0), (0.0, -0.0)]
    for i in range(y0.shape[1]):
        y0[2, i], y0[5, i] = signs[i % 3]
    for x0, x1 in ((0.0, 1.5 * T), (1.5 * T, 0.0)):
        cfg = ob.config_rk4(x0, x1, 30.0) if method == ob.RK4 else ob.config_new(
            method, x0, x1, T / 16, 1e-9, 1e-9, out_type=ob.OUT_DENSE if x0 == 0.0 else ob.OUT_SPARSE)
        g, o = run_both("kepler", cfg, y0, p, out_cap=900)
        assert_same_outputs(g, o, "kepler -0.0 m%d %g->%g" % (method, x0, x1))
    y0, p = W.lorenz_ensemble(192)
    y0[0, ::2] = -0.0
    y0[1, ::3] = -0.0
    y0[2, ::5] = -0.0
    y0[:, 7] = -0.0  # the fixed point at the origin, every component -0.0
    cfg = ob.config_rk4(0.0, 1.0, 1e-2) if method == ob.RK4 else ob.config_new(method, 0.0, 2.0, 0.05, 1e-7, 1e-7)
    out_type_runs = [cfg.out_type] if method == ob.RK4 else [ob.OUT_DENSE, ob.OUT_SPARSE] + (
        [ob.OUT_CONTINUOUS] if method == ob.DOPRI5 else [ob.OUT_CONTINUOUS8])
    for ot in out_type_runs:
        cfg.out_type = ot
        g, o = run_both("lorenz", cfg, y0, p, out_cap=400, com_cap=0 if method == ob.RK4 else 400)
        assert_same_outputs(g, o, "lorenz -0.0 m%d o%d" % (method, ot))
    y0, p = W.bouncing_ball_ensemble(128)
    y0[2, :] = -0.0  # dy[2] = +0.0 (Q13): the state stays -0.0 + 0.0 * h
    cfg = ob.config_rk4(0.0, 2.0, 0.01) if method == ob.RK4 else ob.config_new(method, 0.0, 10.0, 0.01, 1e-2, 1e-6,
                                                                                 out_type=ob.OUT_SPARSE)
    g, o = run_both("bouncing_ball", cfg, y0, p, out_cap=300)
    assert_same_outputs(g, o, "ball -0.0 m%d" % method)


def test_two_launches_of_one_context_in_flight_on_different_streams():
    """ode_b200_integrate_batch_device on two caller streams without synchronising in between: every launch
    owns its work-queue head (round-1 advisor finding: they used to share one counter)."""
    import ctypes as C

    import torch
    from ode_solvers_b200 import _abi
    lib = _abi.load()
    ctx = ob.Context(0)
    dev = torch.device("cuda", 0)
    f = ob.System.builtin("lorenz")
    n = 40000
    cfgs = [ob.config_new(ob.DOPRI5, 0.0, 6.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_SPARSE),
            ob.config_new(ob.DOP853, 0.0, 4.0, 0.1, 1e-7, 1e-7, out_type=ob.OUT_SPARSE)]
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    runs = []
    for rep in range(3):  # several rounds, alternating streams, nothing synchronised in between
        for k, (cfg, st) in enumerate(zip(cfgs, streams)):
            y0, p = W.lorenz_ensemble(n, offset=1000 * rep + 17 * k)
            d_y0, d_p = torch.from_numpy(y0).to(dev), torch.from_numpy(np.asarray(p)).to(dev)
            bufs = {"y_final": torch.full((3, n), float("nan"), dtype=torch.float64, device=dev),
                    "accepted": torch.full((n,), -1, dtype=torch.int32, device=dev),
                    "status": torch.full((n,), -9, dtype=torch.int32, device=dev)}
            torch.cuda.synchronize()
            o = _abi.Outputs()
            o.y_final = C.cast(bufs["y_final"].data_ptr(), _abi.c_double_p)
            o.accepted = C.cast(bufs["accepted"].data_ptr(), _abi.c_u32_p)
            o.status = C.cast(bufs["status"].data_ptr(), _abi.c_i32_p)
            runs.append((cfg, y0, p, bufs, d_y0, d_p, o, st))
    for cfg, y0, p, bufs, d_y0, d_p, o, st in runs:
        rc = lib.ode_b200_integrate_batch_device(ctx.handle, f.sys_id, C.byref(cfg), n, d_y0.data_ptr(), d_p.data_ptr(),
                                                 0, C.byref(o), st.cuda_stream)
        assert rc == 0, _abi.last_error()
    torch.cuda.synchronize()
    import oracle
    for cfg, y0, p, bufs, *_ in runs:
        ref = oracle.integrate_batch("lorenz", cfg, y0, p)
        assert (bufs["status"].cpu().numpy() == 0).all(), "a trajectory was never integrated"
        assert np.array_equal(bufs["accepted"].cpu().numpy().astype(np.uint32), ref.accepted)
        assert np.array_equal(bufs["y_final"].cpu().numpy().view(np.uint64), ref.y_final.view(np.uint64))
    # NULL outputs / config are rejected instead of dereferenced
    assert lib.ode_b200_integrate_batch_device(ctx.handle, f.sys_id, C.byref(cfgs[0]), n, runs[0][4].data_ptr(),
                                               runs[0][5].data_ptr(), 0, None, None) == -1
    assert lib.ode_b200_integrate_batch_device(ctx.handle, f.sys_id, None, n, runs[0][4].data_ptr(),
                                               runs[0][5].data_ptr(), 0, C.byref(runs[0][6]), None) == -1


@pytest.mark.parametrize("system", ["lorenz", "kepler", "threebody18"])
def test_trajectories_that_end_in_the_pass_they_were_claimed_in(system):
    """More trajectories than resident lanes, each of which finishes at its first attempt (zero Rk4 steps;
    Dop853 with n_max = 0 -> MaxNumStepReached at once): every lane is idle right after a refill while the
    queue still holds work. (Round 2 found the scheduler's exit test `nobody active after a refill` wrong
    for exactly this case; it is now `nobody active and every lane has had a claim refused`.)"""
    n = 400000 if system != "threebody18" else 60000
    gen = {"lorenz": W.lorenz_ensemble, "kepler": W.kepler_sweep, "threebody18": W.threebody18_ensemble}[system]
    y0, p = gen(n, offset=5)
    f = ob.System.builtin(system)
    for cfg, want in ((ob.config_rk4(1.0, 1.0, 0.1), ob.OK),
                      (ob.config_from_param(ob.DOP853, 0.0, 1.0, 0.1, 1e-6, 1e-6, 0.9, 0.0, 0.333, 6.0, 1.0, 1e-3, 0, 1000,
                                            ob.OUT_SPARSE), ob.MAX_NUM_STEP_REACHED)):
        for queue in (True, False):
            ctx = ob.Context(0)
            ctx.set_work_queue(queue)
            g = ob.integrate_batch(f, cfg, y0, p, ctx=ctx)
            ctx.close()
            assert (g.status == want).all(), "%d trajectories were never integrated" % int((g.status == -1).sum())
            assert (g.y_final == y0).all() and (g.accepted == 0).all()


@pytest.mark.parametrize("n,cap,steps", [(32, 8, 7), (33, 7, 9), (31, 4, 3), (64, 5, 12), (4096 + 5, 13, 12), (4096, 16, 15),
                                         (2 * 148 * 640 + 77, 6, 5), (1000, 3, 0)])
def test_rk4_warp_level_result_stores(n, cap, steps, monkeypatch):
    """Rk4 with every step recorded: full warps of consecutive trajectories hand their samples to the TMA engine as
    one tensor store per group (push_warp_), partial warps through per-lane bulk copies -- capacities that are no
    multiple of the group, that cut the run short (the counts go on), several generations per lane, zero steps,
    both schedules, and the switch off: identical bits"""
    import oracle
    y0, p = W.lorenz_ensemble(n, offset=17)
    cfg = ob.config_rk4(0.0, steps * 1e-3, 1e-3)
    sel = np.arange(n) if n <= 5000 else np.concatenate([np.arange(0, 700), np.arange(n - 700, n), np.arange(0, n, 389)])
    o = oracle.integrate_batch("lorenz", cfg, y0[:, sel], p, out_cap=cap)
    for tensor in ("1", "0"):
        monkeypatch.setenv("ODE_B200_TENSOR_STORES", tensor)
        for queue in (True, False):
            ctx = ob.Context(0)
            ctx.set_work_queue(queue)
            g = ob.integrate_batch(ob.System.builtin("lorenz"), cfg, y0, p, out_cap=cap, ctx=ctx)
            ctx.close()
            assert (g.out_count == steps + 1).all() and (g.status == (ob.OK if steps + 1 <= cap else ob.OUTPUT_CAPACITY_EXCEEDED)).all()
            c = min(cap, steps + 1)
            assert (g.out_x[sel, :c].view(np.uint64) == o.out_x[:, :c].view(np.uint64)).all(), (tensor, queue)
            assert (g.out_y[sel, :c].view(np.uint64) == o.out_y[:, :c].view(np.uint64)).all(), (tensor, queue)
            assert (g.y_final[:, sel].view(np.uint64) == o.y_final.view(np.uint64)).all()
