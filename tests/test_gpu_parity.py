"""GPU parity tests proper: the CUDA step-loop kernels, called through the C ABI, against the
CPU oracle on the same seeded inputs. Bar: BIT-EXACT state, step sizes, Stats, status and
every recorded sample (which implies the north star's 1e-13 / 1e-10 tolerances and
"identical accepted and rejected step counts")."""
import numpy as np
import pytest

import ode_solvers_b200 as ob
import oracle
from ode_solvers_b200 import workloads as W
from parity_util import ENSEMBLES, assert_same_outputs, run_both

pytestmark = pytest.mark.gpu

N = 384  # 3 CTAs of 128 lanes; not a multiple of the queue's refill granularity on purpose below


def cfg_for(system, method, out_type):
    if system == "lorenz":
        x_end, dx, tol = 10.0, 0.05, 1e-6
    elif system == "kepler":
        x_end, dx, tol = 2.0 * W.kepler_period(), 600.0, 1e-10
    elif system == "cr3bp":
        x_end, dx, tol = 6.0, 0.25, 1e-12
    elif system == "robertson":
        x_end, dx, tol = 0.3, 0.01, None
    elif system == "bouncing_ball":
        x_end, dx, tol = 10.0, 0.01, None
    elif system == "threebody18":
        x_end, dx, tol = 2.0, 0.1, 1e-9
    if method == ob.RK4:
        step = {"lorenz": 1e-3, "kepler": 5.0, "cr3bp": 1e-3, "robertson": 1e-5, "bouncing_ball": 0.01,
                "threebody18": 1e-3}[system]
        return ob.config_rk4(0.0, x_end if system != "lorenz" else 2.0, step)
    if tol is None:
        c = ob.config_new(method, 0.0, x_end, dx, 1e-2, 1e-6, out_type=out_type)  # chemical_reaction.rs:18
    else:
        c = ob.config_new(method, 0.0, x_end, dx, tol, tol, out_type=out_type)
    return c


CASES = []
for sysname in ("lorenz", "kepler", "cr3bp", "robertson", "bouncing_ball", "threebody18"):
    CASES.append((sysname, ob.RK4, ob.OUT_SPARSE))
    for m in (ob.DOPRI5, ob.DOP853):
        for ot in (ob.OUT_SPARSE, ob.OUT_DENSE, ob.OUT_CONTINUOUS):
            CASES.append((sysname, m, ot))
    CASES.append((sysname, ob.DOP853, ob.OUT_CONTINUOUS8))  # extension: Dop853 -> ContinuousOutputModel


@pytest.mark.parametrize("system,method,out_type", CASES)
def test_bitwise_parity(system, method, out_type):
    n = N if system != "threebody18" else 128
    y0, p = ENSEMBLES[system](n, offset=1000)
    cfg = cfg_for(system, method, out_type)
    cap = 0
    if method == ob.RK4:
        cap = 0 if system in ("lorenz", "cr3bp", "robertson") else 2100
    elif out_type == ob.OUT_DENSE:
        cap = int((cfg.x_end - cfg.x0) / cfg.dx) + 3
    elif system in ("kepler", "bouncing_ball", "robertson"):
        cap = 1200
    com_cap = 1200 if (out_type == ob.OUT_CONTINUOUS and method == ob.DOPRI5 and system != "lorenz") else 0
    if out_type == ob.OUT_CONTINUOUS8:
        cap, com_cap = 700, 700
    g, o = run_both(system, cfg, y0, p, out_cap=cap, com_cap=com_cap)
    assert (o.status != -1).all()
    assert_same_outputs(g, o, "%s m%d o%d" % (system, method, out_type))
    assert int(g.accepted.sum()) > 0


def test_static_schedule_matches_work_queue():
    """the persistent work queue only changes WHICH lane integrates a trajectory, never a bit"""
    y0, p = W.kepler_sweep(1000, offset=7)
    cfg = ob.config_new(ob.DOP853, 0.0, W.kepler_period(), 60.0, 1e-10, 1e-10, out_type=ob.OUT_SPARSE)
    ctx = ob.Context(0)
    f = ob.System.builtin("kepler")
    a = ob.integrate_batch(f, cfg, y0, p, ctx=ctx)
    ctx.set_work_queue(False)
    b = ob.integrate_batch(f, cfg, y0, p, ctx=ctx)
    ctx.close()
    assert_same_outputs(a, b, "queue vs static")


def test_per_trajectory_params_and_odd_sizes():
    for n in (1, 31, 33, 129, 1000):
        y0, k = W.robertson_sweep(n, offset=5)
        cfg = ob.config_new(ob.DOPRI5, 0.0, 0.3, 0.3, 1e-2, 1e-6, out_type=ob.OUT_SPARSE)
        g, o = run_both("robertson", cfg, y0, k)
        assert_same_outputs(g, o, "robertson n=%d" % n)


@pytest.mark.parametrize("n", [1, 31, 255, 257, 600])
def test_lockstep_ctas_with_odd_sizes_and_both_schedules(n):
    """Dop853 with n >= 6 runs one 256-thread lockstep CTA per SM with batched refills: partial CTAs,
    a single trajectory, more trajectories than lanes per CTA, with and without the work queue"""
    y0, p = W.kepler_sweep(n, offset=40 + n)
    cfg = ob.config_new(ob.DOP853, 0.0, 0.7 * W.kepler_period(), 60.0, 1e-9, 1e-9, out_type=ob.OUT_SPARSE)
    o = oracle.integrate_batch("kepler", cfg, y0, p, out_cap=300)
    ctx = ob.Context(0)
    for queue in (True, False):
        ctx.set_work_queue(queue)
        g = ob.integrate_batch(ob.System.builtin("kepler"), cfg, y0, p, out_cap=300, ctx=ctx)
        assert_same_outputs(g, o, "kepler lockstep n=%d queue=%d" % (n, queue))
    ctx.close()


def test_work_queue_refills_lockstep_ctas_many_times():
    """far more trajectories than resident lanes (148 SMs x 256), very different lengths: every lane
    refills several times, in batches"""
    n = 148 * 256 * 3 + 77
    y0, p = W.kepler_sweep(n, offset=5)
    cfg = ob.config_new(ob.DOP853, 0.0, 0.25 * W.kepler_period(), 60.0, 1e-7, 1e-7, out_type=ob.OUT_SPARSE)
    g = ob.integrate_batch(ob.System.builtin("kepler"), cfg, y0, p)
    assert (g.status == ob.OK).all() and (g.x_final == cfg.x_end).all()
    sel = np.concatenate([np.arange(0, 300), np.arange(n - 300, n), np.arange(0, n, 997)])
    o = oracle.integrate_batch("kepler", cfg, y0[:, sel], p)
    for k in ("accepted", "rejected", "num_eval", "n_step"):
        assert (getattr(g, k)[sel] == getattr(o, k)).all(), k
    assert (g.y_final[:, sel].view(np.uint64) == o.y_final.view(np.uint64)).all()
    assert (g.h_next[sel].view(np.uint64) == o.h_next.view(np.uint64)).all()


@pytest.mark.parametrize("cap", [2001, 2002, 2003, 2004, 1000])
def test_staged_results_through_tma_bulk_copies(cap):
    """output-heavy calls record through the shared-memory rings and per-lane TMA bulk copies: row
    capacities of every residue mod 4 (padded device rows, tail groups written by the lane itself) and a
    capacity that overflows (every push still counted), every sample bitwise"""
    n = 1500
    y0, p = W.lorenz_ensemble(n, offset=17)
    for cfg in (ob.config_rk4(0.0, 2.0, 1e-3),                                  # 2001 samples, one per step
                ob.config_new(ob.DOPRI5, 0.0, 20.0, 0.01, 1e-6, 1e-6)):        # Dense, 2001 samples, 0-3 per step
        g, o = run_both("lorenz", cfg, y0, p, out_cap=cap)
        assert_same_outputs(g, o, "staged m%d cap %d" % (cfg.method, cap))
        assert (g.out_count == 2001).all()
        assert ((g.status == ob.OUTPUT_CAPACITY_EXCEEDED) == (cap < 2001)).all()


@pytest.mark.parametrize("com_cap", [901, 902, 399])
def test_staged_continuous_output_model_intervals(com_cap):
    """Dopri5 Continuous: samples and add_interval records (pairs of intervals per bulk copy) with odd /
    even / overflowing capacities, ragged interval counts per trajectory"""
    n = 1200
    y0, p = W.lorenz_ensemble(n, offset=23)
    cfg = ob.config_new(ob.DOPRI5, 0.0, 20.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_CONTINUOUS)
    g, o = run_both("lorenz", cfg, y0, p, out_cap=com_cap + 1, com_cap=com_cap)
    assert_same_outputs(g, o, "staged com cap %d" % com_cap)
    assert g.com_count.min() > 300 and len(set(g.com_count.tolist())) > 50
    assert (g.status == ob.OUTPUT_CAPACITY_EXCEEDED).any()  # some trajectories need more than 900 intervals


def test_empty_batch():
    f = ob.System.builtin("lorenz")
    cfg = ob.config_new(ob.DOPRI5, 0.0, 1.0, 0.1, 1e-6, 1e-6)
    g = ob.integrate_batch(f, cfg, np.zeros((3, 0)), W.LORENZ_PARAMS)
    assert g.n == 0
