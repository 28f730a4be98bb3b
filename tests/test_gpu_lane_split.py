"""Dopri5 and Dop853 for large states: one trajectory per group of three lanes (csrc/ode_split.cuh), against the
oracle and against the one-trajectory-per-thread kernels, bit for bit: every output mode, both
schedules, odd batch sizes (partial groups / warps / CTAs), the stiffness test (a whole-warp phase),
every early exit, per-trajectory parameters, non-finite and badly scaled inputs (exact redo path)."""
import numpy as np
import pytest

import ode_solvers_b200 as ob
import oracle
from ode_solvers_b200 import workloads as W
from parity_util import assert_same_outputs

pytestmark = pytest.mark.gpu
F = None


def run(cfg, y0, p, out_cap=0, com_cap=0, split=True, queue=True):
    ctx = ob.Context(0)
    ctx.set_lane_split(split)
    ctx.set_work_queue(queue)
    g = ob.integrate_batch(ob.System.builtin("threebody18"), cfg, y0, p, out_cap=out_cap, com_cap=com_cap, ctx=ctx)
    ctx.close()
    return g


@pytest.mark.parametrize("method,out_type,cap,com_cap",
                         [(ob.DOP853, ob.OUT_SPARSE, 0, 0), (ob.DOP853, ob.OUT_SPARSE, 200, 0), (ob.DOP853, ob.OUT_DENSE, 60, 0),
                          (ob.DOP853, ob.OUT_CONTINUOUS, 200, 0), (ob.DOP853, ob.OUT_CONTINUOUS8, 200, 200),
                          (ob.DOP853, ob.OUT_CONTINUOUS8, 20, 20),
                          (ob.DOPRI5, ob.OUT_SPARSE, 0, 0), (ob.DOPRI5, ob.OUT_SPARSE, 700, 0), (ob.DOPRI5, ob.OUT_DENSE, 60, 0),
                          (ob.DOPRI5, ob.OUT_CONTINUOUS, 700, 700), (ob.DOPRI5, ob.OUT_CONTINUOUS, 700, 0),
                          (ob.DOPRI5, ob.OUT_CONTINUOUS, 20, 20)])
@pytest.mark.parametrize("n", [1, 9, 10, 11, 333, 4000])
def test_split_matches_oracle_and_the_per_thread_kernel(method, out_type, cap, com_cap, n):
    y0, p = W.threebody18_ensemble(n, offset=100 + n)
    cfg = ob.config_new(method, 0.0, 2.5, 0.05, 1e-9, 1e-9, out_type=out_type)
    o = oracle.integrate_batch("threebody18", cfg, y0, p, out_cap=cap, com_cap=com_cap)
    for queue in (True, False):
        g = run(cfg, y0, p, cap, com_cap, split=True, queue=queue)
        assert_same_outputs(g, o, "split m%d o%d n%d queue%d" % (method, out_type, n, queue))
    if n <= 333:
        g1 = run(cfg, y0, p, cap, com_cap, split=False)
        assert_same_outputs(g1, o, "per-thread m%d o%d n%d" % (method, out_type, n))
    assert (o.status == (ob.OK if cap != 20 else ob.OUTPUT_CAPACITY_EXCEEDED)).all()


@pytest.mark.parametrize("method", [ob.DOPRI5, ob.DOP853])
def test_split_stiffness_phase_and_every_exit(method):
    n = 700
    cont = ob.OUT_CONTINUOUS8 if method == ob.DOP853 else ob.OUT_CONTINUOUS
    y0, p = W.threebody18_ensemble(n, offset=9)
    # n_stiff = 3: the stiffness quotient (nalgebra-order dot products over the gathered state) is
    # evaluated every third accepted step; n_max = 40 cuts some trajectories short (MaxNumStepReached)
    cfg = ob.config_from_param(method, 0.0, 4.0, 0.1, 1e-10, 1e-12, 0.9, 0.0, 0.333, 6.0, 4.0, 0.0, 40, 3, ob.OUT_SPARSE)
    o = oracle.integrate_batch("threebody18", cfg, y0, p, out_cap=64)
    g = run(cfg, y0, p, 64)
    assert_same_outputs(g, o, "split stiff/nmax")
    assert (o.status == ob.MAX_NUM_STEP_REACHED).any()
    # close encounters: two bodies almost on top of each other -> huge accelerations, step-size underflow /
    # stiffness detection, and operands far outside the fast paths' window (exact redo)
    y0b = y0.copy()
    y0b[3:6, ::2] = y0b[0:3, ::2] + 1e-9
    y0b[6:9, 1::7] = y0b[0:3, 1::7] * (1 + 1e-15)
    cfg = ob.config_from_param(method, 0.0, 1.0, 0.1, 1e-9, 1e-9, 0.9, 0.0, 0.333, 6.0, 1.0, 0.0, 3000, 2, ob.OUT_DENSE)
    o = oracle.integrate_batch("threebody18", cfg, y0b, p, out_cap=16)
    g = run(cfg, y0b, p, 16)
    assert_same_outputs(g, o, "split close encounters")
    assert len(set(o.status.tolist())) >= 2, set(o.status.tolist())
    # non-finite and badly scaled states
    y0c = y0.copy()
    y0c[0, 0] = np.inf
    y0c[10, 1] = np.nan
    y0c[:, 2] *= 1e-200
    y0c[:, 3] *= 1e150
    y0c[0:9, 4] = 0.0  # all three bodies at the origin: 0/0
    y0c[17, 5] = -0.0
    for ot in (ob.OUT_SPARSE, cont, ob.OUT_DENSE):
        cfg = ob.config_new(method, 0.0, 1.0, 0.1, 1e-9, 1e-9, out_type=ot)
        cfg.n_max = 300
        o = oracle.integrate_batch("threebody18", cfg, y0c[:, :64], p, out_cap=40, com_cap=40)
        g = run(cfg, y0c[:, :64], p, 40, 40)
        assert_same_outputs(g, o, "split non-finite o%d" % ot)


@pytest.mark.parametrize("method", [ob.DOPRI5, ob.DOP853])
def test_split_per_trajectory_masses_backward_and_explicit_first_step(method):
    n = 500
    y0, _ = W.threebody18_ensemble(n, offset=77)
    gm = np.stack([W.uniform(np.arange(n, dtype=np.uint64), 90 + c, 0.5, 1.5) for c in range(3)])
    for x0, x1, h0 in ((0.0, 1.5, 0.0), (1.0, -0.5, 0.0), (0.0, 1.0, 1e-3), (0.0, 1.0, -1e-3)):
        cfg = ob.config_from_param(method, x0, x1, 0.1, 1e-8, 1e-10, 0.9, 0.1, 0.333, 6.0, 0.2, h0, 100000, 1000,
                                   ob.OUT_SPARSE)
        o = oracle.integrate_batch("threebody18", cfg, y0, gm, out_cap=400)
        g = run(cfg, y0, gm, 400)
        assert_same_outputs(g, o, "split masses %g->%g h0=%g" % (x0, x1, h0))


@pytest.mark.parametrize("method", [ob.DOPRI5, ob.DOP853])
def test_split_refills_many_times(method):
    n = 148 * 120 * 3 + 31  # ~3 generations of the 120 trajectories per SM of the Sparse kernel
    y0, p = W.threebody18_ensemble(n, offset=5)
    cfg = ob.config_new(method, 0.0, 1.0, 0.1, 1e-7, 1e-7, out_type=ob.OUT_SPARSE)
    g = run(cfg, y0, p)
    assert (g.status == ob.OK).all() and (g.x_final == cfg.x_end).all()
    sel = np.concatenate([np.arange(0, 200), np.arange(n - 200, n), np.arange(0, n, 997)])
    o = oracle.integrate_batch("threebody18", cfg, y0[:, sel], p)
    for k in ("accepted", "rejected", "num_eval", "n_step"):
        assert (getattr(g, k)[sel] == getattr(o, k)).all(), k
    assert (g.y_final[:, sel].view(np.uint64) == o.y_final.view(np.uint64)).all()
    assert (g.h_next[sel].view(np.uint64) == o.h_next.view(np.uint64)).all()
