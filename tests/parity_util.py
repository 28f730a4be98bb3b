"""Shared helpers of the GPU parity tests: run one configuration through the CUDA library
(C ABI via ctypes) and through the CPU oracle on the same inputs, and compare bit for bit."""
import numpy as np

import ode_solvers_b200 as ob
import oracle
from ode_solvers_b200 import workloads as W


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def same_f64(a, b):
    """bitwise equality, NaN == NaN (payload / sign of NaN is not compared)"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))


def assert_same_outputs(g, o, label, check_samples=True):
    """g: CUDA HostOutputs, o: oracle HostOutputs -- identical Stats, status and bits."""
    for k in ("status", "accepted", "rejected", "num_eval", "n_step", "out_count", "com_count"):
        ga, oa = getattr(g, k), getattr(o, k)
        bad = np.nonzero(ga != oa)[0]
        assert bad.size == 0, "%s: %s differs for %d trajectories, first %d: gpu %s oracle %s" % (
            label, k, bad.size, bad[0], ga[bad[0]], oa[bad[0]])
    for k in ("x_final", "h_next", "com_lower"):
        ok = same_f64(getattr(g, k), getattr(o, k))
        assert ok.all(), "%s: %s differs for %d trajectories" % (label, k, (~ok).sum())
    ok = same_f64(g.y_final, o.y_final)
    if not ok.all():
        t = np.nonzero(~ok.all(axis=0))[0]
        raise AssertionError("%s: y_final differs for %d trajectories, first %d: gpu %r oracle %r" % (
            label, t.size, t[0], g.y_final[:, t[0]], o.y_final[:, t[0]]))
    if check_samples and g.out_cap:
        cnt = np.minimum(g.out_count, g.out_cap)
        for i in range(g.n):
            c = int(cnt[i])
            assert same_f64(g.out_x[i, :c], o.out_x[i, :c]).all(), "%s: out_x differs, trajectory %d" % (label, i)
            assert same_f64(g.out_y[i, :c], o.out_y[i, :c]).all(), "%s: out_y differs, trajectory %d" % (label, i)
    if check_samples and g.com_cap:
        cnt = np.minimum(g.com_count, g.com_cap)
        for i in range(g.n):
            c = int(cnt[i])
            assert same_f64(g.com_break[i, :c], o.com_break[i, :c]).all(), label + ": com_break differs"
            assert same_f64(g.com_step[i, :c], o.com_step[i, :c]).all(), label + ": com_step differs"
            assert same_f64(g.com_coef[i, :c], o.com_coef[i, :c]).all(), label + ": com_coef differs"


def run_both(system, cfg, y0, params, out_cap=0, com_cap=0, ctx=None, n_gpus=1):
    f = ob.System.builtin(system)
    g = ob.integrate_batch(f, cfg, y0, params, out_cap=out_cap, com_cap=com_cap, ctx=ctx, n_gpus=n_gpus)
    o = oracle.integrate_batch(system, cfg, y0, params, out_cap=out_cap, com_cap=com_cap)
    return g, o


ENSEMBLES = {
    "lorenz": W.lorenz_ensemble,
    "kepler": W.kepler_sweep,
    "cr3bp": W.cr3bp_ensemble,
    "robertson": W.robertson_sweep,
    "bouncing_ball": W.bouncing_ball_ensemble,
    "threebody18": W.threebody18_ensemble,
}
