#!/usr/bin/env python3
"""Kernel-time micro-harness for tuning experiments: runs named workloads through the C ABI's
host-buffer entry and prints the device time of the step-loop kernel (CUDA events inside the
library). ODE_B200_LIB=<path> picks an alternative build of the library.

usage: kbench.py [workload ...]   workloads: lorenz_dp5 lorenz_dp8 lorenz_rk4 kepler_dp8 kepler_dp5
                                             cr3bp_dp8_dense lorenz_dp5_dense ball_dp5_dense ...
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import ode_solvers_b200 as ob  # noqa: E402
from ode_solvers_b200 import workloads as W  # noqa: E402


def build(name, n):
    if name == "lorenz_dp5":
        y0, p = W.lorenz_ensemble(n)
        return "lorenz", ob.config_new(ob.DOPRI5, 0.0, 20.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_SPARSE), y0, p, 0, 262
    if name == "lorenz_dp8":
        y0, p = W.lorenz_ensemble(n)
        return "lorenz", ob.config_new(ob.DOP853, 0.0, 20.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_SPARSE), y0, p, 0, 616
    if name == "lorenz_rk4":
        y0, p = W.lorenz_ensemble(n)
        return "lorenz", ob.config_rk4(0.0, 20.0, 1e-3), y0, p, 0, 75
    if name == "lorenz_dp5_dense":
        y0, p = W.lorenz_ensemble(n)
        return "lorenz", ob.config_new(ob.DOPRI5, 0.0, 20.0, 0.1, 1e-6, 1e-6), y0, p, 203, 298
    if name == "lorenz_dp5_dense_hbm":  # dense-output-heavy: 2001 samples x 32 B per trajectory
        y0, p = W.lorenz_ensemble(n)
        return "lorenz", ob.config_new(ob.DOPRI5, 0.0, 20.0, 0.01, 1e-6, 1e-6), y0, p, 2003, 298
    if name == "lorenz_rk4_all_steps":  # every step recorded: 2001 x 32 B per trajectory
        y0, p = W.lorenz_ensemble(n)
        return "lorenz", ob.config_rk4(0.0, 2.0, 1e-3), y0, p, 2002, 75
    if name == "lorenz_dp5_com":        # ContinuousOutputModel: 136 B per accepted step
        y0, p = W.lorenz_ensemble(n)
        return ("lorenz", ob.config_new(ob.DOPRI5, 0.0, 20.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_CONTINUOUS), y0, p,
                -1000, 298)
    if name == "kepler_dp8":
        y0, p = W.kepler_sweep(n)
        return ("kepler", ob.config_new(ob.DOP853, 0.0, 5.0 * W.kepler_period(), 60.0, 1e-10, 1e-10,
                                        out_type=ob.OUT_SPARSE), y0, p, 0, 1120)
    if name == "kepler_dp5":
        y0, p = W.kepler_sweep(n)
        return ("kepler", ob.config_new(ob.DOPRI5, 0.0, 5.0 * W.kepler_period(), 60.0, 1e-10, 1e-10,
                                        out_type=ob.OUT_SPARSE), y0, p, 0, 487)
    if name == "kepler_dp5_dense":      # 2335 samples per trajectory (dx = 60 s over 5 periods)
        y0, p = W.kepler_sweep(n)
        return ("kepler", ob.config_new(ob.DOPRI5, 0.0, 5.0 * W.kepler_period(), 60.0, 1e-10, 1e-10), y0, p, 2400, 487)
    if name == "kepler_dp5_com":
        y0, p = W.kepler_sweep(n)
        return ("kepler", ob.config_new(ob.DOPRI5, 0.0, 5.0 * W.kepler_period(), 60.0, 1e-10, 1e-10,
                                        out_type=ob.OUT_CONTINUOUS), y0, p, -1400, 487)
    if name == "cr3bp_dp5":
        y0, p = W.cr3bp_ensemble(n)
        return ("cr3bp", ob.config_new(ob.DOPRI5, 0.0, 20.0, 0.5, 1e-10, 1e-10, out_type=ob.OUT_SPARSE), y0, p, 0, 700)
    if name == "cr3bp_dp8_dense":
        y0, p = W.cr3bp_ensemble(n)
        return "cr3bp", ob.config_new(ob.DOP853, 0.0, 20.0, 0.5, 1e-12, 1e-12), y0, p, 43, 1395
    if name == "cr3bp_dp8":
        y0, p = W.cr3bp_ensemble(n)
        return ("cr3bp", ob.config_new(ob.DOP853, 0.0, 20.0, 0.5, 1e-12, 1e-12, out_type=ob.OUT_SPARSE), y0, p, 0,
                1395)
    if name == "threebody18_dp8":
        y0, p = W.threebody18_ensemble(n)
        return ("threebody18", ob.config_new(ob.DOP853, 0.0, 5.0, 0.1, 1e-9, 1e-9, out_type=ob.OUT_SPARSE), y0, p, 0,
                3708)
    if name == "threebody18_dp8_com":   # config 3: Dop853 with dense output into a ContinuousOutputModel (extension)
        y0, p = W.threebody18_ensemble(n)
        return ("threebody18", ob.config_new(ob.DOP853, 0.0, 5.0, 0.1, 1e-9, 1e-9, out_type=ob.OUT_CONTINUOUS8), y0, p,
                -120, 3708 + 153 * 18 + 3 * 90)
    if name == "threebody18_dp5":       # 6 RHS of ~90 flops + 20 stage terms x 3 x 18 + error norm ~18 x 18
        y0, p = W.threebody18_ensemble(n)
        return ("threebody18", ob.config_new(ob.DOPRI5, 0.0, 5.0, 0.1, 1e-9, 1e-9, out_type=ob.OUT_SPARSE), y0, p, 0,
                1944)
    if name == "threebody18_dp5_com":   # every accepted step + its ContinuousOutputModel interval (5 x 18 + 2 doubles)
        y0, p = W.threebody18_ensemble(n)
        return ("threebody18", ob.config_new(ob.DOPRI5, 0.0, 5.0, 0.1, 1e-9, 1e-9, out_type=ob.OUT_CONTINUOUS), y0, p,
                -300, 1944 + 14 * 18)
    if name == "robertson_dp8":
        y0, p = W.robertson_sweep(n)
        return ("robertson", ob.config_new(ob.DOP853, 0.0, 0.3, 0.3, 1e-2, 1e-6, out_type=ob.OUT_SPARSE), y0, p, 0,
                638)
    if name == "ball_dp5_dense":
        y0, p = W.bouncing_ball_ensemble(n)
        return "bouncing_ball", ob.config_new(ob.DOPRI5, 0.0, 10.0, 0.01, 1e-2, 1e-6), y0, p, 1003, 250
    raise SystemExit("unknown workload " + name)


def f32_rows(n, reps, ctx):
    """the T = f32 steppers (csrc/ode_f32.cuh) on the Lorenz ensemble; tolerances f32 can reach"""
    y0, p = W.lorenz_ensemble(n)
    f = ob.System.builtin("lorenz")
    for name, cfg in (("lorenz_dp5_f32", ob.config_new_f32(ob.DOPRI5, 0.0, 20.0, 0.1, 1e-4, 1e-4, out_type=ob.OUT_SPARSE)),
                      ("lorenz_dp8_f32", ob.config_new_f32(ob.DOP853, 0.0, 20.0, 0.1, 1e-4, 1e-4, out_type=ob.OUT_SPARSE)),
                      ("lorenz_rk4_f32", ob.config_rk4_f32(0.0, 2.0, 1e-3))):
        best = 1e30
        for _ in range(reps):
            o = ob.integrate_batch_f32(f, cfg, y0.astype(np.float32), p.astype(np.float32), ctx=ctx)
            best = min(best, o.kernel_ms)
        acc = int(o.accepted.sum())
        print("%-18s n=%d kernel %.3f ms  %.3e acc-steps/s  acc/att=%d/%d bad=%d" % (
            name, n, best, acc / best * 1e3, acc, int(o.n_step.sum()), int((o.status != 0).sum())), flush=True)


def main():
    names = sys.argv[1:] or ["lorenz_dp5", "lorenz_dp8", "kepler_dp8"]
    n = int(os.environ.get("KBENCH_N", 1 << 20))
    reps = int(os.environ.get("KBENCH_REPS", 3))
    peak, _ = ob.measure_fp64_peak(0)
    ctx = ob.default_context(0)
    if os.environ.get("KBENCH_NOSPLIT"):
        ctx.set_lane_split(False)
    for q in (1, 0) if os.environ.get("KBENCH_STATIC") else (1,):
        ctx.set_work_queue(bool(q))
        for name in names:
            if name == "f32":
                f32_rows(n, reps, ctx)
                continue
            nn = n if not any(k in name for k in ("threebody", "dense", "all_steps", "com")) else min(n, 1 << 18)
            sysname, cfg, y0, p, cap, flops = build(name, nn)
            f = ob.System.builtin(sysname)
            best = 1e30
            com_cap = -cap if cap < 0 else 0
            cap = abs(cap)
            for _ in range(reps):
                o = ob.integrate_batch(f, cfg, y0, p, out_cap=cap, com_cap=com_cap, ctx=ctx)
                best = min(best, o.kernel_ms)
            if cap:
                dim = y0.shape[0]
                stored = int(np.minimum(o.out_count, cap).sum()) * (1 + dim) * 8
                if com_cap:
                    m = 5 if cfg.method == ob.DOPRI5 else 8
                    stored += int(np.minimum(o.com_count, com_cap).sum()) * (m * dim + 2) * 8
                print("   output bytes written %.3f GB -> %.1f GB/s of HBM write (algorithmic)" % (
                    stored / 1e9, stored / 1e9 / (best * 1e-3)))
            acc, att = int(o.accepted.sum()), int(o.n_step.sum())
            print("%-18s queue=%d n=%d kernel %.3f ms  %.3e acc-steps/s  %.2f TFLOP/s alg (%.1f%% of unfused FP64 "
                  "peak)  acc/att=%d/%d bad=%d" % (name, q, nn, best, acc / best * 1e3, flops * att / best / 1e9,
                                                 100 * flops * att / (best * 1e-3) / peak, acc, att,
                                                 int((o.status != 0).sum())), flush=True)


if __name__ == "__main__":
    main()
