#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native ensemble ODE integrator.

Metric (BASELINE.json): accepted trajectory-steps/sec (f64) = sum over trajectories of
Stats.accepted_steps / device seconds.

Workload (BASELINE.json configs[1]): Lorenz ensemble, 2^20 random initial conditions PER GPU
(weak scaling; y0 ~ U[-20,20]x[-25,25]x[0,50] from a counter-based RNG keyed by the global
trajectory index), sigma=10, beta=8/3, rho=28, x in [0, 20], Dopri5 rtol=atol=1e-6,
OutputType::Sparse with final-state outputs. One "step" = one pass of the hot path (one
persistent-kernel launch) over the whole batch.

  value      kernel path, inputs resident in HBM (CUDA events on the launching stream)
  e2e        the same work through the host-buffer C-ABI call (pinned host memory ->
             H2D + kernel + D2H of final state, Stats and status), i.e. what a reference-side
             FFI binding pays
  roofline   FP64-pipe bound (not HBM, not tensor): algorithmic flops / kernel time against
             the FP64 peak measured live by a DFMA-chain probe
  cpu_baseline  the CPU oracle (operation-order port of the crate) on the host cores,
             bounded sample

`--impl reference` times the reference's CPU path (the oracle port: the Rust crate cannot be
built in this image) on all host threads with the same config/metric.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

N_PER_GPU = 1 << 20
X_END = 20.0
TOL = 1e-6
METRIC = "accepted trajectory-steps/sec (f64)"
UNIT = "accepted_steps/s"
# nominal algorithmic FP64 ops per ATTEMPTED Dopri5 step, 63n + 6F + 25 with n = 3, F = 8
FLOPS_PER_ATTEMPT = 63 * 3 + 6 * 8 + 25
# from the committed ncu capture of this kernel on this workload (profiles/r1_ncu_summaries.json,
# "r1e_lorenz_dopri5_final")
NCU_DRAM_BYTES_PER_LAUNCH = 32995072 + 34777600
NCU_FP64_PIPE_BUSY = 0.702  # sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active / 100


def workload_config(n_gpus):
    return {
        "workload": "lorenz_ensemble_dopri5",
        "system": "lorenz (sigma=10, beta=8/3, rho=28), y0~U[-20,20]x[-25,25]x[0,50]",
        "method": "Dopri5 rtol=atol=1e-6, x 0->20, OutputType::Sparse, final-state outputs",
        "trajectories_per_gpu": N_PER_GPU,
        "trajectories_total": N_PER_GPU * n_gpus,
        "sharding": "contiguous index ranges, one rank per GPU, no collective on the step path",
        "l2": "flushed between timed steps (256 MiB write, outside the per-step event pairs)",
    }


# ---------------------------------------------------------------- reference arm (CPU)
def oracle_rate(n_sample, offset, threads=0):
    """accepted steps/s of the CPU oracle on trajectories [offset, offset + n_sample)."""
    import oracle
    from ode_solvers_b200 import _abi, workloads
    y0, p = workloads.lorenz_ensemble(n_sample, offset)
    cfg = oracle.config_new(_abi.DOPRI5, 0.0, X_END, 0.1, TOL, TOL, out_type=_abi.OUT_SPARSE)
    t0 = time.perf_counter()
    o = oracle.integrate_batch("lorenz", cfg, y0, p, n_threads=threads)
    dt = time.perf_counter() - t0
    return float(o.accepted.sum()) / dt, dt, int(o.accepted.sum())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    cores = oracle.lib().ode_oracle_max_threads()
    rate, dt, _ = oracle_rate(2048, 0)  # calibrate the sample so one step is ~2 s of CPU work
    n_sample = int(min(N_PER_GPU, max(2048, 2048 * 2.0 / max(dt, 1e-3))))
    for _ in range(args.warmup):
        oracle_rate(n_sample, 0)
    tot_steps, tot_t = 0, 0.0
    for k in range(args.steps):
        r, dt, acc = oracle_rate(n_sample, (k * n_sample) % N_PER_GPU)
        tot_steps += acc
        tot_t += dt
    value = tot_steps / tot_t
    sample = "%d of %d trajectories per step (first indices of the ensemble), OpenMP over trajectories" % (
        n_sample, N_PER_GPU)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "C port of the crate's arithmetic without its per-term temporaries: a faster "
                                 "stand-in for the Rust crate, which cannot be built here (no cargo/rustc)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, row in self.rows:
            f = [c.strip() for c in row.split(",")]
            if len(f) < 7:
                continue
            try:
                mx.append(float(f[1]))
                if t0 - 0.05 <= t <= t1 + 0.05:
                    sm.append(float(f[0]))
                    for nm, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples_in_region": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------- the other configs (GPU)
def other_workloads(ob, W, ctx, peak_inst):
    """Kernel time (CUDA events inside the library, best of 2 after one warm-up) of the other
    BASELINE.json configs through the host-buffer entry: steps/s, nominal algorithmic TFLOP/s as a
    fraction of the unfused FP64 peak (flop formulas: BASELINE.md section 2), and GB/s of recorded
    samples for the output-heavy runs. Parity of every one of these kernels is a tests/ matter."""
    per = W.kepler_period()
    cases = [
        ("configs[1] Lorenz 2^20, Rk4 h=1e-3 (20000 steps each), final state", "lorenz", W.lorenz_ensemble, 1 << 20,
         ob.config_rk4(0.0, 20.0, 1e-3), 0, 13 * 3 + 4 * 8 + 4),
        ("configs[1] Lorenz 2^20, Dop853 1e-6", "lorenz", W.lorenz_ensemble, 1 << 20,
         ob.config_new(ob.DOP853, 0.0, 20.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_SPARSE), 0, 146 * 3 + 11 * 8 + 90),
        ("configs[2] Kepler e-sweep 2^22 (4M), Dop853 1e-10, 5 periods", "kepler", W.kepler_sweep, 1 << 22,
         ob.config_new(ob.DOP853, 0.0, 5.0 * per, 60.0, 1e-10, 1e-10, out_type=ob.OUT_SPARSE), 0, 146 * 6 + 11 * 14 + 90),
        ("configs[3] CR3BP 2^18, Dop853 1e-12, Dense dx=0.5", "cr3bp", W.cr3bp_ensemble, 1 << 18,
         ob.config_new(ob.DOP853, 0.0, 20.0, 0.5, 1e-12, 1e-12), 43, 146 * 6 + 11 * 39 + 90),
        ("configs[3] Newtonian 3-body SVector<f64,18> 2^18, Dop853 1e-9", "threebody18", W.threebody18_ensemble, 1 << 18,
         ob.config_new(ob.DOP853, 0.0, 5.0, 0.1, 1e-9, 1e-9, out_type=ob.OUT_SPARSE), 0, 146 * 18 + 11 * 90 + 90),
        ("configs[4] Robertson rate sweep 2^20, Dop853 rtol=1e-2 atol=1e-6", "robertson", W.robertson_sweep, 1 << 20,
         ob.config_new(ob.DOP853, 0.0, 0.3, 0.3, 1e-2, 1e-6, out_type=ob.OUT_SPARSE), 0, 146 * 3 + 11 * 10 + 90),
        ("output-heavy: Lorenz 2^17, Rk4 h=1e-3 0->2, every step recorded (2001 x 32 B each)", "lorenz",
         W.lorenz_ensemble, 1 << 17, ob.config_rk4(0.0, 2.0, 1e-3), 2004, 13 * 3 + 4 * 8 + 4),
        ("output-heavy: Lorenz 2^17, Dopri5 1e-6 Dense dx=0.01 (2001 samples each)", "lorenz", W.lorenz_ensemble,
         1 << 17, ob.config_new(ob.DOPRI5, 0.0, 20.0, 0.01, 1e-6, 1e-6), 2004, 63 * 3 + 6 * 8 + 25 + 12 * 3),
    ]
    rows = []
    for label, sysname, gen, n, cfg, cap, flops in cases:
        try:
            rows.append(_one_workload(ob, ctx, peak_inst, label, sysname, gen, n, cfg, cap, flops))
        except Exception as e:
            rows.append({"workload": label, "error": repr(e)})
    return rows


def _one_workload(ob, ctx, peak_inst, label, sysname, gen, n, cfg, cap, flops):
    y0, p = gen(n)
    f = ob.System.builtin(sysname)
    best = None
    for _ in range(2 if cap else 3):
        o = ob.integrate_batch(f, cfg, y0, p, out_cap=cap, ctx=ctx)
        best = o.kernel_ms if best is None else min(best, o.kernel_ms)
    acc, att = int(o.accepted.sum()), int(o.n_step.sum())
    row = {"workload": label, "trajectories": n, "kernel_ms": round(best, 3),
           "accepted_steps_per_s": acc / (best * 1e-3), "attempted_steps": att,
           "algorithmic_tflops": flops * att / (best * 1e-3) / 1e12,
           "unfused_frac": flops * att / (best * 1e-3) / peak_inst, "trajectories_not_ok": int((o.status != 0).sum())}
    if row["trajectories_not_ok"]:
        st, cnt = np.unique(o.status[o.status != 0], return_counts=True)
        names = {1: "MaxNumStepReached", 2: "StepSizeUnderflow", 3: "StiffnessDetected", 4: "OutputCapacityExceeded",
                 5: "UnsupportedDomain"}
        row["not_ok_statuses"] = {names.get(int(a), str(int(a))): int(b) for a, b in zip(st, cnt)}
        row["not_ok_note"] = "IntegrationError results of the reference for these inputs (same statuses in the oracle)"
    if cap:
        stored = int(np.minimum(o.out_count, cap).sum()) * (1 + y0.shape[0]) * 8
        row["sample_bytes"] = stored
        row["sample_write_gbs"] = stored / (best * 1e-3) / 1e9
    del o
    return row


# ---------------------------------------------------------------- native arm (GPU)
def run_native(args):
    import torch
    import ode_solvers_b200 as ob
    from ode_solvers_b200 import _abi, workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lib = _abi.load()
    ctx = ob.Context(local)
    # Up to 4 ranks per host the kernel stores its per-trajectory results straight into the pinned host
    # arrays; with 8 ranks sharing the host's PCIe one DMA copy per array after the kernel is faster.
    # Measured on this pool, ms per end-to-end pass at 1/2/4/8 ranks: zero-copy 29.2 / 29.2 / 30.0 /
    # 37.5, staged copies 30.3 / 30.7 / 32.6 / 33.0 (profiles/r1_scaling_1_2_4_8.json).
    zero_copy = world <= 4
    ctx.set_zero_copy(zero_copy)
    n = N_PER_GPU
    f = ob.System.builtin("lorenz")
    cfg = ob.config_new(ob.DOPRI5, 0.0, X_END, 0.1, TOL, TOL, out_type=ob.OUT_SPARSE)

    # ---- inputs: this rank's contiguous index range of the global ensemble
    y0_h, p_h = workloads.lorenz_ensemble(n, offset=rank * n)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)  # noqa: E731
    y0_pin = pin((3, n), torch.float64)
    y0_pin.numpy()[:] = y0_h
    p_pin = pin((3,), torch.float64)
    p_pin.numpy()[:] = p_h

    # device-resident buffers (torch owns the memory; kernels are ours)
    d_y0 = y0_pin.to(dev)
    d_p = p_pin.to(dev)
    d = {k: torch.empty(n, dtype=torch.float64, device=dev) for k in ("x_final", "h_next")}
    d["y_final"] = torch.empty((3, n), dtype=torch.float64, device=dev)
    for k in ("num_eval", "accepted", "rejected", "n_step", "status"):
        d[k] = torch.empty(n, dtype=torch.int32, device=dev)

    def outputs_struct(bufs, ptr):
        o = _abi.Outputs()
        for k in ("y_final", "x_final", "h_next"):
            setattr(o, k, C.cast(ptr(bufs[k]), _abi.c_double_p))
        for k in ("num_eval", "accepted", "rejected", "n_step"):
            setattr(o, k, C.cast(ptr(bufs[k]), _abi.c_u32_p))
        o.status = C.cast(ptr(bufs["status"]), _abi.c_i32_p)
        return o

    o_dev = outputs_struct(d, lambda t: t.data_ptr())
    stream = torch.cuda.current_stream()

    def launch_device():
        rc = lib.ode_b200_integrate_batch_device(ctx.handle, f.sys_id, C.byref(cfg), n, d_y0.data_ptr(),
                                                 d_p.data_ptr(), 0, C.byref(o_dev), stream.cuda_stream)
        if rc != 0:
            raise RuntimeError(_abi.last_error())

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- FP64 peak probe (roofline denominator), before the timed region
    peak_inst, _mhz = ob.measure_fp64_peak(local)

    # ---- kernel path: W warm-up, K timed steps
    for _ in range(args.warmup):
        launch_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.25 if rank == 0 else 0.0)
    barrier()
    launches0 = ctx.launch_count
    t_wall0 = time.perf_counter()
    evs = []
    for _ in range(args.steps):
        flush.zero_()  # evict L2 between timed steps (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launch_device()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    t_wall1 = time.perf_counter()
    gpu_launches = ctx.launch_count - launches0
    step_ms = [a.elapsed_time(b) for a, b in evs]
    my_ms = float(sum(step_ms))
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    accepted = int(d["accepted"].to(torch.int64).sum().item())
    attempts = int(d["n_step"].to(torch.int64).sum().item())
    n_bad = int((d["status"] != 0).sum().item())

    # ---- end to end through the host-buffer C-ABI call (pinned host buffers)
    h = {k: pin(tuple(d[k].shape), d[k].dtype) for k in d}
    o_host = outputs_struct(h, lambda t: t.data_ptr())
    h2d = y0_pin.numel() * 8 + p_pin.numel() * 8
    d2h = sum(t.numel() * t.element_size() for t in h.values())

    def run_e2e():
        rc = lib.ode_b200_integrate_batch(ctx.handle, f.sys_id, C.byref(cfg), n, y0_pin.data_ptr(),
                                          p_pin.data_ptr(), 0, C.byref(o_host))
        if rc != 0:
            raise RuntimeError(_abi.last_error())

    for _ in range(min(args.warmup, 2)):
        run_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_kernel_ms = ctx.last_kernel_ms
    acc_e2e = int(h["accepted"].to(torch.int64).sum().item())
    assert acc_e2e == accepted, "e2e and device paths disagree"

    # ---- reduce over ranks: max time, sum work
    from ode_solvers_b200 import sharding
    my_ms, (tot_acc, tot_att, n_bad) = sharding.reduce_report(dist, dev, my_ms, [accepted, attempts, n_bad])
    e2e_s, _ = sharding.reduce_report(dist, dev, e2e_s, [0])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return None

    ms_per_step = my_ms / args.steps
    value = tot_acc / (ms_per_step * 1e-3)
    e2e_value = tot_acc * args.steps / e2e_s

    # ---- roofline of the dominant (only) kernel, this rank's launch
    flops_launch = FLOPS_PER_ATTEMPT * attempts
    kern_s = (float(sum(step_ms)) / args.steps) * 1e-3
    peak_tflops_fma = 2.0 * peak_inst / 1e12
    achieved = flops_launch / kern_s / 1e12
    roofline = {
        "bound": "fp64", "achieved": achieved, "peak": peak_tflops_fma, "unit": "TFLOP/s",
        "frac": achieved / peak_tflops_fma, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this "
                          "kernel on this workload (profiles/r1_ncu_summaries.json, r1e_lorenz_dopri5_final): 33.0 MB + 34.8 MB",
        "fp64_pipe_busy_ncu": NCU_FP64_PIPE_BUSY,
        "kernel": "ode_step_loop_kernel<Dopri5Stepper<SysLorenz, Sparse>>",
        "peak_source": "measured live: DFMA-chain probe (ode_b200_measure_fp64_peak), FMA counted as 2 flops; "
                       "MEASURED_PEAKS.json has no FP64 entry",
        "fp64_pipe_peak_inst_per_s": peak_inst,
        "unfused_frac": achieved / (peak_inst / 1e12),
        "note": "parity forbids FMA contraction, so every algorithmic flop is one FP64-pipe instruction: "
                "frac <= 0.5 by construction; unfused_frac = algorithmic flops / (pipe instructions/s peak). "
                "Algorithmic flops = 262 per attempted step (63n+6F+25, n=3, F=8); divisions, sqrt and the two "
                "pow() per step count as 1 flop each although they cost 10-60 pipe instructions (see profiles/).",
        "algorithmic_bytes_per_launch": (3 * 8 + 3 * 8 + 2 * 8 + 5 * 4) * n,
    }

    # ---- CPU baseline on a bounded sample (rank 0 only, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        import oracle
        cores = oracle.lib().ode_oracle_max_threads()
        _r, dt, _a = oracle_rate(2048, 0)
        ns = int(min(n, max(2048, 2048 * 10.0 / max(dt, 1e-3))))
        r, dt, _a = oracle_rate(ns, 0)
        cpu = {"value": r, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "first %d of %d trajectories, %.1f s, OpenMP schedule(dynamic) over trajectories" % (
                   ns, n, dt)}

    # ---- the other BASELINE.json configs on the same kernels (N = 1 only; reported, not the metric)
    others = None
    if world == 1 and not args.no_extra:
        try:
            others = other_workloads(ob, workloads, ctx, peak_inst)
        except Exception as e:  # reported workloads must never cost the headline line
            others = {"error": repr(e)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s / args.steps, "kernel_ms_inside": e2e_kernel_ms,
                "api": "ode_b200_integrate_batch (host pointers, pinned): y0/params by cudaMemcpyAsync, per-trajectory "
                       "results " + ("written by the kernel straight into the pinned host arrays over PCIe" if zero_copy
                                     else "copied back by cudaMemcpyAsync (ode_b200_set_zero_copy(ctx, 0): 8 "
                                          "ranks share the host's PCIe)")},
        "gpu_launches": int(gpu_launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "other_workloads": others,
        "accepted_steps_per_pass": tot_acc, "attempted_steps_per_pass": tot_att,
        "trajectories_not_ok": n_bad,
        "step_ms": [round(v, 4) for v in step_ms],
    }
    if dist is not None:
        dist.destroy_process_group()
    return line


class QuietStdout:
    """Route fd 1 to stderr while libraries initialise (NCCL prints its version banner to stdout),
    so that the ONE JSON line is the only thing on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the other_workloads leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    with QuietStdout():
        line = run_native(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
