#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native ensemble ODE integrator.

Metric (BASELINE.json): accepted trajectory-steps/sec (f64) = sum over trajectories of
Stats.accepted_steps / device seconds.

Headline workload (BASELINE.json configs[1]): Lorenz ensemble, 2^20 random initial conditions PER GPU
(weak scaling; y0 ~ U[-20,20]x[-25,25]x[0,50] from a counter-based RNG keyed by the global
trajectory index), sigma=10, beta=8/3, rho=28, x in [0, 20], Dopri5 rtol=atol=1e-6,
OutputType::Sparse with final-state outputs. One "step" = one pass of the hot path (one
persistent-kernel launch) over the whole batch.

  value      kernel path, inputs resident in HBM (CUDA events on the launching stream)
  e2e        the same work through the host-buffer C-ABI call (pinned host memory ->
             H2D + kernel + D2H of final state, Stats and status), i.e. what a reference-side
             FFI binding pays
  roofline   FP64-pipe bound (not HBM, not tensor): algorithmic flops / kernel time against
             the FP64 peak measured live by a DFMA-chain probe; `traffic` and the FP64-pipe
             utilisation are read at run time from the committed ncu summaries under profiles/
             (newest round first) and the run FAILS if no capture of the kernel is committed
  cpu_baseline  the CPU oracle (operation-order port of the crate) on the host cores,
             bounded sample, thread count passed explicitly (never OMP_NUM_THREADS)

`co_headline` is the second full record the north star asks for (BASELINE.json configs[2]): the
Kepler eccentricity sweep, 2^22 = 4M trajectories IN TOTAL (strong scaling: each of the N ranks takes
the contiguous range of 4M/N), Dop853 rtol=atol=1e-10 over 5 periods. It carries its own value
(kernel path, max over ranks), roofline, cpu_baseline (N = 1) and two end-to-end numbers with host
buffers: `e2e` through ode_b200_integrate_batch on each rank's range, and `e2e_multi` through the
single-process entry ode_b200_integrate_batch_multi (contiguous ranges over the N GPUs of the box,
final host gather into ONE set of arrays), run by rank 0 after the other ranks are done.

`--impl reference` times the reference's CPU path (the oracle port: the Rust crate cannot be
built in this image) on ALL host cores with the same config/metric.
"""
import argparse
import ctypes as C
import glob
import json
import os
import re
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
# nominal algorithmic FP64 ops per ATTEMPTED step: Dopri5 63n + 6F + 25,
# Dop853 146n + 11F + 90; Lorenz n = 3, F = 8; Kepler n = 6, F = 14
FLOPS_LORENZ_DOPRI5 = 63 * 3 + 6 * 8 + 25
FLOPS_KEPLER_DOP853 = 146 * 6 + 11 * 14 + 90
KEPLER_N_TOTAL = 1 << 22
KEPLER_TOL = 1e-10
# labels of the committed `ncu --set full` captures of the two kernels AT THE BENCH'S SIZES
# (profiles/r*_ncu_summaries.json, made with tools/ncu_capture.sh + tools/ncu_summary.py); the first label
# found in the newest round's file wins
NCU_LABELS_HEADLINE = ["bench_lorenz_dopri5_2p20", "r1e_lorenz_dopri5_final"]
NCU_LABELS_KEPLER = ["bench_kepler_dop853_2p22"]


def host_cores():
    """cores this process may run on: the thread count handed to the oracle EXPLICITLY, so that an
    exported OMP_NUM_THREADS (torchrun sets it to 1) cannot throttle the reference arm"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def ncu_record(labels):
    """dram bytes per launch and FP64-pipe utilisation of a kernel from the committed ncu summaries."""
    def rnd(path):
        m = re.search(r"r(\d+)_ncu_summaries", os.path.basename(path))
        return int(m.group(1)) if m else -1
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_summaries.json")), key=rnd, reverse=True)
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for path in files:
        with open(path) as fh:
            d = json.load(fh)
        for lb in labels:
            if lb in d:
                e = d[lb]

                def by(k):
                    return float(e[k]["value"]) * unit[e[k]["unit"]]
                return {
                    "traffic": by("dram__bytes_read.sum") + by("dram__bytes_write.sum"),
                    "dram_read": by("dram__bytes_read.sum"), "dram_write": by("dram__bytes_write.sum"),
                    "fp64_pipe_busy": float(e["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]["value"]) / 100,
                    "ncu_kernel_ms": float(e["gpu__time_duration.sum"]["value"]),
                    "kernel": e.get("kernel"),
                    "source": "profiles/%s [%s]" % (os.path.basename(path), lb),
                }
    raise SystemExit("bench.py: no committed ncu capture for any of %r under profiles/r*_ncu_summaries.json -- "
                     "capture it (tools/ncu_capture.sh) before reporting a roofline" % (labels,))


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


def kepler_config(n_gpus):
    return {
        "workload": "kepler_esweep_4M_dop853",
        "system": "kepler_orbit.rs two-body problem, mu=398600.435436, a=20000 km, e~U[0,0.9], start at periapsis",
        "method": "Dop853 rtol=atol=1e-10, x 0 -> 5 periods, OutputType::Sparse, final-state outputs",
        "trajectories_total": KEPLER_N_TOTAL,
        "trajectories_per_gpu": KEPLER_N_TOTAL // n_gpus,
        "scaling": "strong: the 4M trajectories are split into contiguous ranges over the GPUs",
    }


# ---------------------------------------------------------------- workloads
def make_cfg(mod, which):
    """mod: the module that builds configs (ode_solvers_b200 or the oracle wrapper)"""
    from ode_solvers_b200 import _abi, workloads
    if which == "lorenz":
        return mod.config_new(_abi.DOPRI5, 0.0, X_END, 0.1, TOL, TOL, out_type=_abi.OUT_SPARSE)
    return mod.config_new(_abi.DOP853, 0.0, 5.0 * workloads.kepler_period(), 60.0, KEPLER_TOL, KEPLER_TOL,
                          out_type=_abi.OUT_SPARSE)


def gen_inputs(which, n, offset):
    from ode_solvers_b200 import workloads
    return workloads.lorenz_ensemble(n, offset) if which == "lorenz" else workloads.kepler_sweep(n, offset)


# ---------------------------------------------------------------- reference arm (CPU)
def oracle_rate(which, n_sample, offset, threads):
    """accepted steps/s of the CPU oracle on trajectories [offset, offset + n_sample)."""
    import oracle
    y0, p = gen_inputs(which, n_sample, offset)
    cfg = make_cfg(oracle, which)
    t0 = time.perf_counter()
    o = oracle.integrate_batch("lorenz" if which == "lorenz" else "kepler", cfg, y0, p, n_threads=threads)
    dt = time.perf_counter() - t0
    return float(o.accepted.sum()) / dt, dt, int(o.accepted.sum())


def cpu_baseline(which, n_total, seconds, threads):
    """bounded sample: calibrate on 2048 trajectories, then one run of about `seconds` of CPU work"""
    _r, dt, _a = oracle_rate(which, 2048, 0, threads)
    ns = int(min(n_total, max(2048, 2048 * seconds / max(dt, 1e-3))))
    r, dt, _a = oracle_rate(which, ns, 0, threads)
    return {"value": r, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "first %d of %d trajectories, %.1f s, OpenMP schedule(dynamic) over trajectories, "
                      "num_threads(%d) passed explicitly" % (ns, n_total, dt, threads)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    rate, dt, _ = oracle_rate("lorenz", 2048, 0, cores)  # calibrate the sample so one step is ~2 s of CPU work
    n_sample = int(min(N_PER_GPU, max(2048, 2048 * 2.0 / max(dt, 1e-3))))
    for _ in range(args.warmup):
        oracle_rate("lorenz", n_sample, 0, cores)
    tot_steps, tot_t = 0, 0.0
    for k in range(args.steps):
        r, dt, acc = oracle_rate("lorenz", n_sample, (k * n_sample) % N_PER_GPU, cores)
        tot_steps += acc
        tot_t += dt
    value = tot_steps / tot_t
    sample = "%d of %d trajectories per step (consecutive index ranges of the ensemble), OpenMP over trajectories, " \
             "num_threads(%d) passed explicitly (OMP_NUM_THREADS=%s is ignored)" % (
                 n_sample, N_PER_GPU, cores, os.environ.get("OMP_NUM_THREADS", "unset"))
    note = ("C port of the crate's arithmetic without its per-term temporaries: a faster stand-in for the Rust "
            "crate, which cannot be built here (no cargo/rustc)")
    kep = cpu_baseline("kepler", KEPLER_N_TOTAL, 4.0, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "note": note},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "co_headline": {"metric": METRIC, "unit": UNIT, "config": kepler_config(args.gpus), "value": kep["value"],
                        "cpu_baseline": kep,
                        "e2e": {"value": kep["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
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
         ob.config_rk4(0.0, 20.0, 1e-3), 0, 0, 13 * 3 + 4 * 8 + 4),
        ("configs[1] Lorenz 2^20, Dop853 1e-6", "lorenz", W.lorenz_ensemble, 1 << 20,
         ob.config_new(ob.DOP853, 0.0, 20.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_SPARSE), 0, 0, 146 * 3 + 11 * 8 + 90),
        ("configs[2] Kepler e-sweep 2^20, Dopri5 1e-10, 5 periods", "kepler", W.kepler_sweep, 1 << 20,
         ob.config_new(ob.DOPRI5, 0.0, 5.0 * per, 60.0, 1e-10, 1e-10, out_type=ob.OUT_SPARSE), 0, 0, 63 * 6 + 6 * 14 + 25),
        ("configs[3] CR3BP 2^18, Dop853 1e-12, Dense dx=0.5", "cr3bp", W.cr3bp_ensemble, 1 << 18,
         ob.config_new(ob.DOP853, 0.0, 20.0, 0.5, 1e-12, 1e-12), 43, 0, 146 * 6 + 11 * 39 + 90),
        ("configs[3] Newtonian 3-body SVector<f64,18> 2^18, Dop853 1e-9", "threebody18", W.threebody18_ensemble, 1 << 18,
         ob.config_new(ob.DOP853, 0.0, 5.0, 0.1, 1e-9, 1e-9, out_type=ob.OUT_SPARSE), 0, 0, 146 * 18 + 11 * 90 + 90),
        ("configs[3] Newtonian 3-body SVector<f64,18> 2^16, Dop853 1e-9 into a ContinuousOutputModel (Continuous8)",
         "threebody18", W.threebody18_ensemble, 1 << 16,
         ob.config_new(ob.DOP853, 0.0, 5.0, 0.1, 1e-9, 1e-9, out_type=ob.OUT_CONTINUOUS8), 120, 120,
         146 * 18 + 11 * 90 + 90 + 153 * 18 + 3 * 90),
        ("configs[4] Robertson rate sweep 2^20, Dop853 rtol=1e-2 atol=1e-6", "robertson", W.robertson_sweep, 1 << 20,
         ob.config_new(ob.DOP853, 0.0, 0.3, 0.3, 1e-2, 1e-6, out_type=ob.OUT_SPARSE), 0, 0, 146 * 3 + 11 * 10 + 90),
        ("output-heavy: Lorenz 2^18, Rk4 h=1e-3 0->2, every step recorded (2001 x 32 B each)", "lorenz",
         W.lorenz_ensemble, 1 << 18, ob.config_rk4(0.0, 2.0, 1e-3), 2004, 0, 13 * 3 + 4 * 8 + 4),
        ("output-heavy: Lorenz 2^18, Dopri5 1e-6 Dense dx=0.01 (2001 samples each)", "lorenz", W.lorenz_ensemble,
         1 << 18, ob.config_new(ob.DOPRI5, 0.0, 20.0, 0.01, 1e-6, 1e-6), 2004, 0, 63 * 3 + 6 * 8 + 25 + 12 * 3),
    ]
    rows = []
    for label, sysname, gen, n, cfg, cap, com_cap, flops in cases:
        try:
            rows.append(_one_workload(ob, ctx, peak_inst, label, sysname, gen, n, cfg, cap, com_cap, flops))
        except Exception as e:
            rows.append({"workload": label, "error": repr(e)})
    return rows


def _one_workload(ob, ctx, peak_inst, label, sysname, gen, n, cfg, cap, com_cap, flops):
    y0, p = gen(n)
    f = ob.System.builtin(sysname)
    best = None
    for _ in range(2 if cap else 3):
        o = ob.integrate_batch(f, cfg, y0, p, out_cap=cap, com_cap=com_cap, ctx=ctx)
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
        if com_cap:
            m = 5 if cfg.method == ob.DOPRI5 else 8
            stored += int(np.minimum(o.com_count, com_cap).sum()) * (m * y0.shape[0] + 2) * 8
        row["sample_bytes"] = stored
        row["sample_write_gbs"] = stored / (best * 1e-3) / 1e9
        from ode_solvers_b200 import _abi  # noqa: F401
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            row["hbm_frac_of_measured_peak"] = row["sample_write_gbs"] / float(peaks["hbm_gbs"])
        except Exception:
            pass
    del o
    return row


# ---------------------------------------------------------------- native arm (GPU)
class DeviceRun:
    """One workload on one rank: device-resident inputs/outputs (torch owns the memory, the kernels are
    ours), pinned host mirrors, the two timed paths."""

    def __init__(self, torch, ob, lib, ctx, dev, which, n, offset):
        from ode_solvers_b200 import _abi
        self.torch, self.lib, self.ctx, self.dev, self._abi = torch, lib, ctx, dev, _abi
        self.n = n
        self.f = ob.System.builtin("lorenz" if which == "lorenz" else "kepler")
        self.cfg = make_cfg(ob, which)
        y0_h, p_h = gen_inputs(which, n, offset)
        self.dim = y0_h.shape[0]
        pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)  # noqa: E731
        self.y0_pin = pin((self.dim, n), torch.float64)
        self.y0_pin.numpy()[:] = y0_h
        self.p_pin = pin((len(p_h),), torch.float64)
        self.p_pin.numpy()[:] = p_h
        self.d_y0 = self.y0_pin.to(dev)
        self.d_p = self.p_pin.to(dev)
        self.d = {k: torch.empty(n, dtype=torch.float64, device=dev) for k in ("x_final", "h_next")}
        self.d["y_final"] = torch.empty((self.dim, n), dtype=torch.float64, device=dev)
        for k in ("num_eval", "accepted", "rejected", "n_step", "status"):
            self.d[k] = torch.empty(n, dtype=torch.int32, device=dev)
        self.o_dev = self._outputs(self.d)
        self.h = {k: pin(tuple(self.d[k].shape), self.d[k].dtype) for k in self.d}
        self.o_host = self._outputs(self.h)
        self.h2d = self.y0_pin.numel() * 8 + self.p_pin.numel() * 8
        self.d2h = sum(t.numel() * t.element_size() for t in self.h.values())

    def _outputs(self, bufs):
        _abi = self._abi
        o = _abi.Outputs()
        for k in ("y_final", "x_final", "h_next"):
            setattr(o, k, C.cast(bufs[k].data_ptr(), _abi.c_double_p))
        for k in ("num_eval", "accepted", "rejected", "n_step"):
            setattr(o, k, C.cast(bufs[k].data_ptr(), _abi.c_u32_p))
        o.status = C.cast(bufs["status"].data_ptr(), _abi.c_i32_p)
        return o

    def launch_device(self, stream):
        rc = self.lib.ode_b200_integrate_batch_device(self.ctx.handle, self.f.sys_id, C.byref(self.cfg), self.n,
                                                      self.d_y0.data_ptr(), self.d_p.data_ptr(), 0,
                                                      C.byref(self.o_dev), stream.cuda_stream)
        if rc != 0:
            raise RuntimeError(self._abi.last_error())

    def run_e2e(self):
        rc = self.lib.ode_b200_integrate_batch(self.ctx.handle, self.f.sys_id, C.byref(self.cfg), self.n,
                                               self.y0_pin.data_ptr(), self.p_pin.data_ptr(), 0, C.byref(self.o_host))
        if rc != 0:
            raise RuntimeError(self._abi.last_error())

    def timed_kernel_path(self, steps, warmup, flush, stream, barrier):
        torch = self.torch
        for _ in range(warmup):
            self.launch_device(stream)
        barrier()
        launches0 = self.ctx.launch_count
        t_wall0 = time.perf_counter()
        evs = []
        for _ in range(steps):
            flush.zero_()  # evict L2 between timed steps (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            self.launch_device(stream)
            e1.record(stream)
            evs.append((e0, e1))
        barrier()
        t_wall1 = time.perf_counter()
        step_ms = [a.elapsed_time(b) for a, b in evs]
        return step_ms, self.ctx.launch_count - launches0, (t_wall0, t_wall1)

    def counters(self):
        torch = self.torch
        return (int(self.d["accepted"].to(torch.int64).sum().item()), int(self.d["n_step"].to(torch.int64).sum().item()),
                int((self.d["status"] != 0).sum().item()))

    def timed_e2e(self, steps, warmup, barrier):
        for _ in range(warmup):
            self.run_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.run_e2e()
        self.torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return dt, self.ctx.last_kernel_ms, int(self.h["accepted"].to(self.torch.int64).sum().item())


def fp64_roofline(flops_per_attempt, attempts, kern_s, peak_inst, ncu, kernel_name, flop_note, alg_bytes):
    peak_tflops_fma = 2.0 * peak_inst / 1e12
    achieved = flops_per_attempt * attempts / kern_s / 1e12
    return {
        "bound": "fp64", "achieved": achieved, "peak": peak_tflops_fma, "unit": "TFLOP/s",
        "frac": achieved / peak_tflops_fma, "traffic": ncu["traffic"],
        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum (%.1f MB + %.1f MB) of the committed "
                          "`ncu --set full` capture of this kernel at this size, read at run time from %s "
                          "(ncu duration %.2f ms)" % (ncu["dram_read"] / 1e6, ncu["dram_write"] / 1e6, ncu["source"],
                                                     ncu["ncu_kernel_ms"]),
        "fp64_pipe_busy_ncu": ncu["fp64_pipe_busy"],
        "kernel": kernel_name, "ncu_kernel": ncu["kernel"],
        "peak_source": "measured live: DFMA-chain probe (ode_b200_measure_fp64_peak), FMA counted as 2 flops; "
                       "MEASURED_PEAKS.json has no FP64 entry",
        "fp64_pipe_peak_inst_per_s": peak_inst,
        "unfused_frac": achieved / (peak_inst / 1e12),
        "note": "parity forbids FMA contraction, so every algorithmic flop is one FP64-pipe instruction: "
                "frac <= 0.5 by construction; unfused_frac = algorithmic flops / (pipe instructions/s peak). " + flop_note,
        "algorithmic_bytes_per_launch": alg_bytes,
    }


def multi_gpu_leg(ob, lib, torch, n_gpus, steps):
    """Kepler 4M through ode_b200_integrate_batch_multi: ONE process, contiguous ranges over n_gpus GPUs,
    one set of pinned host arrays in, one set out (the final host gather), wall clock around the call."""
    from ode_solvers_b200 import _abi
    n = KEPLER_N_TOTAL
    f = ob.System.builtin("kepler")
    cfg = make_cfg(ob, "kepler")
    y0_h, p_h = gen_inputs("kepler", n, 0)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)  # noqa: E731
    y0 = pin((6, n), torch.float64)
    y0.numpy()[:] = y0_h
    p = pin((1,), torch.float64)
    p.numpy()[:] = p_h
    h = {"y_final": pin((6, n), torch.float64), "x_final": pin((n,), torch.float64), "h_next": pin((n,), torch.float64)}
    for k in ("num_eval", "accepted", "rejected", "n_step", "status"):
        h[k] = pin((n,), torch.int32)
    o = _abi.Outputs()
    for k in ("y_final", "x_final", "h_next"):
        setattr(o, k, C.cast(h[k].data_ptr(), _abi.c_double_p))
    for k in ("num_eval", "accepted", "rejected", "n_step"):
        setattr(o, k, C.cast(h[k].data_ptr(), _abi.c_u32_p))
    o.status = C.cast(h["status"].data_ptr(), _abi.c_i32_p)

    def call():
        rc = lib.ode_b200_integrate_batch_multi(f.sys_id, C.byref(cfg), n, y0.data_ptr(), p.data_ptr(), 0, n_gpus,
                                                C.byref(o))
        if rc != 0:
            raise RuntimeError(_abi.last_error())
    call()  # contexts, workspaces
    call()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    dt = time.perf_counter() - t0
    acc = int(h["accepted"].to(torch.int64).sum().item())
    bad = int((h["status"] != 0).sum().item())
    h2d = y0.numel() * 8 + 8
    d2h = sum(t.numel() * t.element_size() for t in h.values())
    return {"value": acc * steps / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / steps, "n_gpus": n_gpus, "steps": steps,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "accepted_steps_per_pass": acc,
            "trajectories_not_ok": bad,
            "api": "ode_b200_integrate_batch_multi (one process, one host thread + context + stream per GPU, contiguous "
                   "index ranges, outputs gathered into the caller's single set of pinned host arrays); wall clock"}


def run_native(args):
    import torch
    import ode_solvers_b200 as ob
    from ode_solvers_b200 import _abi, sharding, workloads

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    # fail before any GPU time is spent if the roofline's ncu evidence is not committed
    ncu_head, ncu_kep = ncu_record(NCU_LABELS_HEADLINE), ncu_record(NCU_LABELS_KEPLER)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lib = _abi.load()
    if args.no_pipeline:  # profiler runs only: ncu serialises kernels and cannot profile one that waits for host copies
        os.environ["ODE_B200_PIPELINE"] = "0"  # read by every context the library creates (the multi-GPU entry's too)
    ctx = ob.Context(local)
    # The host-buffer entry runs its default transfer mode: the chunked upload / integrate / copy-back pipeline
    # (include/ode_b200.h, ode_b200_set_pipeline): the kernel starts before its inputs have arrived and finished
    # chunks of 2^16 trajectories are copied back while the rest is still being integrated.
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- FP64 peak probe (roofline denominator), before the timed region
    peak_inst, _mhz = ob.measure_fp64_peak(local)

    # ================= headline: Lorenz / Dopri5, 2^20 per GPU (weak scaling)
    n = N_PER_GPU
    head = DeviceRun(torch, ob, lib, ctx, dev, "lorenz", n, rank * n)
    for _ in range(args.warmup):
        head.launch_device(stream)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.25 if rank == 0 else 0.0)
    step_ms, gpu_launches, (t_wall0, t_wall1) = head.timed_kernel_path(args.steps, 0, flush, stream, barrier)
    my_ms = float(sum(step_ms))
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    accepted, attempts, n_bad = head.counters()
    e2e_s, e2e_kernel_ms, acc_e2e = head.timed_e2e(args.steps, min(args.warmup, 2), barrier)
    assert acc_e2e == accepted, "e2e and device paths disagree"
    my_ms, (tot_acc, tot_att, n_bad) = sharding.reduce_report(dist, dev, my_ms, [accepted, attempts, n_bad])
    e2e_s, _ = sharding.reduce_report(dist, dev, e2e_s, [0])
    h2d, d2h = head.h2d, head.d2h
    del head

    # ================= co-headline: Kepler 4M / Dop853 1e-10 (strong scaling over the ranks)
    ksteps = max(2, min(args.steps, 5))
    lo, hi = sharding.shard_range(KEPLER_N_TOTAL, world, rank)
    kep = DeviceRun(torch, ob, lib, ctx, dev, "kepler", hi - lo, lo)
    kstep_ms, klaunches, _ = kep.timed_kernel_path(ksteps, 2, flush, stream, barrier)
    kacc, katt, kbad = kep.counters()
    ke2e_s, ke2e_kernel_ms, kacc_e2e = kep.timed_e2e(ksteps, 1, barrier)
    assert kacc_e2e == kacc, "Kepler: e2e and device paths disagree"
    k_ms, (ktot_acc, ktot_att, kbad) = sharding.reduce_report(dist, dev, float(sum(kstep_ms)), [kacc, katt, kbad])
    ke2e_s, _ = sharding.reduce_report(dist, dev, ke2e_s, [0])
    kh2d, kd2h, k_n_rank = kep.h2d, kep.d2h, kep.n
    del kep
    gpu_launches += klaunches

    if dist is not None:
        barrier()
        dist.destroy_process_group()
    if rank != 0:
        return None  # the other ranks are done: rank 0 alone drives all GPUs in the single-process leg below

    ms_per_step = my_ms / args.steps
    value = tot_acc / (ms_per_step * 1e-3)
    e2e_value = tot_acc * args.steps / e2e_s

    roofline = fp64_roofline(
        FLOPS_LORENZ_DOPRI5, attempts, (float(sum(step_ms)) / args.steps) * 1e-3, peak_inst, ncu_head,
        "ode_step_loop_kernel<Dopri5Stepper<SysLorenz, Sparse>>",
        "Algorithmic flops = 262 per attempted step (63n+6F+25, n=3, F=8); divisions, sqrt and the two pow() per "
        "step count as 1 flop each although they cost 10-60 pipe instructions (see profiles/).",
        (3 * 8 + 3 * 8 + 2 * 8 + 5 * 4) * n)

    k_ms_per_step = k_ms / ksteps
    co = {
        "metric": METRIC, "unit": UNIT, "config": kepler_config(world), "steps": ksteps, "warmup": 2,
        "value": ktot_acc / (k_ms_per_step * 1e-3), "ms_per_step": k_ms_per_step, "scaling": "strong",
        "accepted_steps_per_pass": ktot_acc, "attempted_steps_per_pass": ktot_att, "trajectories_not_ok": kbad,
        "step_ms": [round(v, 4) for v in kstep_ms],
        "e2e": {"value": ktot_acc * ksteps / ke2e_s, "unit": UNIT, "h2d_bytes_per_step": kh2d * world,
                "d2h_bytes_per_step": kd2h * world, "ms_per_step": 1e3 * ke2e_s / ksteps,
                "kernel_ms_inside": ke2e_kernel_ms,
                "api": "ode_b200_integrate_batch (host pointers, pinned) on each rank's contiguous range of the 4M"},
        "roofline": fp64_roofline(
            FLOPS_KEPLER_DOP853, katt, (float(sum(kstep_ms)) / ksteps) * 1e-3, peak_inst, ncu_kep,
            "ode_step_loop_kernel<Dop853Stepper<SysKepler, Sparse>>",
            "Algorithmic flops = 1120 per attempted step (146n+11F+90, n=6, F=14); this rank's launch (%d "
            "trajectories); the ncu capture is the single-GPU 4M launch." % k_n_rank,
            (6 * 8 + 6 * 8 + 2 * 8 + 5 * 4) * k_n_rank),
    }

    # ---- single-process multi-GPU entry (all N GPUs of the box; N = 1 exercises the same code path)
    try:
        avail = ob.device_count()
        co["e2e_multi"] = multi_gpu_leg(ob, lib, torch, min(world, avail), 3)
    except Exception as e:  # must never cost the headline line
        co["e2e_multi"] = {"error": repr(e)}

    # ---- CPU baselines on bounded samples (rank 0 only, N = 1 only), all cores, passed explicitly
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = host_cores()
        cpu = cpu_baseline("lorenz", n, 10.0, cores)
        co["cpu_baseline"] = cpu_baseline("kepler", KEPLER_N_TOTAL, 8.0, cores)

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
                "api": "ode_b200_integrate_batch (host pointers, pinned), default transfer mode = pipeline: the kernel is "
                       "launched first, y0 is uploaded in chunks of 2^16 trajectories behind it (lanes wait for their "
                       "chunk), finished chunks are copied back by cudaMemcpyAsync while the rest is integrated"},
        "gpu_launches": int(gpu_launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "co_headline": co,
        "other_workloads": others,
        "accepted_steps_per_pass": tot_acc, "attempted_steps_per_pass": tot_att,
        "trajectories_not_ok": n_bad,
        "step_ms": [round(v, 4) for v in step_ms],
    }
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
    ap.add_argument("--no-pipeline", action="store_true",
                    help="e2e leg without the upload / integrate / copy-back pipeline (for runs under ncu, which "
                         "serialises kernels and cannot profile one that waits for concurrent host copies)")
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
