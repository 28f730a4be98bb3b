#!/usr/bin/env python3
"""End-to-end time of ode_b200_integrate_batch (pinned host buffers in, pinned host buffers out) in its three
transfer modes, for the two bench workloads: pipeline (chunked upload / integrate / copy-back, the default),
one-shot copies, and zero-copy result stores. Prints ms per call (best of 5) and the kernel time inside."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
import ode_solvers_b200 as ob  # noqa: E402
from ode_solvers_b200 import _abi  # noqa: E402


def main():
    lib = _abi.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = ob.Context(0)
    for which, n in (("lorenz", 1 << 20), ("kepler", 1 << 22)):
        run = bench.DeviceRun(torch, ob, lib, ctx, dev, which, n, 0)
        stream = torch.cuda.current_stream()
        for _ in range(2):
            run.launch_device(stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run.launch_device(stream)
        e1.record(stream)
        torch.cuda.synchronize()
        print("%-7s n=%d  kernel path (device-resident): %.3f ms" % (which, n, e0.elapsed_time(e1)), flush=True)
        for label, pipe, zc in (("pipeline 2^16", 16, 0), ("pipeline 2^15", 15, 0), ("pipeline 2^17", 17, 0),
                                ("one-shot copies", 0, 0), ("zero-copy stores", 0, 1)):
            ctx.set_pipeline(pipe)
            ctx.set_zero_copy(bool(zc))
            best, kms = 1e30, 0.0
            for _ in range(6):
                t0 = time.perf_counter()
                run.run_e2e()
                dt = (time.perf_counter() - t0) * 1e3
                if dt < best:
                    best, kms = dt, ctx.last_kernel_ms
            print("   %-18s e2e %.3f ms   (kernel inside %.3f ms)" % (label, best, kms), flush=True)
        del run
    ctx.close()


if __name__ == "__main__":
    main()
