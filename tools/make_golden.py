#!/usr/bin/env python3
"""Generate tests/golden/full_size_checksums.json with the CPU ORACLE (test infrastructure):
checksums of the oracle's results on BASELINE.json's full-size ensembles, so that the GPU tier can
prove bitwise parity at full size without re-running minutes of CPU work on the GPU box.

  lorenz_dopri5_1M           all 2^20 trajectories of config 2 (Dopri5 rtol=atol=1e-6, 0->20)
  kepler_dop853_4M_stride64  every 64th of the 2^22 trajectories of config 3 (Dop853 1e-10, 5 periods)

sha256 is taken over the raw little-endian bytes of the SoA arrays, so equality means every
trajectory's final state / step count is bit-identical.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from ode_solvers_b200 import _abi as A  # noqa: E402
from ode_solvers_b200 import workloads as W  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def summary(o):
    return {"accepted": int(o.accepted.sum()), "rejected": int(o.rejected.sum()), "n_step": int(o.n_step.sum()),
            "num_eval": int(o.num_eval.sum()), "not_ok": int((o.status != 0).sum()),
            "sha256_y_final": sha(o.y_final), "sha256_accepted": sha(o.accepted), "sha256_h_next": sha(o.h_next)}


def main():
    gold = {}
    n = 1 << 20
    y0, p = W.lorenz_ensemble(n)
    cfg = oracle.config_new(A.DOPRI5, 0.0, 20.0, 0.1, 1e-6, 1e-6, out_type=A.OUT_SPARSE)
    gold["lorenz_dopri5_1M"] = summary(oracle.integrate_batch("lorenz", cfg, y0, p))
    n = 1 << 22
    y0, p = W.kepler_sweep(n)
    cfg = oracle.config_new(A.DOP853, 0.0, 5 * W.kepler_period(), 60.0, 1e-10, 1e-10, out_type=A.OUT_SPARSE)
    gold["kepler_dop853_4M_stride64"] = summary(oracle.integrate_batch("kepler", cfg, y0[:, ::64], p))
    out = os.path.join(ROOT, "tests", "golden", "full_size_checksums.json")
    with open(out, "w") as f:
        json.dump(gold, f, indent=1)
    print(json.dumps(gold, indent=1))


if __name__ == "__main__":
    main()
