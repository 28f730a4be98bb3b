"""CUDA kernels against (1) the reference crate's own golden vectors (tests/golden/reference_goldens.json)
through the batched C ABI, and (2) the full-size checksums of the oracle
(tests/golden/full_size_checksums.json, made by tools/make_golden.py): BASELINE.json's
1M-trajectory Lorenz ensemble and 4M-trajectory Kepler sweep, bit for bit."""
import hashlib
import json
import math
import os

import numpy as np
import pytest

import ode_solvers_b200 as ob
from ode_solvers_b200 import workloads as W

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", json.load(open(os.path.join(GOLD, "reference_goldens.json")))["cases"],
                         ids=lambda c: c["cite"])
def test_reference_golden_vectors(case):
    f = ob.System.builtin(case["system"], case.get("params"))
    if case["method"] == "rk4":
        cfg = ob.config_rk4(case["x"], case["x_end"], case["step"])
    else:
        m = ob.DOPRI5 if case["method"] == "dopri5" else ob.DOP853
        cfg = ob.config_new(m, case["x"], case["x_end"], case["dx"], case["rtol"], case["atol"])
    # a batch of identical trajectories: every lane must reproduce the reference's numbers
    n = 70
    y0 = np.repeat(np.asarray(case["y0"], dtype=float).reshape(-1, 1), n, axis=1)
    o = ob.integrate_batch(f, cfg, y0, out_cap=128)
    for i in (0, 31, 32, 69):
        cnt = int(o.out_count[i])
        for idx, want, tol in case.get("checks", []):
            assert abs(o.out_y[i, idx, 0] - want) < tol
        if "last_x" in case:
            assert abs(o.out_x[i, cnt - 1] - case["last_x"][0]) < case["last_x"][1]
        if "len" in case:
            assert cnt == case["len"]
        if "analytic" in case:
            for s in range(cnt):
                assert abs(o.out_y[i, s, 0] + math.log(o.out_x[i, s] ** 3 + 1.0)) < case["analytic_tol"]
    assert (o.out_count == o.out_count[0]).all() and (o.status == ob.OK).all()


def test_full_size_lorenz_1m_bitwise_checksum():
    g = json.load(open(os.path.join(GOLD, "full_size_checksums.json")))["lorenz_dopri5_1M"]
    n = 1 << 20
    y0, p = W.lorenz_ensemble(n)
    cfg = ob.config_new(ob.DOPRI5, 0.0, 20.0, 0.1, 1e-6, 1e-6, out_type=ob.OUT_SPARSE)
    o = ob.integrate_batch(ob.System.builtin("lorenz"), cfg, y0, p)
    assert int((o.status != 0).sum()) == g["not_ok"] == 0
    assert (int(o.accepted.sum()), int(o.rejected.sum()), int(o.n_step.sum()), int(o.num_eval.sum())) == (
        g["accepted"], g["rejected"], g["n_step"], g["num_eval"])
    assert sha(o.accepted) == g["sha256_accepted"]
    assert sha(o.h_next) == g["sha256_h_next"]
    assert sha(o.y_final) == g["sha256_y_final"]  # all 2^20 final states bit-identical to the oracle
    assert (o.x_final == 20.0).all()


def test_full_size_kepler_4m_checksum_and_invariants():
    g = json.load(open(os.path.join(GOLD, "full_size_checksums.json")))["kepler_dop853_4M_stride64"]
    n = 1 << 22
    y0, p = W.kepler_sweep(n)
    x_end = 5 * W.kepler_period()
    cfg = ob.config_new(ob.DOP853, 0.0, x_end, 60.0, 1e-10, 1e-10, out_type=ob.OUT_SPARSE)
    o = ob.integrate_batch(ob.System.builtin("kepler"), cfg, y0, p)
    assert (o.status == ob.OK).all() and (o.x_final == x_end).all()
    sub = slice(0, n, 64)
    assert (int(o.accepted[sub].sum()), int(o.rejected[sub].sum()), int(o.n_step[sub].sum())) == (
        g["accepted"], g["rejected"], g["n_step"])
    assert sha(o.y_final[:, sub]) == g["sha256_y_final"]
    # size-independent properties on ALL 4M trajectories: the two-body invariants survive 5 periods
    mu = p[0]

    def energy(y):
        return 0.5 * (y[3] ** 2 + y[4] ** 2 + y[5] ** 2) - mu / np.sqrt(y[0] ** 2 + y[1] ** 2 + y[2] ** 2)

    def h_z(y):
        return y[0] * y[4] - y[1] * y[3]

    assert np.max(np.abs(energy(o.y_final) / energy(y0) - 1.0)) < 1e-7
    assert np.max(np.abs(h_z(o.y_final) / h_z(y0) - 1.0)) < 1e-7
    assert (o.y_final[2] == 0.0).all() and (o.y_final[5] == 0.0).all()  # planar orbits stay planar
    # after an integer number of periods the state returns to periapsis
    assert np.max(np.abs(o.y_final[0] / y0[0] - 1.0)) < 1e-3
