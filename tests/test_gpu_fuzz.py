"""Randomised differential test: random solver configurations (from_param space, directions, output
modes, capacities) x random systems, CUDA kernels vs the CPU oracle, everything bit for bit. The
seeds are fixed, so a failure names a reproducible case."""
import numpy as np
import pytest

import ode_solvers_b200 as ob
from ode_solvers_b200 import workloads as W
from parity_util import ENSEMBLES, assert_same_outputs, run_both

pytestmark = pytest.mark.gpu

SPAN = {"lorenz": 3.0, "kepler": 0.6 * W.kepler_period(), "cr3bp": 3.0, "robertson": 0.3, "bouncing_ball": 4.0,
        "threebody18": 1.0}


def random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    system = str(rng.choice(["lorenz", "kepler", "cr3bp", "robertson", "bouncing_ball", "threebody18"],
                            p=[0.25, 0.25, 0.2, 0.1, 0.1, 0.1]))
    method = int(rng.choice([ob.DOPRI5, ob.DOP853]))
    out_type = int(rng.choice([ob.OUT_DENSE, ob.OUT_SPARSE, ob.OUT_CONTINUOUS] +
                              ([ob.OUT_CONTINUOUS8] if method == ob.DOP853 else [])))
    span = SPAN[system] * rng.uniform(0.2, 1.0)
    x0 = float(rng.choice([0.0, 0.0, 0.37 * span]))
    backward = system == "kepler" and out_type != ob.OUT_DENSE and rng.random() < 0.3
    x_end = x0 - span if backward else x0 + span
    samples = int(rng.choice([7, 40, 300, 1500]))
    dx = span / samples * float(rng.choice([1.0, 1.0, 0.999, 1.37]))
    rtol = 10.0 ** rng.uniform(-11, -2)
    atol = rtol * 10.0 ** rng.uniform(-4, 1)
    beta = float(rng.choice([0.0, 0.04, 0.08, 0.1]))
    fac_min = rng.uniform(0.1, 0.5)
    fac_max = rng.uniform(1.5, 10.0)
    h_max = span * float(rng.choice([1.0, 1.0, 0.3, 0.02])) * (-1.0 if rng.random() < 0.2 else 1.0)
    h0 = float(rng.choice([0.0, 0.0, span * 1e-4])) * (-1.0 if backward else 1.0)
    n_max = int(rng.choice([100000, 100000, 40]))
    n_stiff = int(rng.choice([1000, 1, 3, 17]))
    cfg = ob.config_from_param(method, x0, x_end, dx, rtol, atol, rng.uniform(0.7, 0.95), beta, fac_min, fac_max, h_max,
                               h0, n_max, n_stiff, out_type)
    cap = int(rng.choice([samples + 3, samples + 4, 2 * samples + 5, max(3, samples // 2), 2001]))
    n = 64 if system == "threebody18" else int(rng.choice([33, 96, 200]))
    if rng.random() < 0.12:  # the fixed-step method now and then
        step = span / float(rng.choice([10, 333, 2500])) * (1.0 if rng.random() < 0.9 else -1.0)
        cfg = ob.config_rk4(x0, x_end, step)
        cap = int(rng.choice([0, 12, 335, 2502]))
    return system, cfg, n, cap


def perturb_inputs(seed, system, y0, p):
    """now and then: stiff Robertson rates (StiffnessDetected / StepSizeUnderflow paths), non-finite or
    huge initial components, rescaled Kepler units (operands outside the fast division's window)"""
    rng = np.random.default_rng(5000 + seed)
    y0 = y0.copy()
    p = np.array(p, dtype=np.float64)
    r = rng.random()
    if system == "robertson" and r < 0.6:
        p = p * np.array([[1.0], [10.0 ** rng.uniform(2, 6)], [10.0 ** rng.uniform(2, 8)]])
    elif r < 0.08:
        y0[rng.integers(0, y0.shape[0]), ::7] = rng.choice([np.inf, -np.inf, np.nan, 1e308, -1e300, 5e-324, 0.0])
    elif system == "kepler" and r < 0.3:
        s = 10.0 ** rng.uniform(-80, 60)
        p = p.reshape(1, -1) * np.ones((1, y0.shape[1]))
        y0[:, 1::3] *= s
        p[:, 1::3] *= s ** 3
    return y0, p


import os

# ODE_B200_FUZZ_SEEDS="lo:hi" widens the hunt (e.g. 600:6000 runs in about a minute on a B200)
_LO, _HI = (int(v) for v in os.environ.get("ODE_B200_FUZZ_SEEDS", "0:600").split(":"))


@pytest.mark.parametrize("seed", range(_LO, _HI))
def test_random_configuration_matches_oracle_bitwise(seed):
    system, cfg, n, cap = random_case(seed)
    y0, p = ENSEMBLES[system](n, offset=9000 + 50 * seed)
    y0, p = perturb_inputs(seed, system, y0, p)
    com = cap if cfg.out_type in (ob.OUT_CONTINUOUS, ob.OUT_CONTINUOUS8) else 0
    if cfg.method == ob.RK4 or (cfg.method == ob.DOP853 and cfg.out_type == ob.OUT_CONTINUOUS):
        com = 0  # Dop853 treats Continuous like Sparse (dop853.rs:398,540-542)
    ctx = None
    if seed % 5 == 4:  # the static one-thread-per-trajectory schedule must give the same bits
        ctx = ob.Context(0)
        ctx.set_work_queue(False)
    g, o = run_both(system, cfg, y0, p, out_cap=cap, com_cap=com, ctx=ctx)
    if ctx is not None:
        ctx.close()
    label = "fuzz %d: %s m%d o%d x %g->%g dx %g rtol %.2e atol %.2e beta %g n_max %d n_stiff %d cap %d" % (
        seed, system, cfg.method, cfg.out_type, cfg.x0, cfg.x_end, cfg.dx, cfg.rtol, cfg.atol, cfg.beta, cfg.n_max,
        cfg.n_stiff, cap)
    assert_same_outputs(g, o, label)
