"""Independent cross-check of the CPU oracle's numerics for the systems the crate's own tests hold no
golden values for (Kepler, CR3BP, Robertson, Newtonian 3-body-18, Lorenz): scipy's DOP853 / RK45 are
different implementations of the same Dormand-Prince pairs (other controller, other dense output), so
agreement is to integration tolerance only -- it shows that the restated tableaux, right-hand sides
and step loops solve the right equations; bit-level pinning is tests/test_oracle_golden.py's job."""
import math

import numpy as np
import pytest

scipy_integrate = pytest.importorskip("scipy.integrate")

import ode_solvers_b200 as ob  # noqa: E402
import oracle  # noqa: E402
from ode_solvers_b200 import workloads as W  # noqa: E402


def lorenz(t, y, s=10.0, b=8.0 / 3.0, r=28.0):
    return [s * (y[1] - y[0]), y[0] * (r - y[2]) - y[1], y[0] * y[1] - b * y[2]]


def kepler(t, y, mu=W.KEPLER_MU):
    r3 = math.sqrt(y[0] ** 2 + y[1] ** 2 + y[2] ** 2) ** 3
    return [y[3], y[4], y[5], -mu * y[0] / r3, -mu * y[1] / r3, -mu * y[2] / r3]


def cr3bp(t, y, mu=W.CR3BP_MU):
    d3 = math.sqrt((y[0] + mu) ** 2 + y[1] ** 2 + y[2] ** 2) ** 3
    r3 = math.sqrt((y[0] - 1 + mu) ** 2 + y[1] ** 2 + y[2] ** 2) ** 3
    return [y[3], y[4], y[5],
            y[0] + 2 * y[4] - (1 - mu) * (y[0] + mu) / d3 - mu * (y[0] - 1 + mu) / r3,
            -2 * y[3] + y[1] - (1 - mu) * y[1] / d3 - mu * y[1] / r3,
            -(1 - mu) * y[2] / d3 - mu * y[2] / r3]


def robertson(t, y, k1=0.04, k2=1e4, k3=3e7):
    return [-k1 * y[0] + k2 * y[1] * y[2], k1 * y[0] - k2 * y[1] * y[2] - k3 * y[1] ** 2, k3 * y[1] ** 2]


def threebody18(t, y, gm):
    dy = np.zeros(18)
    dy[:9] = y[9:]
    for a in range(3):
        for b in range(3):
            if a != b:
                d = y[3 * b:3 * b + 3] - y[3 * a:3 * a + 3]
                dy[9 + 3 * a:12 + 3 * a] += gm[b] * d / np.linalg.norm(d) ** 3
    return dy


CASES = [
    ("lorenz", lorenz, np.array([1.0, 1.0, 1.0]), W.LORENZ_PARAMS, 3.0, ()),
    ("kepler", kepler, np.array([-5007.248417988539, -1444.918140151374, 3628.534606178356, 0.717716656891,
                                 -10.224093784269, 0.748229399696]), np.array([W.KEPLER_MU]), 20000.0, ()),
    ("cr3bp", cr3bp, np.array(W.CR3BP_Y0, dtype=float), np.array([W.CR3BP_MU]), 5.0, ()),
    ("robertson", robertson, np.array([1.0, 0.0, 0.0]), np.array([0.04, 1e4, 3e7]), 0.3, ()),
]


@pytest.mark.parametrize("method,scipy_method", [(ob.DOPRI5, "RK45"), (ob.DOP853, "DOP853")])
@pytest.mark.parametrize("name,f,y0,p,x_end,args", CASES)
def test_oracle_agrees_with_scipy_to_tolerance(name, f, y0, p, x_end, args, method, scipy_method):
    tol = 1e-11 if name != "robertson" else 1e-9
    cfg = oracle.config_new(method, 0.0, x_end, x_end, tol, tol * 1e-3, out_type=ob.OUT_SPARSE)
    o = oracle.integrate_batch(name, cfg, y0.reshape(-1, 1), p)
    assert o.status[0] == ob.OK
    ref = scipy_integrate.solve_ivp(f, (0.0, x_end), y0, method="DOP853", rtol=1e-13, atol=1e-16, args=args)
    assert ref.success
    scale = np.maximum(np.abs(ref.y[:, -1]), np.abs(y0).max() * 1e-3)
    err = np.abs(o.y_final[:, 0] - ref.y[:, -1]) / scale
    assert err.max() < (2e-7 if name == "lorenz" else 2e-8), (name, scipy_method, err)


def test_oracle_threebody18_agrees_with_scipy():
    y0, gm = W.threebody18_ensemble(3, offset=4)
    gm = np.asarray(gm, dtype=float)
    gm0 = gm[:, 0] if gm.ndim == 2 else gm
    for method in (ob.DOPRI5, ob.DOP853):
        cfg = oracle.config_new(method, 0.0, 2.0, 2.0, 1e-11, 1e-14, out_type=ob.OUT_SPARSE)
        o = oracle.integrate_batch("threebody18", cfg, y0, gm)
        ref = scipy_integrate.solve_ivp(threebody18, (0.0, 2.0), y0[:, 0], method="DOP853", rtol=1e-13, atol=1e-16,
                                        args=(gm0,))
        assert o.status[0] == ob.OK and ref.success
        assert np.abs(o.y_final[:, 0] - ref.y[:, -1]).max() < 1e-7 * max(1.0, np.abs(ref.y[:, -1]).max())
