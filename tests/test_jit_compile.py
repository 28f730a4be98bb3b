"""User systems (the `System` trait supplied as CUDA source) compile with NVRTC for sm_100a on a
machine without a GPU; a source error comes back with the compiler's message."""
import pytest

import ode_solvers_b200 as ob

LORENZ = ("dy[0] = p[0] * (y[1] - y[0]);\n dy[1] = y[0] * (p[2] - y[2]) - y[1];\n"
          " dy[2] = y[0] * y[1] - p[1] * y[2];")


def test_user_system_compiles_for_every_method_without_a_gpu():
    f = ob.System.from_source("cpu_user_lorenz", 3, 3, LORENZ, "return y[2] > 1.0e3;")
    assert (f.dim, f.n_params, f.name) == (3, 3, "cpu_user_lorenz")
    for method, out in ((ob.RK4, ob.OUT_SPARSE), (ob.DOPRI5, ob.OUT_SPARSE), (ob.DOPRI5, ob.OUT_DENSE),
                        (ob.DOPRI5, ob.OUT_CONTINUOUS), (ob.DOP853, ob.OUT_SPARSE), (ob.DOP853, ob.OUT_DENSE)):
        f.compile(method, out)


def test_compile_error_carries_the_nvrtc_log():
    g = ob.System.from_source("cpu_broken", 1, 0, "dy[0] = undefined_symbol;")
    with pytest.raises(ob.Ode200Error) as e:
        g.compile(ob.DOPRI5)
    assert "undefined_symbol" in str(e.value)


def test_bad_registration_arguments():
    with pytest.raises(ob.Ode200Error):
        ob.System.from_source("too_big", 33, 0, "dy[0] = 0.0;")
    with pytest.raises(ob.Ode200Error):
        ob.System.from_source("too_many_params", 3, 17, "dy[0] = 0.0;")


def test_speculative_arithmetic_macros_and_runtime_dimension_compile_without_a_gpu():
    """div_d / sqrt_d (the branch-free exact division / root of the step loop) in a user RHS, for the
    lockstep Dop853 kernel (dim 6) as well as Dopri5 and Rk4; a DynSystem resolves one compiled system
    per state length and reuses it"""
    kepler = ("const double r = sqrt_d(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);\n const double r3 = r * r * r;\n"
              "dy[0] = y[3]; dy[1] = y[4]; dy[2] = y[5];\n"
              "dy[3] = div_d(-p[0] * y[0], r3); dy[4] = div_d(-p[0] * y[1], r3); dy[5] = div_d(-p[0] * y[2], r3);")
    f = ob.System.from_source("cpu_user_kepler", 6, 1, kepler)
    for method, out in ((ob.RK4, ob.OUT_SPARSE), (ob.DOPRI5, ob.OUT_DENSE), (ob.DOP853, ob.OUT_SPARSE),
                        (ob.DOP853, ob.OUT_CONTINUOUS8)):
        f.compile(method, out)
    dyn = ob.DynSystem("cpu_chain", 1, "for (int i = 0; i < DIM; ++i) dy[i] = div_d(x - y[i], p[0]);", params=[2.0])
    a, b = dyn.resolve(4), dyn.resolve(9)
    assert (a.dim, b.dim) == (4, 9) and dyn.resolve(4) is a and a.name == "cpu_chain_dim4"
    a.compile(ob.DOPRI5)
    b.compile(ob.DOP853)  # n > 8: the one-CTA-per-SM kernel
    with pytest.raises(ValueError):
        dyn.resolve(0)
