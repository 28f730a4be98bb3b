"""The reference's example programs, run literally through the crate-shaped host API on the GPU
(N = 1 batches). Expected Stats / final states are the survey's fixtures,
which tests/test_oracle_golden.py also pins on the CPU oracle."""
import math

import numpy as np
import pytest

import ode_solvers_b200 as ob

pytestmark = pytest.mark.gpu


def test_example_lorenz_literal():
    """examples/lorenz.rs:17-28: Dop853, 0 -> 100, dx = 1e-3, rtol = atol = 1e-4, y0 = (1, 1, 1)"""
    system = ob.System.builtin("lorenz", [10.0, 8.0 / 3.0, 28.0])
    stepper = ob.Dop853.new(system, 0.0, 100.0, 1e-3, [1.0, 1.0, 1.0], 1e-4, 1e-4)
    stats = stepper.integrate()
    assert (stats.num_eval, stats.accepted_steps, stats.rejected_steps) == (15787, 855, 269)
    assert len(stepper.x_out()) == len(stepper.y_out()) == 100001
    assert tuple(stepper.y) == (11.050144919666055, 16.885442582557765, 21.887923123488036)
    assert str(stats) == ("Number of function evaluations: 15787\nNumber of accepted steps: 855\n"
                          "Number of rejected steps: 269")


def test_example_kepler_orbit_literal():
    """examples/kepler_orbit.rs:18-37: Dopri5, 5 periods, dx = 60, rtol = atol = 1e-10"""
    mu, a = 398600.435436, 20000.0
    period = 2.0 * math.pi * math.sqrt(a ** 3 / mu)
    y0 = [-5007.248417988539, -1444.918140151374, 3628.534606178356, 0.717716656891, -10.224093784269,
          0.748229399696]
    stepper = ob.Dopri5.new(ob.System.builtin("kepler", [mu]), 0.0, 5.0 * period, 60.0, y0, 1e-10, 1e-10)
    stats = stepper.integrate()
    assert (stats.num_eval, stats.accepted_steps, stats.rejected_steps) == (8145, 1339, 18)
    assert len(stepper.x_out()) == 2346
    assert tuple(stepper.y) == (-5007.248459071929, -1444.9175767267636, 3628.53456600154, 0.7177162284560319,
                                -10.224093906800825, 0.7482297100304597)


def test_example_three_body_literal():
    """examples/three_body_system.rs:15-26: Dop853, 0 -> 150, dx = 0.002, rtol = atol = 1e-14"""
    stepper = ob.Dop853.new(ob.System.builtin("cr3bp", [0.012300118882173]), 0.0, 150.0, 0.002,
                            [-0.271, -0.42, 0.0, 0.3, -1.0, 0.0], 1e-14, 1e-14)
    stats = stepper.integrate()
    assert (stats.num_eval, stats.accepted_steps, stats.rejected_steps) == (129654, 8474, 231)
    assert len(stepper.x_out()) == 75001
    assert tuple(stepper.y) == (0.38237697246428154, -0.6300497681904966, 0.0, 0.19966199753987168,
                                -0.07418622375983154, 0.0)


def test_example_chemical_reaction_literal():
    """examples/chemical_reaction.rs:10-19: Dop853, 0 -> 0.3, dx = 0.3, rtol = 1e-2, atol = 1e-6"""
    stepper = ob.Dop853.new(ob.System.builtin("robertson", [0.04, 1.0e4, 3.0e7]), 0.0, 0.3, 0.3, [1.0, 0.0, 0.0],
                            1e-2, 1e-6)
    stats = stepper.integrate()
    assert (stats.num_eval, stats.accepted_steps, stats.rejected_steps) == (2098, 103, 50)
    assert len(stepper.x_out()) == 2
    assert tuple(stepper.y) == (0.9886737692266352, 3.4503275486359126e-05, 0.011291727497878943)


def test_example_bouncing_ball_f64():
    """examples/bouncing_ball.rs:24-64 in f64: Rk4 h = 0.01 with solout y[0] < 0, then the host-side
    bounce loop (SolverResult::append / stepper.into())"""
    system = ob.System.builtin("bouncing_ball", [9.81])
    y0 = np.array([10.0, 0.0, 0.0])
    combined = ob.SolverResult()
    bounces = 0
    while bounces < 5 and y0[0] >= 0.001:
        stepper = ob.Rk4.new(system, 0.0, y0, 10.0, 0.01)
        stats = stepper.integrate()
        if bounces == 0:
            assert (stats.num_eval, stats.accepted_steps) == (572, 143)
            assert stepper.x == 1.430000000000001
            assert tuple(stepper.y) == (-0.03023450000000058, -14.02830000000002, 0.0)
        bounces += 1
        last = next(y for y in stepper.y_out() if y[0] <= 0.0)
        y0 = np.array([abs(last[0]), -1.0 * last[1] * 0.75, 0.0])
        combined.append(stepper.into())
    assert bounces == 5 and len(combined.get()[0]) == len(combined.get()[1]) > 400


# ---- the DVector (runtime-dimension) twins of the examples and of the crate's Rk4 tests
KEPLER_DVEC = ("const double r = sqrt_d(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);\n"
               "const double r3 = r * r * r;\n"  # r.powi(3) = (r * r) * r
               "dy[0] = y[3]; dy[1] = y[4]; dy[2] = y[5];\n"
               "dy[3] = div_d(-p[0] * y[0], r3); dy[4] = div_d(-p[0] * y[1], r3); dy[5] = div_d(-p[0] * y[2], r3);")

CR3BP_DVEC = ("const double mu = p[0];\n"
              "const double a = y[0] + mu, b = y[0] - 1.0 + mu;\n"
              "const double d = sqrt_d(a * a + y[1] * y[1] + y[2] * y[2]);\n"
              "const double r = sqrt_d(b * b + y[1] * y[1] + y[2] * y[2]);\n"
              "const double d3 = d * d * d, r3 = r * r * r;\n"
              "dy[0] = y[3]; dy[1] = y[4]; dy[2] = y[5];\n"
              "dy[3] = y[0] + 2.0 * y[4] - div_d((1.0 - mu) * (y[0] + mu), d3) - div_d(mu * (y[0] - 1.0 + mu), r3);\n"
              "dy[4] = -2.0 * y[3] + y[1] - div_d((1.0 - mu) * y[1], d3) - div_d(mu * y[1], r3);\n"
              "dy[5] = div_d(-(1.0 - mu) * y[2], d3) - div_d(mu * y[2], r3);")


def test_example_kepler_orbit_dvector_literal(tmp_path):
    """examples/kepler_orbit_dvector.rs:18-47: `type State = DVector<f64>`; same numbers as the
    SVector example, and the example's save() file"""
    mu, a = 398600.435436, 20000.0
    period = 2.0 * math.pi * math.sqrt(a ** 3 / mu)
    y0 = [-5007.248417988539, -1444.918140151374, 3628.534606178356, 0.717716656891, -10.224093784269,
          0.748229399696]
    system = ob.DynSystem("kepler_orbit_dvector", 1, KEPLER_DVEC, params=[mu])
    stepper = ob.Dopri5.new(system, 0.0, 5.0 * period, 60.0, y0, 1e-10, 1e-10)
    stats = stepper.integrate()
    assert (stats.num_eval, stats.accepted_steps, stats.rejected_steps) == (8145, 1339, 18)
    assert tuple(stepper.y) == (-5007.248459071929, -1444.9175767267636, 3628.53456600154, 0.7177162284560319,
                                -10.224093906800825, 0.7482297100304597)
    path = tmp_path / "outputs" / "kepler_orbit_dopri5_dvector.dat"
    ob.save(stepper.x_out(), stepper.y_out(), str(path))
    lines = path.read_text().splitlines()
    assert len(lines) == 2346
    assert lines[0] == "0, -5007.248417988539, -1444.918140151374, 3628.534606178356, 0.717716656891, " \
                       "-10.224093784269, 0.748229399696"
    assert lines[1].startswith("60, ")
    back = np.array([[float(v) for v in ln.split(", ")] for ln in lines])
    assert (back[:, 0] == np.array(stepper.x_out())).all() and (back[:, 1:] == np.array(stepper.y_out())).all()


def test_example_three_body_dvector_literal():
    """examples/three_body_system_dvector.rs:15-37"""
    system = ob.DynSystem("three_body_dvector", 1, CR3BP_DVEC, params=[0.012300118882173])
    stepper = ob.Dop853.new(system, 0.0, 150.0, 0.002, [-0.271, -0.42, 0.0, 0.3, -1.0, 0.0], 1e-14, 1e-14)
    stats = stepper.integrate()
    assert (stats.num_eval, stats.accepted_steps, stats.rejected_steps) == (129654, 8474, 231)
    assert len(stepper.x_out()) == 75001
    assert tuple(stepper.y) == (0.38237697246428154, -0.6300497681904966, 0.0, 0.19966199753987168,
                                -0.07418622375983154, 0.0)


def test_runtime_dimension_is_taken_from_the_state():
    """one DynSystem, states of different lengths (src/rk4.rs:272-306 runs its tests on DVector
    states): n decoupled copies of Test1 y' = (x - y)/2 reproduce the crate's golden values
    component by component, batched and through the stepper"""
    system = ob.DynSystem("test1_dvector", 0, "for (int i = 0; i < DIM; ++i) dy[i] = (x - y[i]) / 2.0;")
    for n in (1, 2, 5):
        stepper = ob.Rk4.new(system, 0.0, [1.0] * n, 0.2, 0.1)
        stepper.integrate()
        out = stepper.y_out()
        assert len(out) == 3 and len(out[1]) == n
        assert all(abs(v - 0.95369) < 1e-5 for v in out[1]) and all(abs(v - 0.91451) < 1e-5 for v in out[2])
    y0 = np.ones((4, 10))
    g = ob.integrate_batch(system, ob.config_rk4(0.0, 0.2, 0.1), y0, out_cap=4)
    assert g.y_final.shape == (4, 10) and np.abs(g.y_final - 0.91451).max() < 1e-5
    d5 = ob.Dopri5.new(system, 0.0, 1.0, 0.5, [1.0, 1.0, 1.0], 1e-8, 1e-8)
    d5.integrate()
    exact = 3.0 * math.exp(-0.5) - 2.0 + 1.0  # y = 3 e^{-x/2} - 2 + x
    assert all(abs(v - exact) < 1e-7 for v in d5.y)
