"""The f32 instantiation of the oracle (oracle/ode_oracle_core.inc included for T = float) and the
f32 constructors: CPU tier. The crate is generic over FloatNumber = {f32, f64} (src/dop_shared.rs:52-77);
its tests hold no f32 value, so the f32 oracle is pinned on (i) being the SAME source as the f64 oracle
(which is pinned on the crate's goldens), (ii) the constructors' literal semantics (T::from of an f64
literal = the value rounded to f32; dopri5.rs:89 vs :160 = quirk Q3), and (iii) agreement with the f64
run of the same problem to f32 accuracy."""
import ctypes as C

import numpy as np

import oracle
import ode_solvers_b200 as ob
from ode_solvers_b200 import _abi, workloads as W


def _fields(c):
    return [getattr(c, f) for f, _ in type(c)._fields_]


def test_f32_constructors_match_the_oracle_and_the_rust_semantics():
    f32 = np.float32
    for m in (ob.DOPRI5, ob.DOP853):
        a, b = ob.config_new_f32(m, 0.0, 10.0, 0.01, 1e-2, 1e-6), oracle.config_new_f32(m, 0.0, 10.0, 0.01, 1e-2, 1e-6)
        assert _fields(a) == _fields(b)
        args = (m, 1.0, -3.0, 0.25, 1e-4, 1e-5, 0.85, 0.08, 0.25, 8.0, 2.0, 1e-3, 5000, 77, ob.OUT_SPARSE)
        assert _fields(ob.config_from_param_f32(*args)) == _fields(oracle.config_from_param_f32(*args))
    c = ob.config_new_f32(ob.DOPRI5, 0.0, 10.0, 0.01, 1e-2, 1e-6)
    # dopri5.rs:14-27 with T = f32: T::from(0.2 - 0.04 * 0.75) rounds the f64 VALUE; uround = f32::EPSILON (:89)
    assert f32(c.alpha) == f32(0.2 - 0.04 * 0.75) and f32(c.beta) == f32(0.04) and f32(c.fac_min) == f32(0.2)
    assert f32(c.uround) == np.finfo(np.float32).eps and c.h_max == 10.0 and c.n_max == 100000
    d = ob.config_from_param_f32(ob.DOPRI5, 0.0, 10.0, 0.01, 1e-2, 1e-6, 0.9, 0.04, 0.2, 10.0, 10.0, 0.0, 100000, 1000,
                                 ob.OUT_DENSE)
    # from_param computes alpha IN f32 (0.2f32 - beta * 0.75f32, dopri5.rs:147) and takes uround = f64::EPSILON (:160): Q3
    assert f32(d.alpha) == f32(0.2) - f32(0.04) * f32(0.75)
    assert f32(d.uround) == f32(2.220446049250313e-16) and d.uround != c.uround
    e = ob.config_new_f32(ob.DOP853, 0.0, 1.0, 0.1, 1e-4, 1e-4)
    assert (f32(e.alpha), e.beta, f32(e.fac_min), f32(e.uround)) == (f32(0.125), 0.0, f32(0.333), f32(2.220446049250313e-16))
    assert _fields(ob.config_rk4_f32(0.0, 2.0, 0.01)) == _fields(oracle.config_rk4_f32(0.0, 2.0, 0.01))


def test_f32_oracle_tracks_the_f64_oracle_to_single_precision():
    y0, p = W.lorenz_ensemble(40, offset=3)
    for m in (ob.DOPRI5, ob.DOP853):
        c64 = oracle.config_new(m, 0.0, 0.5, 0.05, 1e-5, 1e-5)
        c32 = oracle.config_new_f32(m, 0.0, 0.5, 0.05, 1e-5, 1e-5)
        a = oracle.integrate_batch("lorenz", c64, y0, p, out_cap=14)
        b = oracle.integrate_batch_f32("lorenz", c32, y0, p, out_cap=14)
        # xd += dx accumulates in f32: the sample at x_end itself can fall just outside (the endpoint rule's 1e-9
        # is below f32 resolution), so the f32 run may record one sample less -- the reference would too
        assert (b.status == ob.OK).all() and (np.abs(b.out_count.astype(int) - a.out_count.astype(int)) <= 1).all()
        assert np.abs(b.y_final - a.y_final).max() < 5e-3 * max(1.0, np.abs(a.y_final).max())
        assert np.abs(b.out_y[:, :10] - a.out_y[:, :10]).max() < 5e-3 * max(1.0, np.abs(a.out_y).max())
        assert (np.abs(b.accepted.astype(int) - a.accepted.astype(int)) <= 3).all()
    # the f32 run of rk4.rs Test1 (rk4.rs:223-232) reproduces the crate's table to its 1e-5
    b = oracle.integrate_batch_f32("test_rk4_1", oracle.config_rk4_f32(0.0, 0.2, 0.1), np.ones((1, 1)), None, out_cap=4)
    assert abs(b.out_y[0, 1, 0] - 0.95369) < 1e-5 and abs(b.out_y[0, 2, 0] - 0.91451) < 1e-5


def test_f32_dopri5_bouncing_ball_alternative_of_the_example():
    """examples/bouncing_ball.rs:35 (commented alternative): Dopri5::new(system, 0., 10.0, 0.01, y0, 1.0e-2, 1.0e-6)
    with State = Vector3<f32>: Dense output, solout stops below the floor"""
    cfg = oracle.config_new_f32(ob.DOPRI5, 0.0, 10.0, 0.01, 1e-2, 1e-6)
    o = oracle.integrate_batch_f32("bouncing_ball", cfg, np.array([[10.0], [0.0], [0.0]]), [9.81], out_cap=1100)
    # the ball reaches the floor at sqrt(2 * 10 / 9.81) = 1.43 s; solout is only asked at the end of an accepted step, so
    # the stepper overshoots ("the results contain a lot of invalid data with y[0] < 0", bouncing_ball.rs:62)
    assert o.status[0] == ob.OK and o.y_final[0, 0] < 0.0 and 1.43 < o.x_final[0] < 10.0
    n = int(o.out_count[0])
    assert n >= 143 and np.allclose(np.diff(o.out_x[0, :n]), 0.01, atol=1e-5)
    first_below = int(np.argmax(o.out_y[0, :n, 0] <= 0.0))  # the host bounce loop's `find(|y| y[0] <= 0.)`
    assert abs(o.out_x[0, first_below] - 1.43) < 0.011
    # Q3 matters in f32: `new` (uround = f32::EPSILON) and `from_param` (f64::EPSILON) are different steppers
    stiff = oracle.config_new_f32(ob.DOPRI5, 0.0, 0.3, 0.3, 1e-2, 1e-6, out_type=ob.OUT_SPARSE)
    same = oracle.config_from_param_f32(ob.DOPRI5, 0.0, 0.3, 0.3, 1e-2, 1e-6, 0.9, 0.04, 0.2, 10.0, 0.3, 0.0, 100000, 1000,
                                        ob.OUT_SPARSE)
    assert stiff.uround != same.uround


def test_f32_powf_clone_is_what_the_oracle_calls():
    lib = oracle.lib()
    chk = C.CDLL(oracle.ORACLE_LIB.replace("libode_oracle.so", "libode_powcheck.so"))
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(1e-8, 4.0, 2000), [1e-4, 1.0, 0.5, 1e-30, 3e38]]).astype(np.float32)
    ys = np.concatenate([rng.choice([0.17, -0.04, 0.125, 0.2, -0.08], 2000), [0.17, 0.2, -0.04, 0.125, 0.17]]).astype(np.float32)
    if hasattr(chk, "ode_check_powf"):
        chk.ode_check_powf.argtypes = [C.c_float, C.c_float]
        chk.ode_check_powf.restype = C.c_float
        for x, y in zip(xs, ys):
            a, b = lib.ode_oracle_powf(float(x), float(y)), chk.ode_check_powf(float(x), float(y))
            assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32), (x, y)
