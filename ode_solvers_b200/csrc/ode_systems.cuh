/*
 * ode_systems.cuh -- built-in `System` implementations as CUDA device functors.
 *
 * The reference's extension point is `trait System<T,V>` (src/dop_shared.rs:31-42):
 *   fn system(&self, x, &y, &mut dy)        -> rhs(x, y, dy, p)
 *   fn solout(&mut self, x, &y, &dy) -> bool -> solout(x, y, dy, p)   (default: false)
 * with the implementing struct's fields becoming the per-trajectory parameter row p[NP].
 * Each functor below restates one example / test system of the reference with the SAME
 * operation order (powi(2) = v*v, powi(3) = (v*v)*v, left-associated sums), so that with
 * -fmad=false the device produces the bits the crate produces on the host.
 * rhs must write every component of dy.
 *
 * Systems whose rhs is a template on the scalar type T (HAS_F32) also instantiate for f32, the
 * reference's other FloatNumber (src/dop_shared.rs:74; examples/bouncing_ball.rs:15 runs in f32).
 */
#pragma once

#include "ode_device.cuh"

namespace ode_b200 {

/* examples/lorenz.rs:42-54 ; p = {sigma, beta, rho} */
struct SysLorenz {
    static constexpr int DIM = 3, NP = 3;
    static constexpr bool SPEC = false;
    static constexpr bool HAS_SOLOUT = false, HAS_F32 = true;
    template <class T, class M>
    ODE_DEV static void rhs(T, const T* y, T* dy, const T* p, M&) {
        dy[0] = p[0] * (y[1] - y[0]);
        dy[1] = y[0] * (p[2] - y[2]) - y[1];
        dy[2] = y[0] * y[1] - p[1] * y[2];
    }
    template <class T>
    ODE_DEV static bool solout(T, const T*, const T*, const T*) { return false; }
};

/* examples/kepler_orbit.rs:51-66 ; p = {mu} */
struct SysKepler {
    static constexpr int DIM = 6, NP = 1;
    static constexpr bool SPEC = true;
    static constexpr bool HAS_SOLOUT = false, HAS_F32 = false;
    template <class M>
    ODE_DEV static void rhs(double, const double* y, double* dy, const double* p, M& m) {
        const double r = m.sqrt_nz(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
        const double r3 = r * r * r;
        /* three exact IEEE quotients by the same r3: one refined reciprocal (ode_device.cuh) */
        const double ir3 = m.rcp(r3);
        dy[0] = y[3];
        dy[1] = y[4];
        dy[2] = y[5];
        dy[3] = m.div_by(-p[0] * y[0], r3, ir3);
        dy[4] = m.div_by(-p[0] * y[1], r3, ir3);
        dy[5] = m.div_by(-p[0] * y[2], r3, ir3);
    }
    ODE_DEV static bool solout(double, const double*, const double*, const double*) { return false; }
};

/* examples/three_body_system.rs:40-59 (circular restricted three-body problem) ; p = {mu} */
struct SysCr3bp {
    static constexpr int DIM = 6, NP = 1;
    static constexpr bool SPEC = true;
    static constexpr bool HAS_SOLOUT = false, HAS_F32 = false;
    template <class M>
    ODE_DEV static void rhs(double, const double* y, double* dy, const double* p, M& m) {
        const double mu = p[0];
        const double a = y[0] + mu, b = y[0] - 1.0 + mu;
        const double d = m.sqrt_nz(a * a + y[1] * y[1] + y[2] * y[2]);
        const double r = m.sqrt_nz(b * b + y[1] * y[1] + y[2] * y[2]);
        const double d3 = d * d * d, r3 = r * r * r;
        const double id3 = m.rcp(d3), ir3 = m.rcp(r3); /* one refined reciprocal per denominator */
        dy[0] = y[3];
        dy[1] = y[4];
        dy[2] = y[5];
        dy[3] = y[0] + 2.0 * y[4] - m.div_by((1.0 - mu) * (y[0] + mu), d3, id3) -
                m.div_by(mu * (y[0] - 1.0 + mu), r3, ir3);
        dy[4] = -2.0 * y[3] + y[1] - m.div_by((1.0 - mu) * y[1], d3, id3) - m.div_by(mu * y[1], r3, ir3);
        dy[5] = m.div_by(-(1.0 - mu) * y[2], d3, id3) - m.div_by(mu * y[2], r3, ir3);
    }
    ODE_DEV static bool solout(double, const double*, const double*, const double*) { return false; }
};

/* examples/chemical_reaction.rs:28-36 (Robertson) with the literals lifted into
 * p = {k1, k2, k3}; p = (0.04, 1e4, 3e7) is bitwise the example */
struct SysRobertson {
    static constexpr int DIM = 3, NP = 3;
    static constexpr bool SPEC = false;
    static constexpr bool HAS_SOLOUT = false, HAS_F32 = true;
    template <class T, class M>
    ODE_DEV static void rhs(T, const T* y, T* dy, const T* p, M&) {
        dy[0] = -p[0] * y[0] + p[1] * y[1] * y[2];
        dy[1] = p[0] * y[0] - p[1] * y[1] * y[2] - p[2] * y[1] * y[1];
        dy[2] = p[2] * y[1] * y[1];
    }
    template <class T>
    ODE_DEV static bool solout(T, const T*, const T*, const T*) { return false; }
};

/* examples/bouncing_ball.rs:74-86 in f64 ; p = {g}. dy[2] = 0 written explicitly (Q13). */
struct SysBouncingBall {
    static constexpr int DIM = 3, NP = 1;
    static constexpr bool SPEC = false;
    static constexpr bool HAS_SOLOUT = true, HAS_F32 = true;
    template <class T, class M>
    ODE_DEV static void rhs(T, const T* y, T* dy, const T* p, M&) {
        dy[0] = y[1];
        dy[1] = -p[0];
        dy[2] = T(0);
    }
    template <class T>
    ODE_DEV static bool solout(T, const T* y, const T*, const T*) { return y[0] < T(0); }
};

/* Newtonian three-body problem, 18 states (not in the reference; order defined by the oracle,
 * oracle/ode_oracle_systems.c threebody18_rhs) ; p = {G*m1, G*m2, G*m3} */
struct SysThreeBody18 {
    static constexpr int DIM = 18, NP = 3;
    static constexpr bool SPEC = true;
    static constexpr bool HAS_SOLOUT = false, HAS_F32 = false;
    template <class M>
    ODE_DEV static void rhs(double, const double* y, double* dy, const double* p, M& m) {
#pragma unroll
        for (int c = 0; c < 9; ++c) dy[c] = y[9 + c];
#pragma unroll
        for (int c = 0; c < 9; ++c) dy[9 + c] = 0.0;
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = a + 1; b < 3; ++b) {
                const double dx = y[3 * b] - y[3 * a], dyy = y[3 * b + 1] - y[3 * a + 1],
                             dz = y[3 * b + 2] - y[3 * a + 2];
                const double r2 = dx * dx + dyy * dyy + dz * dz;
                const double r = m.sqrt_nz(r2);
                const double inv = m.div_nz(1.0, r2 * r);
                const double fa = p[b] * inv, fb = p[a] * inv;
                dy[9 + 3 * a] = dy[9 + 3 * a] + fa * dx;
                dy[9 + 3 * a + 1] = dy[9 + 3 * a + 1] + fa * dyy;
                dy[9 + 3 * a + 2] = dy[9 + 3 * a + 2] + fa * dz;
                dy[9 + 3 * b] = dy[9 + 3 * b] - fb * dx;
                dy[9 + 3 * b + 1] = dy[9 + 3 * b + 1] - fb * dyy;
                dy[9 + 3 * b + 2] = dy[9 + 3 * b + 2] - fb * dz;
            }
    }
    ODE_DEV static bool solout(double, const double*, const double*, const double*) { return false; }
};

/* src/rk4.rs:178-186 Test1 */
struct SysTestRk4_1 {
    static constexpr int DIM = 1, NP = 0;
    static constexpr bool SPEC = false;
    static constexpr bool HAS_SOLOUT = false, HAS_F32 = true;
    template <class T, class M>
    ODE_DEV static void rhs(T x, const T* y, T* dy, const T*, M&) { dy[0] = (x - y[0]) / T(2); }
    template <class T>
    ODE_DEV static bool solout(T, const T*, const T*, const T*) { return false; }
};

/* src/rk4.rs:188-196 Test2 */
struct SysTestRk4_2 {
    static constexpr int DIM = 1, NP = 0;
    static constexpr bool SPEC = false;
    static constexpr bool HAS_SOLOUT = false, HAS_F32 = true;
    template <class T, class M>
    ODE_DEV static void rhs(T x, const T* y, T* dy, const T*, M&) { dy[0] = T(-2) * x - y[0]; }
    template <class T>
    ODE_DEV static bool solout(T, const T*, const T*, const T*) { return false; }
};

/* src/rk4.rs:198-206 Test3 (= dopri5.rs:531-538, dop853.rs:595-602): uses f64::exp -> exp_d */
struct SysTestExp {
    static constexpr int DIM = 1, NP = 0;
    static constexpr bool SPEC = false;
    static constexpr bool HAS_SOLOUT = false, HAS_F32 = false;
    template <class M>
    ODE_DEV static void rhs(double x, const double* y, double* dy, const double*, M&) {
        dy[0] = (5.0 * x * x - y[0]) / exp_d(x + y[0]);
    }
    ODE_DEV static bool solout(double, const double*, const double*, const double*) { return false; }
};

/* the same with `solout: x >= 0.5` (rk4.rs:219-221, dopri5.rs:540-542, dop853.rs:604-606) ; p = {x_stop} */
struct SysTestExpStop {
    static constexpr int DIM = 1, NP = 1;
    static constexpr bool SPEC = false;
    static constexpr bool HAS_SOLOUT = true, HAS_F32 = false;
    template <class M>
    ODE_DEV static void rhs(double x, const double* y, double* dy, const double*, M&) {
        dy[0] = (5.0 * x * x - y[0]) / exp_d(x + y[0]);
    }
    ODE_DEV static bool solout(double x, const double*, const double*, const double* p) { return x >= p[0]; }
};

/* tests/dopri5.rs:13-19 : y' = -3 t^2 e^y */
struct SysTestCubic {
    static constexpr int DIM = 1, NP = 0;
    static constexpr bool SPEC = false;
    static constexpr bool HAS_SOLOUT = false, HAS_F32 = false;
    template <class M>
    ODE_DEV static void rhs(double x, const double* y, double* dy, const double*, M&) {
        dy[0] = -3.0 * x * x * exp_d(y[0]);
    }
    ODE_DEV static bool solout(double, const double*, const double*, const double*) { return false; }
};

/* tests/dopri5.rs:21-31 with `solout: x >= 4.1` ; p = {x_stop} */
struct SysTestCubicStop {
    static constexpr int DIM = 1, NP = 1;
    static constexpr bool SPEC = false;
    static constexpr bool HAS_SOLOUT = true, HAS_F32 = false;
    template <class M>
    ODE_DEV static void rhs(double x, const double* y, double* dy, const double*, M&) {
        dy[0] = -3.0 * x * x * exp_d(y[0]);
    }
    ODE_DEV static bool solout(double x, const double*, const double*, const double* p) { return x >= p[0]; }
};

}  // namespace ode_b200
