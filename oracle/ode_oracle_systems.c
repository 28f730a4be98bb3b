/*
 * ode_oracle_systems.c -- CPU ORACLE (test infrastructure only): the `System` impls of the
 * reference's examples and tests, restated with the same operation order.
 * `powi(2)` = v*v, `powi(3)` = (v*v)*v (LLVM expands constant powi; compiler-rt's
 * __powidf2 gives v*(v*v), bitwise the same product).
 */
#include "ode_oracle.h"

#include <math.h>
#include <string.h>

/* examples/lorenz.rs:48-54 ; p = {sigma, beta, rho} (struct order :42-46) */
static void lorenz_rhs(double x, const double* y, double* dy, const double* p) {
    (void)x;
    dy[0] = p[0] * (y[1] - y[0]);
    dy[1] = y[0] * (p[2] - y[2]) - y[1];
    dy[2] = y[0] * y[1] - p[1] * y[2];
}

/* examples/kepler_orbit.rs:55-66 ; p = {mu} */
static void kepler_rhs(double x, const double* y, double* dy, const double* p) {
    (void)x;
    double r = sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
    double r3 = r * r * r;
    dy[0] = y[3];
    dy[1] = y[4];
    dy[2] = y[5];
    dy[3] = -p[0] * y[0] / r3;
    dy[4] = -p[0] * y[1] / r3;
    dy[5] = -p[0] * y[2] / r3;
}

/* examples/three_body_system.rs:44-59 (circular restricted 3-body) ; p = {mu} */
static void cr3bp_rhs(double x, const double* y, double* dy, const double* p) {
    (void)x;
    double mu = p[0];
    double a = y[0] + mu, b = y[0] - 1.0 + mu;
    double d = sqrt(a * a + y[1] * y[1] + y[2] * y[2]);
    double r = sqrt(b * b + y[1] * y[1] + y[2] * y[2]);
    double d3 = d * d * d, r3 = r * r * r;
    dy[0] = y[3];
    dy[1] = y[4];
    dy[2] = y[5];
    dy[3] = y[0] + 2.0 * y[4] - (1.0 - mu) * (y[0] + mu) / d3 - mu * (y[0] - 1.0 + mu) / r3;
    dy[4] = -2.0 * y[3] + y[1] - (1.0 - mu) * y[1] / d3 - mu * y[1] / r3;
    dy[5] = -(1.0 - mu) * y[2] / d3 - mu * y[2] / r3;
}

/* examples/chemical_reaction.rs:30-36 with the hard-coded rate constants lifted into
 * p = {k1 = 0.04, k2 = 1e4, k3 = 3e7}: `3.*10_f64.powi(7)` is exactly 3e7, and
 * (-0.04)*y0 == -(0.04*y0), so p = (0.04, 1e4, 3e7) is bitwise the example. */
static void robertson_rhs(double x, const double* y, double* dy, const double* p) {
    (void)x;
    dy[0] = -p[0] * y[0] + p[1] * y[1] * y[2];
    dy[1] = p[0] * y[0] - p[1] * y[1] * y[2] - p[2] * y[1] * y[1];
    dy[2] = p[2] * y[1] * y[1];
}

/* examples/bouncing_ball.rs:76-86 in f64 ; p = {g = 9.81}. The example never writes dy[2];
 * in the crate that slot keeps whatever the k-vector held before (zeros for Rk4/Dopri5 and,
 * with y0[2] = 0 as in the example, for Dop853 too). The batched ABI requires rhs to write
 * every component, so dy[2] = 0 is written explicitly. */
static void ball_rhs(double x, const double* y, double* dy, const double* p) {
    (void)x;
    dy[0] = y[1];
    dy[1] = -p[0];
    dy[2] = 0.0;
}
static int ball_solout(double x, const double* y, const double* dy, const double* p) {
    (void)x; (void)dy; (void)p;
    return y[0] < 0.0;
}

/* src/rk4.rs:178-186 Test1 */
static void test_rk4_1_rhs(double x, const double* y, double* dy, const double* p) {
    (void)p;
    dy[0] = (x - y[0]) / 2.0;
}
/* src/rk4.rs:188-196 Test2 */
static void test_rk4_2_rhs(double x, const double* y, double* dy, const double* p) {
    (void)p;
    dy[0] = -2.0 * x - y[0];
}
/* src/rk4.rs:198-206 Test3 = src/dopri5.rs:531-538 = src/dop853.rs:595-602 */
static void test_exp_rhs(double x, const double* y, double* dy, const double* p) {
    (void)p;
    dy[0] = (5.0 * x * x - y[0]) / exp(x + y[0]);
}
/* solout of rk4.rs:219-221, dopri5.rs:540-542, dop853.rs:604-606 ; p = {x_stop = 0.5} */
static int stop_at_x(double x, const double* y, const double* dy, const double* p) {
    (void)y; (void)dy;
    return x >= p[0];
}
/* tests/dopri5.rs:15-19 : y' = -3 t^2 e^y ; stopping variant :23-31 with p = {4.1} */
static void test_cubic_rhs(double x, const double* y, double* dy, const double* p) {
    (void)p;
    dy[0] = -3.0 * x * x * exp(y[0]);
}

/*
 * Newtonian three-body problem, SVector<f64,18>: y = (r1,r2,r3,v1,v2,v3), p = {G*m1,G*m2,G*m3}.
 * NOT in the reference (its "three body" is the 6-dim restricted problem above); BASELINE
 * config 4 asks for it, so the operation order below is DEFINED here and mirrored by the
 * device functor (ode_systems.cuh).
 */
static void threebody18_rhs(double x, const double* y, double* dy, const double* p) {
    (void)x;
    for (int c = 0; c < 9; ++c) dy[c] = y[9 + c];
    for (int c = 0; c < 9; ++c) dy[9 + c] = 0.0;
    for (int a = 0; a < 3; ++a)
        for (int b = a + 1; b < 3; ++b) {
            double dx = y[3 * b] - y[3 * a], dyy = y[3 * b + 1] - y[3 * a + 1], dz = y[3 * b + 2] - y[3 * a + 2];
            double r2 = dx * dx + dyy * dyy + dz * dz;
            double r = sqrt(r2);
            double inv = 1.0 / (r2 * r);
            double fa = p[b] * inv, fb = p[a] * inv;
            dy[9 + 3 * a] = dy[9 + 3 * a] + fa * dx;
            dy[9 + 3 * a + 1] = dy[9 + 3 * a + 1] + fa * dyy;
            dy[9 + 3 * a + 2] = dy[9 + 3 * a + 2] + fa * dz;
            dy[9 + 3 * b] = dy[9 + 3 * b] - fb * dx;
            dy[9 + 3 * b + 1] = dy[9 + 3 * b + 1] - fb * dyy;
            dy[9 + 3 * b + 2] = dy[9 + 3 * b + 2] - fb * dz;
        }
}

static const ode_oracle_system SYSTEMS[] = {
    {"lorenz", 3, 3, lorenz_rhs, 0},
    {"kepler", 6, 1, kepler_rhs, 0},
    {"cr3bp", 6, 1, cr3bp_rhs, 0},
    {"robertson", 3, 3, robertson_rhs, 0},
    {"bouncing_ball", 3, 1, ball_rhs, ball_solout},
    {"threebody18", 18, 3, threebody18_rhs, 0},
    {"test_rk4_1", 1, 0, test_rk4_1_rhs, 0},
    {"test_rk4_2", 1, 0, test_rk4_2_rhs, 0},
    {"test_exp", 1, 0, test_exp_rhs, 0},
    {"test_exp_stop", 1, 1, test_exp_rhs, stop_at_x},
    {"test_cubic", 1, 0, test_cubic_rhs, 0},
    {"test_cubic_stop", 1, 1, test_cubic_rhs, stop_at_x},
};

int ode_oracle_system_count(void) { return (int)(sizeof(SYSTEMS) / sizeof(SYSTEMS[0])); }
const ode_oracle_system* ode_oracle_system_at(int i) {
    return (i >= 0 && i < ode_oracle_system_count()) ? &SYSTEMS[i] : 0;
}
const ode_oracle_system* ode_oracle_system_by_name(const char* name) {
    if (!name) return 0;
    for (int i = 0; i < ode_oracle_system_count(); ++i)
        if (!strcmp(SYSTEMS[i].name, name)) return &SYSTEMS[i];
    return 0;
}

/* ------------------------------------------------------------------ f32 forms (T = f32) */
/* the same statements with every operation in single precision, as rustc compiles the generic
 * code for T = f32 (examples/bouncing_ball.rs:15,77-86 is written for f32) */
static void lorenz_rhs32(float x, const float* y, float* dy, const float* p) {
    (void)x;
    dy[0] = p[0] * (y[1] - y[0]);
    dy[1] = y[0] * (p[2] - y[2]) - y[1];
    dy[2] = y[0] * y[1] - p[1] * y[2];
}
static void robertson_rhs32(float x, const float* y, float* dy, const float* p) {
    (void)x;
    dy[0] = -p[0] * y[0] + p[1] * y[1] * y[2];
    dy[1] = p[0] * y[0] - p[1] * y[1] * y[2] - p[2] * y[1] * y[1];
    dy[2] = p[2] * y[1] * y[1];
}
static void ball_rhs32(float x, const float* y, float* dy, const float* p) {
    (void)x;
    dy[0] = y[1];
    dy[1] = -p[0];
    dy[2] = 0.0f;
}
static int ball_solout32(float x, const float* y, const float* dy, const float* p) {
    (void)x; (void)dy; (void)p;
    return y[0] < 0.0f;
}
static void test_rk4_1_rhs32(float x, const float* y, float* dy, const float* p) {
    (void)p;
    dy[0] = (x - y[0]) / 2.0f;
}
static void test_rk4_2_rhs32(float x, const float* y, float* dy, const float* p) {
    (void)p;
    dy[0] = -2.0f * x - y[0];
}

static const ode_oracle_system32 SYSTEMS32[] = {
    {"lorenz", 3, 3, lorenz_rhs32, 0},         {"robertson", 3, 3, robertson_rhs32, 0},
    {"bouncing_ball", 3, 1, ball_rhs32, ball_solout32}, {"test_rk4_1", 1, 0, test_rk4_1_rhs32, 0},
    {"test_rk4_2", 1, 0, test_rk4_2_rhs32, 0},
};
const ode_oracle_system32* ode_oracle_system32_by_name(const char* name) {
    if (!name) return 0;
    for (unsigned q = 0; q < sizeof(SYSTEMS32) / sizeof(SYSTEMS32[0]); ++q)
        if (!strcmp(SYSTEMS32[q].name, name)) return &SYSTEMS32[q];
    return 0;
}
