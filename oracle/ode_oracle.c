/*
 * ode_oracle.c -- CPU ORACLE (test infrastructure only; see ode_oracle.h).
 *
 * Every arithmetic expression is written with the operation order of the reference
 * (one IEEE rounding per operation, never fused: compile with -ffp-contract=off).
 * nalgebra vector expressions are expanded component-wise, left to right, each
 * producing a rounded temporary, zero tableau coefficients included.
 * The step loops themselves live in ode_oracle_core.inc, written once over the scalar type and
 * included below for T = f64 and T = f32 (the crate is generic over FloatNumber, dop_shared.rs:52-77).
 */
#include "ode_oracle.h"
#include "ode_tableau_c.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD ODE_B200_MAX_DIM

/* ------------------------------------------------------------------ helpers */

/* libm as the crate sees it: f64::powf -> pow, f64::exp -> exp (glibc) */
double ode_oracle_pow(double x, double y) { return pow(x, y); }
double ode_oracle_exp(double x) { return exp(x); }

/* 1-based tableau accessors (butcher_tableau.rs:57-84, :380-426) */
static double a5(int i, int j) { return ODE_DP5_A[i - 1][j - 1]; }
static double c5(int i) { return ODE_DP5_C[i - 1]; }
static double d5(int i) { return ODE_DP5_D[i - 1]; }
static double e5(int i) { return ODE_DP5_E[i - 1]; }
static double a8(int i, int j) { return ODE_DP8_A[i - 1][j - 1]; }
static double b8(int i) { return ODE_DP8_B[i - 1]; }
static double bhh8(int i) { return ODE_DP8_BHH[i - 1]; }
static double c8(int i) { return ODE_DP8_C[i - 1]; }
static double d8(int i, int j) { return ODE_DP8_D[i - 4][j - 1]; }
static double e8(int i) { return ODE_DP8_E[i - 1]; }

double ode_oracle_tableau(const char* w, int i, int j) {
    if (!strcmp(w, "a5")) return a5(i, j);
    if (!strcmp(w, "c5")) return c5(i);
    if (!strcmp(w, "d5")) return d5(i);
    if (!strcmp(w, "e5")) return e5(i);
    if (!strcmp(w, "a8")) return a8(i, j);
    if (!strcmp(w, "b8")) return b8(i);
    if (!strcmp(w, "bhh8")) return bhh8(i);
    if (!strcmp(w, "c8")) return c8(i);
    if (!strcmp(w, "d8")) return d8(i, j);
    if (!strcmp(w, "e8")) return e8(i);
    return NAN;
}

/* ------------------------------------------------------------------ COM::evaluate */

/* continuous_output_model.rs:31-56 (evaluate) and :86-104 (get_interval_index) */
int ode_oracle_com_evaluate(int dim, int m, uint32_t n_int, double lower_bound, const double* bp,
                            const double* steps, const double* coef, double x, double* y) {
    if (x < lower_bound) return 0;
    /* slice::binary_search_by: any matching index is acceptable; the model's breakpoints are
     * strictly increasing in the supported domain so the match is unique. */
    uint32_t lo = 0, hi = n_int;
    int64_t found = -1;
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (bp[mid] == x) {
            found = mid;
            break;
        }
        if (bp[mid] < x)
            lo = mid + 1;
        else
            hi = mid;
    }
    uint32_t index;
    if (found >= 0)
        index = (uint32_t)found;
    else if (lo < n_int)
        index = lo;
    else
        return 0;
    double lower = index == 0 ? lower_bound : bp[index - 1];
    double theta = (x - lower) / steps[index], theta1 = 1.0 - theta;
    const double* c = coef + (size_t)index * m * dim;
    if (m == 0) return 0; /* reference would panic on coefficients[len-1] */
    for (int i = 0; i < dim; ++i) {
        double r = c[(size_t)(m - 1) * dim + i];
        for (int q = m - 2; q >= 0; --q) r = c[(size_t)q * dim + i] + r * ((q % 2 == 0) ? theta : theta1);
        y[i] = r;
    }
    return 1;
}

/* ------------------------------------------------------------------ constructors */

/* dopri5.rs:14-27,76-105 ; dop853.rs:14-27,76-109 */
int ode_oracle_config_new(int method, double x, double x_end, double dx, double rtol, double atol,
                          ode_b200_config* c) {
    memset(c, 0, sizeof(*c));
    c->method = method;
    c->out_type = ODE_B200_OUT_DENSE;
    c->x0 = x;
    c->x_end = x_end;
    c->dx = dx;
    c->rtol = rtol;
    c->atol = atol;
    c->safety_factor = 0.9;
    c->h_max = x_end - x;
    c->h0 = 0.0;
    c->uround = 2.220446049250313e-16;
    c->n_max = 100000;
    c->n_stiff = 1000;
    if (method == ODE_B200_DOPRI5) {
        c->alpha = 0.2 - 0.04 * 0.75;
        c->beta = 0.04;
        c->fac_max = 10.0;
        c->fac_min = 0.2;
    } else if (method == ODE_B200_DOP853) {
        c->alpha = 1.0 / 8.0;
        c->beta = 0.0;
        c->fac_max = 6.0;
        c->fac_min = 0.333;
    } else
        return -1;
    return 0;
}

/* dopri5.rs:129-184 ; dop853.rs:133-192 */
int ode_oracle_config_from_param(int method, double x, double x_end, double dx, double rtol, double atol,
                                 double safety, double beta, double fac_min, double fac_max, double h_max,
                                 double h, uint32_t n_max, uint32_t n_stiff, int out_type, ode_b200_config* c) {
    memset(c, 0, sizeof(*c));
    c->method = method;
    c->out_type = out_type;
    c->x0 = x;
    c->x_end = x_end;
    c->dx = dx;
    c->rtol = rtol;
    c->atol = atol;
    c->safety_factor = safety;
    c->beta = beta;
    c->fac_min = fac_min;
    c->fac_max = fac_max;
    c->h_max = h_max;
    c->h0 = h;
    c->uround = 2.220446049250313e-16;
    c->n_max = n_max;
    c->n_stiff = n_stiff;
    if (method == ODE_B200_DOPRI5)
        c->alpha = 0.2 - beta * 0.75;
    else if (method == ODE_B200_DOP853)
        c->alpha = 1.0 / 8.0 - beta * 0.2;
    else
        return -1;
    return 0;
}

/* rk4.rs:42-68 */
int ode_oracle_config_rk4(double x, double x_end, double step, ode_b200_config* c) {
    memset(c, 0, sizeof(*c));
    c->method = ODE_B200_RK4;
    c->out_type = ODE_B200_OUT_SPARSE;
    c->x0 = x;
    c->x_end = x_end;
    c->step_size = step;
    return 0;
}


/* ---- the same constructors with T = f32: every literal goes through T::from (an `as f32` cast of the
 * f64 value), every operation is an f32 operation */
/* dopri5.rs:14-27,76-105 (uround = T::epsilon(), Q3) ; dop853.rs:14-27,76-109 (uround = T::from(f64::EPSILON)) */
int ode_oracle_config_new_f32(int method, float x, float x_end, float dx, float rtol, float atol,
                              ode_b200_config_f32* c) {
    memset(c, 0, sizeof(*c));
    c->method = method;
    c->out_type = ODE_B200_OUT_DENSE;
    c->x0 = x;
    c->x_end = x_end;
    c->dx = dx;
    c->rtol = rtol;
    c->atol = atol;
    c->safety_factor = (float)0.9;
    c->h_max = x_end - x;
    c->h0 = 0.0f;
    c->n_max = 100000;
    c->n_stiff = 1000;
    if (method == ODE_B200_DOPRI5) {
        c->alpha = (float)(0.2 - 0.04 * 0.75);
        c->beta = (float)0.04;
        c->fac_max = (float)10.0;
        c->fac_min = (float)0.2;
        c->uround = FLT_EPSILON;
    } else if (method == ODE_B200_DOP853) {
        c->alpha = 1.0f / (float)8.0;
        c->beta = 0.0f;
        c->fac_max = (float)6.0;
        c->fac_min = (float)0.333;
        c->uround = (float)2.220446049250313e-16;
    } else
        return -1;
    return 0;
}

/* dopri5.rs:129-184 ; dop853.rs:133-192 */
int ode_oracle_config_from_param_f32(int method, float x, float x_end, float dx, float rtol, float atol,
                                     float safety, float beta, float fac_min, float fac_max, float h_max, float h,
                                     uint32_t n_max, uint32_t n_stiff, int out_type, ode_b200_config_f32* c) {
    memset(c, 0, sizeof(*c));
    c->method = method;
    c->out_type = out_type;
    c->x0 = x;
    c->x_end = x_end;
    c->dx = dx;
    c->rtol = rtol;
    c->atol = atol;
    c->safety_factor = safety;
    c->beta = beta;
    c->fac_min = fac_min;
    c->fac_max = fac_max;
    c->h_max = h_max;
    c->h0 = h;
    c->uround = (float)2.220446049250313e-16;
    c->n_max = n_max;
    c->n_stiff = n_stiff;
    if (method == ODE_B200_DOPRI5) {
        volatile float t = beta * (float)0.75;
        c->alpha = (float)0.2 - t;
    } else if (method == ODE_B200_DOP853) {
        volatile float t = beta * (float)0.2;
        c->alpha = 1.0f / (float)8.0 - t;
    } else
        return -1;
    return 0;
}

/* rk4.rs:42-68 */
int ode_oracle_config_rk4_f32(float x, float x_end, float step, ode_b200_config_f32* c) {
    memset(c, 0, sizeof(*c));
    c->method = ODE_B200_RK4;
    c->out_type = ODE_B200_OUT_SPARSE;
    c->x0 = x;
    c->x_end = x_end;
    c->step_size = step;
    return 0;
}

/* ------------------------------------------------------------------ batch driver */

int ode_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ the step loops, T = f64 */
#define REAL double
#define T_(name) name##_f64
#define RC(lit) (lit)
#define R_FABS fabs
#define R_SQRT sqrt
#define R_POW pow
#define R_FMAX fmax
#define R_FMIN fmin
#define R_CEIL ceil
#define Ta5(i, j) a5(i, j)
#define Tc5(i) c5(i)
#define Td5(i) d5(i)
#define Te5(i) e5(i)
#define Ta8(i, j) a8(i, j)
#define Tb8(i) b8(i)
#define Tbhh8(i) bhh8(i)
#define Tc8(i) c8(i)
#define Td8(i, j) d8(i, j)
#define Te8(i) e8(i)
#define SYS_T ode_oracle_system
#define SYS_BY_NAME ode_oracle_system_by_name
#define CFG_T ode_b200_config
#define OUT_T ode_b200_outputs
#include "ode_oracle_core.inc"
#undef REAL
#undef T_
#undef RC
#undef R_FABS
#undef R_SQRT
#undef R_POW
#undef R_FMAX
#undef R_FMIN
#undef R_CEIL
#undef Ta5
#undef Tc5
#undef Td5
#undef Te5
#undef Ta8
#undef Tb8
#undef Tbhh8
#undef Tc8
#undef Td8
#undef Te8
#undef SYS_T
#undef SYS_BY_NAME
#undef CFG_T
#undef OUT_T

/* ------------------------------------------------------------------ the step loops, T = f32 */
/* T::from(lit) and `.to_subset()` are `as f32` casts of the f64 value (round to nearest): the C cast.
 * f32::powf / sqrt / abs / max / min / ceil -> powf / sqrtf / fabsf / fmaxf / fminf / ceilf of glibc. */
#define REAL float
#define T_(name) name##_f32
#define RC(lit) ((float)(lit))
#define R_FABS fabsf
#define R_SQRT sqrtf
#define R_POW powf
#define R_FMAX fmaxf
#define R_FMIN fminf
#define R_CEIL ceilf
#define Ta5(i, j) ((float)a5(i, j))
#define Tc5(i) ((float)c5(i))
#define Td5(i) ((float)d5(i))
#define Te5(i) ((float)e5(i))
#define Ta8(i, j) ((float)a8(i, j))
#define Tb8(i) ((float)b8(i))
#define Tbhh8(i) ((float)bhh8(i))
#define Tc8(i) ((float)c8(i))
#define Td8(i, j) ((float)d8(i, j))
#define Te8(i) ((float)e8(i))
#define SYS_T ode_oracle_system32
#define SYS_BY_NAME ode_oracle_system32_by_name
#define CFG_T ode_b200_config_f32
#define OUT_T ode_b200_outputs_f32
#include "ode_oracle_core.inc"

int ode_oracle_integrate_batch(const char* system, const ode_b200_config* cfg, int64_t n, const double* y0,
                               const double* params, int per_traj, const ode_b200_outputs* out, int n_threads) {
    return integrate_batch_f64(system, cfg, n, y0, params, per_traj, out, n_threads);
}
int ode_oracle_integrate_batch_f32(const char* system, const ode_b200_config_f32* cfg, int64_t n, const float* y0,
                                   const float* params, int per_traj, const ode_b200_outputs_f32* out, int n_threads) {
    return integrate_batch_f32(system, cfg, n, y0, params, per_traj, out, n_threads);
}
float ode_oracle_powf(float x, float y) { return powf(x, y); }

/* Rk4::new + integrate with T = f32 (the pre-existing entry, now a wrapper of the general one) */
int ode_oracle_rk4_f32_integrate_batch(const char* system, float x, float x_end, float step_size, int64_t n,
                                       const float* y0, const float* params, int per_traj,
                                       const ode_b200_outputs_f32* out, int n_threads) {
    ode_b200_config_f32 c;
    ode_oracle_config_rk4_f32(x, x_end, step_size, &c);
    return integrate_batch_f32(system, &c, n, y0, params, per_traj, out, n_threads);
}
