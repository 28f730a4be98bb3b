/*
 * ode_oracle.c -- CPU ORACLE (test infrastructure only; see ode_oracle.h).
 *
 * Every arithmetic expression is written with the operation order of the reference
 * (one IEEE rounding per operation, never fused: compile with -ffp-contract=off).
 * nalgebra vector expressions are expanded component-wise, left to right, each
 * producing a rounded temporary, zero tableau coefficients included.
 */
#include "ode_oracle.h"
#include "ode_tableau_c.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD ODE_B200_MAX_DIM

/* ------------------------------------------------------------------ helpers */

/* dop_shared.rs:156-162 -- note b == 0 gives the NEGATIVE branch */
static double sign_(double a, double b) { return b > 0.0 ? fabs(a) : -fabs(a); }

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

/*
 * nalgebra 0.33.2 Matrix::dot for a column vector (used only by the stiffness test,
 * dopri5.rs:379-380, dop853.rs:373-374). nalgebra is not vendored under /root/reference;
 * this restates its base/blas.rs `dotx` from memory (PARITY UNPINNED):
 * special cases for dimension 2, 3, 4, otherwise 8 accumulators per block of 8 rows,
 * folded as (acc0+acc4),(acc1+acc5),(acc2+acc6),(acc3+acc7), then a sequential tail.
 */
static double nalgebra_dot(const double* a, const double* b, int n) {
    if (n == 2) return a[0] * b[0] + a[1] * b[1];
    if (n == 3) return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
    if (n == 4) {
        double p = a[0] * b[0], q = a[1] * b[1], r = a[2] * b[2], s = a[3] * b[3];
        p += r;
        q += s;
        return p + q;
    }
    double res = 0.0, acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int i = 0;
    while (n - i >= 8) {
        for (int u = 0; u < 8; ++u) acc[u] += a[i + u] * b[i + u];
        i += 8;
    }
    res += acc[0] + acc[4];
    res += acc[1] + acc[5];
    res += acc[2] + acc[6];
    res += acc[3] + acc[7];
    for (; i < n; ++i) res += a[i] * b[i];
    return res;
}

/* ------------------------------------------------------------------ controller.rs */

typedef struct {
    double alpha, beta, facc1, facc2, fac_old, h_max, safety, posneg;
    int reject;
} controller_t;

/* controller.rs:29-49 */
static void controller_new(controller_t* c, double alpha, double beta, double fac_max, double fac_min,
                           double h_max, double safety, double posneg) {
    c->alpha = alpha;
    c->beta = beta;
    c->facc1 = 1.0 / fac_min;
    c->facc2 = 1.0 / fac_max;
    c->fac_old = 1.0E-4;
    c->h_max = fabs(h_max);
    c->reject = 0;
    c->safety = safety;
    c->posneg = posneg;
}

/* controller.rs:52-77 */
static int controller_accept(controller_t* c, double err, double h, double* h_new) {
    double fac11 = pow(err, c->alpha);
    double fac = fac11 * pow(c->fac_old, -c->beta);
    fac = fmax(c->facc2, fmin(c->facc1, fac / c->safety));
    *h_new = h / fac;
    if (err <= 1.0) {
        c->fac_old = fmax(err, 1.0E-4);
        if (fabs(*h_new) > c->h_max) *h_new = c->posneg * c->h_max;
        if (c->reject) *h_new = c->posneg * fmin(fabs(*h_new), fabs(h));
        c->reject = 0;
        return 1;
    }
    *h_new = h / fmin(c->facc1, fac11 / c->safety);
    c->reject = 1;
    return 0;
}

/* ------------------------------------------------------------------ result sink */

/* SolverResult (dop_shared.rs:46,79-109) + the ContinuousOutputModel being filled */
typedef struct {
    int dim;
    int64_t cap;
    uint32_t count;
    double *x, *y;       /* may be NULL */
    double last[MAXD];   /* results.get().1.last() */
    int64_t com_cap;
    uint32_t com_count;
    double *com_break, *com_step, *com_coef; /* may be NULL */
    int com_m;
} sink_t;

static void sink_push(sink_t* s, double x, const double* y) {
    if (s->x && s->y && (int64_t)s->count < s->cap) {
        s->x[s->count] = x;
        memcpy(s->y + (size_t)s->count * s->dim, y, sizeof(double) * s->dim);
    }
    memcpy(s->last, y, sizeof(double) * s->dim);
    s->count++;
}

/* continuous_output_model.rs:72-83 */
static void sink_add_interval(sink_t* s, double breakpoint, double (*rc)[MAXD], double step) {
    if ((int64_t)s->com_count < s->com_cap) {
        if (s->com_break) s->com_break[s->com_count] = breakpoint;
        if (s->com_step) s->com_step[s->com_count] = step;
        if (s->com_coef)
            for (int r = 0; r < s->com_m; ++r)
                memcpy(s->com_coef + ((size_t)s->com_count * s->com_m + r) * s->dim, rc[r],
                       sizeof(double) * s->dim);
    }
    s->com_count++;
}

typedef struct {
    double y[MAXD];
    double x, h_next;
    uint32_t num_eval, accepted, rejected, n_step;
    int status;
    double com_lower;
} traj_result_t;

#define DENSE_GUARD (1u << 26)

/* ------------------------------------------------------------------ hinit */

/* dopri5.rs:187-243 and dop853.rs:195-251 (identical except the exponent) */
static double hinit(const ode_oracle_system* S, const double* p, double x, double x_end, const double* y,
                    double rtol, double atol, double h_max_abs, double expo) {
    int n = S->dim;
    double f0[MAXD] = {0}, f1[MAXD] = {0}, y1[MAXD];
    S->rhs(x, y, f0, p);
    double posneg = sign_(1.0, x_end - x);
    double d0 = 0.0, d1 = 0.0;
    for (int i = 0; i < n; ++i) {
        double sci = atol + fabs(y[i]) * rtol;
        d0 += (y[i] / sci) * (y[i] / sci);
        d1 += (f0[i] / sci) * (f0[i] / sci);
    }
    double h0 = (d0 < 1.0e-10 || d1 < 1.0e-10) ? 1.0e-6 : 0.01 * sqrt(d0 / d1);
    h0 = fmin(h0, h_max_abs);
    h0 = sign_(h0, posneg);
    for (int i = 0; i < n; ++i) y1[i] = y[i] + f0[i] * h0;
    S->rhs(x + h0, y1, f1, p);
    double d2 = 0.0;
    for (int i = 0; i < n; ++i) {
        double sci = atol + fabs(y[i]) * rtol;
        d2 += ((f1[i] - f0[i]) / sci) * ((f1[i] - f0[i]) / sci);
    }
    d2 = sqrt(d2) / h0;
    double h1;
    if (fmax(sqrt(d1), fabs(d2)) <= 1.0E-15)
        h1 = fmax(1.0E-6, fabs(h0) * 1.0E-3);
    else
        h1 = pow(0.01 / fmax(sqrt(d1), d2), expo); /* second d2 WITHOUT abs, as in the source */
    return sign_(fmin(100.0 * fabs(h0), fmin(h1, h_max_abs)), posneg);
}

/* ------------------------------------------------------------------ rk4.rs */

/* Rk4::integrate + step + populate_buffer, rk4.rs:71-143 */
static void rk4_integrate(const ode_oracle_system* S, const ode_b200_config* cfg, const double* y0,
                          const double* p, sink_t* sink, traj_result_t* R) {
    int n = S->dim;
    double x = cfg->x0, step = cfg->step_size, half = step / 2.0;
    double y[MAXD], k[4][MAXD], buf[MAXD], ynew[MAXD];
    memcpy(y, y0, sizeof(double) * n);
    memset(k, 0, sizeof(k)); /* rk4.rs:48-53: k preallocated as zeros; system() may leave slots untouched */
    sink_push(sink, x, y);
    R->status = ODE_B200_OK;
    double ns = ceil((cfg->x_end - x) / step);
    /* to_usize().unwrap() panics for negative / NaN */
    if (!(ns >= 0.0) || ns > 4.0e9) {
        R->status = ODE_B200_UNSUPPORTED_DOMAIN;
        ns = 0.0;
    }
    uint64_t num_steps = (uint64_t)ns;
    for (uint64_t it = 0; it < num_steps; ++it) {
        S->rhs(x, y, k[0], p);
        for (int i = 0; i < n; ++i) buf[i] = y[i] + k[0][i] * half;
        S->rhs(x + half, buf, k[1], p);
        for (int i = 0; i < n; ++i) buf[i] = y[i] + k[1][i] * half;
        S->rhs(x + half, buf, k[2], p);
        for (int i = 0; i < n; ++i) buf[i] = y[i] + k[2][i] * step;
        S->rhs(x + step, buf, k[3], p);
        double x_new = x + step;
        for (int i = 0; i < n; ++i)
            ynew[i] = y[i] + (((k[0][i] + k[1][i] * 2.0) + k[2][i] * 2.0) + k[3][i]) * (step / 6.0);
        int abort_ = S->solout ? S->solout(x_new, ynew, k[0], p) : 0;
        x = x_new;
        memcpy(y, ynew, sizeof(double) * n);
        sink_push(sink, x_new, ynew);
        R->num_eval += 4;
        R->accepted += 1;
        R->n_step += 1;
        if (abort_) break;
    }
    memcpy(R->y, y, sizeof(double) * n);
    R->x = x;
    R->h_next = step;
}

/* ------------------------------------------------------------------ dopri5.rs */

/* compute_y_out, dopri5.rs:490-495 */
static void dp5_y_out(int n, double (*rc)[MAXD], double th, double th1, double* out) {
    for (int i = 0; i < n; ++i)
        out[i] = rc[0][i] + (rc[1][i] + (rc[2][i] + (rc[3][i] + rc[4][i] * th1) * th) * th1) * th;
}

/* integrate_core, dopri5.rs:263-445 with solution_output :447-487 inlined */
static void dopri5_integrate(const ode_oracle_system* S, const ode_b200_config* cfg, const double* y0,
                             const double* p, sink_t* sink, traj_result_t* R) {
    const int n = S->dim;
    const int out_type = cfg->out_type;
    const int dense = out_type == ODE_B200_OUT_DENSE, cont = out_type == ODE_B200_OUT_CONTINUOUS;
    double x = cfg->x0, x_end = cfg->x_end, xd = cfg->x0, dx = cfg->dx;
    double rtol = cfg->rtol, atol = cfg->atol, uround = cfg->uround;
    double y[MAXD], k[7][MAXD], y_next[MAXD], y_stiff[MAXD], rc[5][MAXD], tmp[MAXD];
    memcpy(y, y0, sizeof(double) * n);
    memset(k, 0, sizeof(k)); /* dopri5.rs:294 */
    memset(rc, 0, sizeof(rc));
    memset(y_next, 0, sizeof(y_next));
    memset(y_stiff, 0, sizeof(y_stiff));
    controller_t C;
    controller_new(&C, cfg->alpha, cfg->beta, cfg->fac_max, cfg->fac_min, cfg->h_max, cfg->safety_factor,
                   sign_(1.0, x_end - x));

    double x_old = x, h = cfg->h0, h_old, h_new = 0.0;
    uint32_t n_step = 0, non_stiff = 0, iasti = 0;
    int last = 0;
    const double posneg = sign_(1.0, x_end - x);
    R->status = ODE_B200_OK;
    R->com_lower = x; /* set_lower_bound, dopri5.rs:279-281 */

    if (h == 0.0) {
        h = hinit(S, p, x, x_end, y, rtol, atol, C.h_max, 1.0 / 5.0);
        R->num_eval += 2;
    }
    h_old = h;
    if (out_type == ODE_B200_OUT_SPARSE) sink_push(sink, x, y);
    S->rhs(x, y, k[0], p);
    R->num_eval += 1;

    while (!last) {
        if (n_step > cfg->n_max) { /* '>' (Q2) */
            h_old = h;
            R->status = ODE_B200_MAX_NUM_STEP_REACHED;
            break;
        }
        if (0.1 * fabs(h) <= uround * fabs(x)) {
            h_old = h;
            R->status = ODE_B200_STEP_SIZE_UNDERFLOW;
            break;
        }
        if ((x + 1.01 * h - x_end) * posneg > 0.0) {
            h = x_end - x;
            last = 1;
        }
        n_step += 1;

        for (int s = 1; s < 7; ++s) {
            for (int i = 0; i < n; ++i) y_next[i] = y[i];
            for (int j = 0; j < s; ++j)
                for (int i = 0; i < n; ++i) y_next[i] = y_next[i] + (k[j][i] * h) * a5(s + 1, j + 1);
            S->rhs(x + h * c5(s + 1), y_next, k[s], p);
            if (s == 5) memcpy(y_stiff, y_next, sizeof(double) * n);
        }
        memcpy(k[1], k[6], sizeof(double) * n);
        R->num_eval += 6;

        if (dense || cont) /* every attempt (Q4) */
            for (int i = 0; i < n; ++i)
                rc[4][i] = (((((k[0][i] * d5(1) + k[2][i] * d5(3)) + k[3][i] * d5(4)) + k[4][i] * d5(5)) +
                             k[5][i] * d5(6)) +
                            k[1][i] * d5(7)) *
                           h;
        for (int i = 0; i < n; ++i)
            tmp[i] = (((((k[0][i] * e5(1) + k[2][i] * e5(3)) + k[3][i] * e5(4)) + k[4][i] * e5(5)) +
                       k[5][i] * e5(6)) +
                      k[1][i] * e5(7)) *
                     h;
        memcpy(k[3], tmp, sizeof(double) * n);

        double err = 0.0;
        for (int i = 0; i < n; ++i) {
            double sc = atol + fmax(fabs(y[i]), fabs(y_next[i])) * rtol;
            err += (k[3][i] / sc) * (k[3][i] / sc);
        }
        err = sqrt(err / (double)n);

        if (controller_accept(&C, err, h, &h_new)) {
            R->accepted += 1;
            if (R->accepted % cfg->n_stiff == 0 || iasti > 0) {
                double dk[MAXD], dyv[MAXD];
                for (int i = 0; i < n; ++i) {
                    dk[i] = k[1][i] - k[5][i];
                    dyv[i] = y_next[i] - y_stiff[i];
                }
                double num = nalgebra_dot(dk, dk, n), den = nalgebra_dot(dyv, dyv, n);
                double h_lamb = den > 0.0 ? h * sqrt(num / den) : 0.0;
                if (h_lamb > 3.25) {
                    iasti += 1;
                    non_stiff = 0;
                    if (iasti == 15) {
                        h_old = h;
                        R->status = ODE_B200_STIFFNESS_DETECTED;
                        break;
                    }
                } else {
                    non_stiff += 1;
                    if (non_stiff == 6) iasti = 0;
                }
            }
            if (dense || cont) {
                for (int i = 0; i < n; ++i) {
                    double ydiff = y_next[i] - y[i];
                    double bspl = k[0][i] * h - ydiff;
                    rc[0][i] = y[i];
                    rc[1][i] = ydiff;
                    rc[2][i] = bspl;
                    rc[3][i] = ((-k[1][i]) * h + ydiff) - bspl;
                }
            }
            memcpy(k[0], k[1], sizeof(double) * n);
            memcpy(y, y_next, sizeof(double) * n);
            x_old = x;
            x += h;
            h_old = h;

            /* solution_output(y_next, com, &k[0]) */
            if (dense) {
                int bad = 0;
                uint32_t guard = 0;
                while (fabs(xd) <= fabs(x)) {
                    int pushed = 0;
                    if (fabs(x_old) <= fabs(xd) && fabs(x) >= fabs(xd)) {
                        double th = (xd - x_old) / h_old, th1 = 1.0 - th;
                        dp5_y_out(n, rc, th, th1, tmp);
                        sink_push(sink, xd, tmp);
                        xd += dx;
                        pushed = 1;
                    }
                    if (sink->count == 0) { bad = 1; break; } /* reference: unwrap() on empty -> panic */
                    if (S->solout && S->solout(xd, sink->last, k[0], p)) break;
                    /* reference: infinite loop (nothing changes between passes once no sample is
                     * pushed, or when dx == 0 leaves xd where it was) */
                    if (!pushed || dx == 0.0 || ++guard > DENSE_GUARD) { bad = 1; break; }
                }
                if (bad) {
                    R->status = ODE_B200_UNSUPPORTED_DOMAIN;
                    break;
                }
                if (fabs(xd - x_end) < 1.0e-9) {
                    double th = (x_end - x_old) / h_old, th1 = 1.0 - th;
                    dp5_y_out(n, rc, th, th1, tmp);
                    sink_push(sink, x_end, tmp);
                    xd += dx;
                }
            } else {
                sink_push(sink, x, y_next);
            }
            if (cont) sink_add_interval(sink, fabs(x), rc, h_old);

            if (sink->count == 0) { /* results.last().unwrap() would panic */
                R->status = ODE_B200_UNSUPPORTED_DOMAIN;
                break;
            }
            if (S->solout && S->solout(x, sink->last, k[0], p)) last = 1;
            if (last) {
                h_old = posneg * h_new;
                break;
            }
        } else {
            last = 0;
            if (R->accepted >= 1) R->rejected += 1;
        }
        h = h_new;
    }
    memcpy(R->y, y, sizeof(double) * n);
    R->x = x;
    R->h_next = h_old;
    R->n_step = n_step;
}

/* ------------------------------------------------------------------ dop853.rs */

/* compute_y_out, dop853.rs:546-559 */
static void dp8_y_out(int n, double (*rc)[MAXD], double th, double th1, double* out) {
    for (int i = 0; i < n; ++i)
        out[i] = rc[0][i] +
                 (rc[1][i] +
                  (rc[2][i] +
                   (rc[3][i] + (rc[4][i] + (rc[5][i] + (rc[6][i] + rc[7][i] * th) * th1) * th) * th1) * th) *
                      th1) *
                     th;
}

typedef struct {
    double x0, x, x_old, x_end, xd, dx, h_old;
    double y[MAXD];
    double rc[8][MAXD];
    int out_type, n;
    int bad;
} dp8_state_t;

/* solution_output, dop853.rs:515-543 */
static void dp8_solution_output(dp8_state_t* st, sink_t* sink, const double* y_next) {
    double tmp[MAXD];
    if (st->out_type == ODE_B200_OUT_DENSE) {
        if (fabs(st->xd - st->x0) < 2.220446049250313e-16) {
            sink_push(sink, st->x0, st->y);
            st->xd += st->dx;
        } else {
            uint32_t guard = 0;
            while (fabs(st->xd) <= fabs(st->x)) {
                if (fabs(st->x_old) <= fabs(st->xd) && fabs(st->x) >= fabs(st->xd)) {
                    double th = (st->xd - st->x_old) / st->h_old, th1 = 1.0 - th;
                    dp8_y_out(st->n, st->rc, th, th1, tmp);
                    sink_push(sink, st->xd, tmp);
                    st->xd += st->dx;
                    if (++guard > DENSE_GUARD) { st->bad = 1; return; }
                } else { /* reference: infinite loop */
                    st->bad = 1;
                    return;
                }
            }
            if (fabs(st->xd - st->x_end) < 1.0e-9) {
                double th = (st->x_end - st->x_old) / st->h_old, th1 = 1.0 - th;
                dp8_y_out(st->n, st->rc, th, th1, tmp);
                sink_push(sink, st->x_end, tmp);
                st->xd += st->dx;
            }
        }
    } else if (st->out_type == ODE_B200_OUT_CONTINUOUS8) {
        sink_push(sink, st->x, y_next); /* extension: (x, y) like Dopri5's Continuous push (dopri5.rs:479-481) */
    } else {
        sink_push(sink, st->x0, y_next); /* pushes x0, not x (Q6) */
    }
}

/* integrate, dop853.rs:254-512 */
static void dop853_integrate(const ode_oracle_system* S, const ode_b200_config* cfg, const double* y0,
                             const double* p, sink_t* sink, traj_result_t* R) {
    const int n = S->dim;
    dp8_state_t st;
    memset(&st, 0, sizeof(st));
    st.n = n;
    st.out_type = cfg->out_type;
    st.x0 = st.x = st.xd = st.x_old = cfg->x0;
    st.x_end = cfg->x_end;
    st.dx = cfg->dx;
    memcpy(st.y, y0, sizeof(double) * n);
    const double rtol = cfg->rtol, atol = cfg->atol, uround = cfg->uround;
    /* ODE_B200_OUT_CONTINUOUS8 (extension, see include/ode_b200.h): Dense-mode numerics, intervals recorded */
    const int cont8 = cfg->out_type == ODE_B200_OUT_CONTINUOUS8;
    const int dense = cfg->out_type == ODE_B200_OUT_DENSE || cont8;
    double k[12][MAXD], y_next[MAXD], err_est[MAXD], err_bhh[MAXD], t1[MAXD], t2[MAXD];
    memset(k, 0, sizeof(k)); /* dop853.rs:276 */
    memset(y_next, 0, sizeof(y_next));
    controller_t C;
    controller_new(&C, cfg->alpha, cfg->beta, cfg->fac_max, cfg->fac_min, cfg->h_max, cfg->safety_factor,
                   sign_(1.0, st.x_end - st.x));
    double h = cfg->h0, h_new = 0.0;
    uint32_t n_step = 0, iasti = 0, non_stiff = 0;
    int last = 0;
    const double posneg = sign_(1.0, st.x_end - st.x);
    R->status = ODE_B200_OK;
    R->com_lower = cont8 ? st.x : 0.0;

    if (h == 0.0) {
        h = hinit(S, p, st.x, st.x_end, st.y, rtol, atol, C.h_max, 0.125);
        R->num_eval += 2;
    }
    st.h_old = h;
    dp8_solution_output(&st, sink, st.y);
    S->rhs(st.x, st.y, k[0], p);
    R->num_eval += 1;

    while (!last && !st.bad) {
        if (n_step >= cfg->n_max) { /* '>=' (Q2) */
            st.h_old = h;
            R->status = ODE_B200_MAX_NUM_STEP_REACHED;
            break;
        }
        if (0.1 * fabs(h) <= uround * fabs(st.x)) {
            st.h_old = h;
            R->status = ODE_B200_STEP_SIZE_UNDERFLOW;
            break;
        }
        if ((st.x + 1.01 * h - st.x_end) * posneg > 0.0) {
            h = st.x_end - st.x;
            last = 1;
        }
        n_step += 1;

        for (int s = 1; s < 12; ++s) {
            for (int i = 0; i < n; ++i) y_next[i] = st.y[i];
            for (int j = 0; j < s; ++j) {
                double ha = h * a8(s + 1, j + 1);
                for (int i = 0; i < n; ++i) y_next[i] = y_next[i] + k[j][i] * ha;
            }
            S->rhs(st.x + (h * c8(s + 1)), y_next, k[s], p);
        }
        memcpy(k[1], k[10], sizeof(double) * n);
        memcpy(k[2], k[11], sizeof(double) * n);
        R->num_eval += 11;

        for (int i = 0; i < n; ++i)
            t1[i] = ((((((k[0][i] * b8(1) + k[5][i] * b8(6)) + k[6][i] * b8(7)) + k[7][i] * b8(8)) +
                       k[8][i] * b8(9)) +
                      k[9][i] * b8(10)) +
                     k[1][i] * b8(11)) +
                    k[2][i] * b8(12);
        memcpy(k[3], t1, sizeof(double) * n);
        for (int i = 0; i < n; ++i) k[4][i] = st.y[i] + k[3][i] * h;

        for (int i = 0; i < n; ++i) err_est[i] = 0.0;
        for (int q = 0; q < 12; ++q)
            for (int i = 0; i < n; ++i) err_est[i] = err_est[i] + k[q][i] * e8(q + 1);

        double err = 0.0, err2 = 0.0;
        for (int i = 0; i < n; ++i)
            err_bhh[i] = ((k[3][i] - k[0][i] * bhh8(1)) - k[8][i] * bhh8(2)) - k[2][i] * bhh8(3);
        for (int i = 0; i < n; ++i) {
            double sc = atol + fmax(fabs(st.y[i]), fabs(k[4][i])) * rtol;
            err += (err_est[i] / sc) * (err_est[i] / sc);
            err2 += (err_bhh[i] / sc) * (err_bhh[i] / sc);
        }
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = fabs(h) * err * sqrt(1.0 / (deno * (double)n));

        if (controller_accept(&C, err, h, &h_new)) {
            R->accepted += 1;
            memcpy(t1, k[4], sizeof(double) * n);
            S->rhs(st.x + h, t1, k[3], p);
            R->num_eval += 1;

            if (R->accepted % cfg->n_stiff == 0 || iasti > 0) {
                for (int i = 0; i < n; ++i) {
                    t1[i] = k[3][i] - k[2][i];
                    t2[i] = k[4][i] - y_next[i];
                }
                double num = nalgebra_dot(t1, t1, n), den = nalgebra_dot(t2, t2, n);
                double h_lamb = den > 0.0 ? h * sqrt(num / den) : 0.0;
                if (h_lamb > 6.1) {
                    non_stiff = 0;
                    iasti += 1;
                    if (iasti == 15) {
                        st.h_old = h;
                        R->status = ODE_B200_STIFFNESS_DETECTED;
                        break;
                    }
                } else {
                    non_stiff += 1;
                    if (non_stiff == 6) iasti = 0;
                }
            }

            if (dense) {
                double (*rc)[MAXD] = st.rc;
                for (int i = 0; i < n; ++i) {
                    rc[0][i] = st.y[i];
                    double y_diff = k[4][i] - st.y[i];
                    rc[1][i] = y_diff;
                    double bspl = k[0][i] * h - y_diff;
                    rc[2][i] = bspl;
                    rc[3][i] = (y_diff - k[3][i] * h) - bspl;
                }
                for (int r = 4; r < 8; ++r) {
                    for (int i = 0; i < n; ++i) rc[r][i] = 0.0;
                    for (int j = 1; j < 13; ++j)
                        for (int i = 0; i < n; ++i) rc[r][i] = rc[r][i] + k[j - 1][i] * d8(r, j);
                }
                double yn[MAXD];
                for (int i = 0; i < n; ++i)
                    yn[i] = st.y[i] + (((((((k[0][i] * a8(14, 1) + k[6][i] * a8(14, 7)) + k[7][i] * a8(14, 8)) +
                                            k[8][i] * a8(14, 9)) +
                                           k[9][i] * a8(14, 10)) +
                                          k[1][i] * a8(14, 11)) +
                                         k[2][i] * a8(14, 12)) +
                                        k[3][i] * a8(14, 13)) *
                                           h;
                S->rhs(st.x + h * c8(14), yn, k[9], p);
                for (int i = 0; i < n; ++i)
                    yn[i] = st.y[i] + (((((((k[0][i] * a8(15, 1) + k[5][i] * a8(15, 6)) + k[6][i] * a8(15, 7)) +
                                            k[7][i] * a8(15, 8)) +
                                           k[1][i] * a8(15, 11)) +
                                          k[2][i] * a8(15, 12)) +
                                         k[3][i] * a8(15, 13)) +
                                        k[9][i] * a8(15, 14)) *
                                           h;
                S->rhs(st.x + h * c8(15), yn, k[1], p);
                for (int i = 0; i < n; ++i)
                    yn[i] = st.y[i] + (((((((k[0][i] * a8(16, 1) + k[5][i] * a8(16, 6)) + k[6][i] * a8(16, 7)) +
                                            k[7][i] * a8(16, 8)) +
                                           k[8][i] * a8(16, 9)) +
                                          k[3][i] * a8(16, 13)) +
                                         k[9][i] * a8(16, 14)) +
                                        k[1][i] * a8(16, 15)) *
                                           h;
                S->rhs(st.x + h * c8(16), yn, k[2], p);
                R->num_eval += 3;
                for (int r = 4; r < 8; ++r)
                    for (int i = 0; i < n; ++i)
                        rc[r][i] = ((((rc[r][i] + k[3][i] * d8(r, 13)) + k[9][i] * d8(r, 14)) +
                                     k[1][i] * d8(r, 15)) +
                                    k[2][i] * d8(r, 16)) *
                                   h;
            }

            memcpy(k[0], k[3], sizeof(double) * n);
            memcpy(st.y, k[4], sizeof(double) * n);
            st.x_old = st.x;
            st.x += h;
            st.h_old = h;

            dp8_solution_output(&st, sink, k[4]);
            if (st.bad) break;
            if (cont8) sink_add_interval(sink, fabs(st.x), st.rc, st.h_old);

            if (sink->count == 0) {
                st.bad = 1;
                break;
            }
            if (S->solout && S->solout(st.x, sink->last, k[0], p)) last = 1;
            if (last) {
                st.h_old = posneg * h_new;
                break;
            }
        } else {
            last = 0;
            if (R->accepted >= 1) R->rejected += 1;
        }
        h = h_new;
    }
    if (st.bad) R->status = ODE_B200_UNSUPPORTED_DOMAIN;
    memcpy(R->y, st.y, sizeof(double) * n);
    R->x = st.x;
    R->h_next = st.h_old;
    R->n_step = n_step;
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

/* ------------------------------------------------------------------ batch driver */

int ode_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int ode_oracle_integrate_batch(const char* system, const ode_b200_config* cfg, int64_t n,
                               const double* y0, const double* params, int per_traj,
                               const ode_b200_outputs* out, int n_threads) {
    const ode_oracle_system* S = ode_oracle_system_by_name(system);
    if (!S || !cfg || n < 0 || !y0 || !out) return -1;
    if (S->n_params > 0 && !params) return -1;
    const int d = S->dim;
    int m = cfg->method == ODE_B200_DOPRI5 ? 5 : 8;
#ifdef _OPENMP
    if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
    n_threads = 1;
#endif
    (void)n_threads;
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
    for (int64_t i = 0; i < n; ++i) {
        double y[MAXD], p[ODE_B200_MAX_PARAMS + 1];
        for (int c = 0; c < d; ++c) y[c] = y0[(size_t)c * n + i];
        for (int c = 0; c < S->n_params; ++c) p[c] = per_traj ? params[(size_t)c * n + i] : params[c];
        sink_t sink;
        memset(&sink, 0, sizeof(sink));
        sink.dim = d;
        sink.cap = out->out_cap;
        if (out->out_x && out->out_y && out->out_cap > 0) {
            sink.x = out->out_x + (size_t)i * out->out_cap;
            sink.y = out->out_y + (size_t)i * out->out_cap * d;
        }
        sink.com_cap = out->com_cap;
        sink.com_m = m;
        if (out->com_cap > 0) {
            if (out->com_break) sink.com_break = out->com_break + (size_t)i * out->com_cap;
            if (out->com_step) sink.com_step = out->com_step + (size_t)i * out->com_cap;
            if (out->com_coef) sink.com_coef = out->com_coef + (size_t)i * out->com_cap * m * d;
        }
        traj_result_t R;
        memset(&R, 0, sizeof(R));
        if (cfg->method != ODE_B200_RK4 && cfg->n_stiff == 0) { /* reference: `% 0` panics */
            memcpy(R.y, y, sizeof(double) * d);
            R.x = cfg->x0;
            R.status = ODE_B200_UNSUPPORTED_DOMAIN;
        } else if (cfg->method == ODE_B200_RK4)
            rk4_integrate(S, cfg, y, p, &sink, &R);
        else if (cfg->method == ODE_B200_DOPRI5)
            dopri5_integrate(S, cfg, y, p, &sink, &R);
        else
            dop853_integrate(S, cfg, y, p, &sink, &R);
        if (R.status == ODE_B200_OK &&
            ((out->out_cap > 0 && (int64_t)sink.count > out->out_cap) ||
             (out->com_cap > 0 && (int64_t)sink.com_count > out->com_cap)))
            R.status = ODE_B200_OUTPUT_CAPACITY_EXCEEDED;
        if (out->y_final)
            for (int c = 0; c < d; ++c) out->y_final[(size_t)c * n + i] = R.y[c];
        if (out->x_final) out->x_final[i] = R.x;
        if (out->h_next) out->h_next[i] = R.h_next;
        if (out->num_eval) out->num_eval[i] = R.num_eval;
        if (out->accepted) out->accepted[i] = R.accepted;
        if (out->rejected) out->rejected[i] = R.rejected;
        if (out->n_step) out->n_step[i] = R.n_step;
        if (out->status) out->status[i] = R.status;
        if (out->out_count) out->out_count[i] = sink.count;
        if (out->com_count) out->com_count[i] = sink.com_count;
        if (out->com_lower) out->com_lower[i] = R.com_lower;
    }
    return 0;
}
