/*
 * ode_f32.cuh -- the three steppers with T = f32 (`impl FloatNumber for f32`, src/dop_shared.rs:74;
 * examples/bouncing_ball.rs:15,35-36 runs Rk4 in f32 and names Dopri5 in f32 as the alternative).
 *
 * A coverage path, not the hot one: ONE generic stepper per method, one trajectory per thread, written
 * with plain loops over the state and the stages exactly as the Rust source is (every term of every
 * stage sum, zero tableau coefficients included; the reference's slot recycling k[1] = k[6] /
 * k[1] = k[10], k[2] = k[11], k[3], k[4] as in src/dopri5.rs:326-360 and src/dop853.rs:307-347), every
 * operation in single precision (-fmad=false keeps FMUL and FADD apart; division and sqrt are IEEE),
 * every literal and tableau coefficient the f64 value rounded to f32 (T::from / to_subset), and
 * f32::powf the bit-exact libm clone of ode_powf.h. The output type is a runtime switch.
 * Reference lines are cited at each block; the statements follow Dopri5Stepper / Dop853Stepper /
 * Rk4Stepper (ode_kernels.cuh) minus their tuning.
 */
#pragma once

#include "ode_device.cuh"
#include "ode_powf.h"

namespace ode_b200 {

static __device__ const unsigned long long g_ode_powf_log2_tab[32] = ODE_POWF_LOG2_TAB_INIT;
static __device__ const unsigned long long g_ode_exp2f_tab[32] = ODE_EXP2F_TAB_INIT;
__shared__ unsigned long long s_ode_powf_tabs[64];

/* every thread of the CTA must call this once before powf_f_ */
ODE_DEV void stage_powf_tables_() {
    for (int i = threadIdx.x; i < 64; i += blockDim.x)
        s_ode_powf_tabs[i] = i < 32 ? g_ode_powf_log2_tab[i] : g_ode_exp2f_tab[i - 32];
    __syncthreads();
}
/* f32::powf as the reference's host computes it (bit-identical to glibc's powf) */
ODE_DEV float powf_f_(float x, float y) {
    ode_powf_tabs T;
    T.lt = s_ode_powf_tabs;
    T.et = s_ode_powf_tabs + 32;
    return ode_powf(x, y, T);
}

#define F32C(lit) ((float)(lit)) /* T::from(lit).unwrap(): the f64 literal rounded to f32 at compile time */

ODE_DEV float signf_(float a, float b) { return b > 0.0f ? fabsf(a) : -fabsf(a); } /* dop_shared.rs:156-162 */

/* nalgebra 0.33.2 `dot` order (see diff_dot_, ode_device.cuh) in f32 */
template <int N>
ODE_DEV float diff_dot_f_(const float (&a)[N], const float (&b)[N]) {
    float d[N];
#pragma unroll
    for (int i = 0; i < N; ++i) d[i] = a[i] - b[i];
    if constexpr (N == 1) {
        float res = 0.0f; /* generic path: no full block of 8, sequential tail */
        res += d[0] * d[0];
        return res;
    } else if constexpr (N == 2) {
        return d[0] * d[0] + d[1] * d[1];
    } else if constexpr (N == 3) {
        return (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
    } else if constexpr (N == 4) {
        float p = d[0] * d[0], q = d[1] * d[1], r = d[2] * d[2], s = d[3] * d[3];
        p += r;
        q += s;
        return p + q;
    } else {
        float res = 0.0f, acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        constexpr int NB = N / 8;
#pragma unroll
        for (int b8 = 0; b8 < NB; ++b8)
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] += d[8 * b8 + u] * d[8 * b8 + u];
        res += acc[0] + acc[4];
        res += acc[1] + acc[5];
        res += acc[2] + acc[6];
        res += acc[3] + acc[7];
#pragma unroll
        for (int i = 8 * NB; i < N; ++i) res += d[i] * d[i];
        return res;
    }
}

template <class Sys, int METHOD>
struct F32Stepper {
    static constexpr int N = Sys::DIM;
    static constexpr int NPP = Sys::NP > 0 ? Sys::NP : 1;
    static constexpr bool kStage = false, kLockstep = false;
    static constexpr int kRefillIdle = 1, kThreads = 128, kMinBlocks = 4;

    struct State {
        float y[N], k0[N], p[NPP], ylast[N]; /* k0 = k[0] = f(x, y); ylast = results.last() */
        float x, h, x_old, h_old, xd;
        float fac_old;
        bool reject;
        Counters c;
        unsigned iasti, non_stiff;
        unsigned long long steps_left;
        long long traj;
        int status;
    };

    ODE_DEV static void rhs(float x, const float (&y)[N], float (&dy)[N], const float (&p)[NPP]) {
        MathExact mx;
        Sys::rhs(x, y, dy, p, mx);
    }

    /* SolverResult::push (dop_shared.rs:88-91) */
    ODE_DEV static void push(const KernelArgsF32& a, State& s, float x, const float (&y)[N]) {
        if (a.out.out_x != nullptr && (long long)s.c.out_count < a.out.out_cap) {
            const size_t q = (size_t)s.traj * (size_t)a.out.out_cap + s.c.out_count;
            a.out.out_x[q] = x;
#pragma unroll
            for (int i = 0; i < N; ++i) a.out.out_y[q * N + i] = y[i];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) s.ylast[i] = y[i];
        ++s.c.out_count;
    }
    /* ContinuousOutputModel::add_interval (continuous_output_model.rs:72-83) */
    ODE_DEV static void add_interval(const KernelArgsF32& a, State& s, float brk, const float (&rc)[5][N], float step) {
        const ode_b200_outputs_f32& o = a.out;
        if (o.com_break != nullptr && (long long)s.c.com_count < o.com_cap) {
            const size_t q = (size_t)s.traj * (size_t)o.com_cap + s.c.com_count;
            o.com_break[q] = brk;
            o.com_step[q] = step;
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int i = 0; i < N; ++i) o.com_coef[(q * 5 + r) * N + i] = rc[r][i];
        }
        ++s.c.com_count;
    }
    ODE_DEV static bool finish(const KernelArgsF32& a, State& s, int status) {
        const ode_b200_outputs_f32& o = a.out;
        const long long t = s.traj;
        if (status == ODE_B200_OK && ((o.out_cap > 0 && (long long)s.c.out_count > o.out_cap) ||
                                      (o.com_cap > 0 && (long long)s.c.com_count > o.com_cap)))
            status = ODE_B200_OUTPUT_CAPACITY_EXCEEDED;
        if (o.y_final)
#pragma unroll
            for (int i = 0; i < N; ++i) o.y_final[(size_t)i * a.n_traj + t] = s.y[i];
        if (o.x_final) o.x_final[t] = s.x;
        if (o.h_next) o.h_next[t] = s.h_old;
        if (o.num_eval) o.num_eval[t] = s.c.num_eval;
        if (o.accepted) o.accepted[t] = s.c.accepted;
        if (o.rejected) o.rejected[t] = s.c.rejected;
        if (o.n_step) o.n_step[t] = s.c.n_step;
        if (o.status) o.status[t] = status;
        if (o.out_count) o.out_count[t] = s.c.out_count;
        if (o.com_count) o.com_count[t] = s.c.com_count;
        if (o.com_lower) /* set_lower_bound, dopri5.rs:279-281 (never reached when `% n_stiff` would panic) */
            o.com_lower[t] = (METHOD == ODE_B200_DOPRI5 && a.cfg.n_stiff != 0) ? a.cfg.x0 : 0.0f;
        return true;
    }

    /* Controller::accept, controller.rs:52-77 (Controller::new's facc1 / facc2 / |h_max| recomputed from the
     * config: the same f32 divisions) */
    ODE_DEV static bool ctrl_accept(const KernelArgsF32& a, State& s, float err, float h, float& h_new) {
        const ode_b200_config_f32& c = a.cfg;
        const float facc1 = 1.0f / c.fac_min, facc2 = 1.0f / c.fac_max, h_max = fabsf(c.h_max);
        const float posneg = signf_(1.0f, c.x_end - c.x0);
        const float fac11 = powf_f_(err, c.alpha);
        float fac = fac11 * powf_f_(s.fac_old, -c.beta);
        fac = fmaxf(facc2, fminf(facc1, fac / c.safety_factor));
        h_new = h / fac;
        if (err <= 1.0f) {
            s.fac_old = fmaxf(err, F32C(1.0E-4));
            if (fabsf(h_new) > h_max) h_new = posneg * h_max;
            if (s.reject) h_new = posneg * fminf(fabsf(h_new), fabsf(h));
            s.reject = false;
            return true;
        }
        h_new = h / fminf(facc1, fac11 / c.safety_factor);
        s.reject = true;
        return false;
    }

    /* hinit, dopri5.rs:187-243 / dop853.rs:195-251 */
    ODE_DEV static float hinit(const KernelArgsF32& a, float x, const float (&y)[N], const float (&p)[NPP], float expo) {
        const ode_b200_config_f32& c = a.cfg;
        float f0[N], f1[N], y1[N];
        rhs(x, y, f0, p);
        const float posneg = signf_(1.0f, c.x_end - x);
        const float h_max_abs = fabsf(c.h_max);
        float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const float sci = c.atol + fabsf(y[i]) * c.rtol;
            d0 += (y[i] / sci) * (y[i] / sci);
            d1 += (f0[i] / sci) * (f0[i] / sci);
        }
        float h0 = (d0 < F32C(1.0e-10) || d1 < F32C(1.0e-10)) ? F32C(1.0e-6) : F32C(0.01) * sqrtf(d0 / d1);
        h0 = fminf(h0, h_max_abs);
        h0 = signf_(h0, posneg);
#pragma unroll
        for (int i = 0; i < N; ++i) y1[i] = y[i] + f0[i] * h0;
        rhs(x + h0, y1, f1, p);
        float d2 = 0.0f;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const float sci = c.atol + fabsf(y[i]) * c.rtol;
            d2 += ((f1[i] - f0[i]) / sci) * ((f1[i] - f0[i]) / sci);
        }
        d2 = sqrtf(d2) / h0;
        float h1;
        if (fmaxf(sqrtf(d1), fabsf(d2)) <= F32C(1.0E-15))
            h1 = fmaxf(F32C(1.0E-6), fabsf(h0) * F32C(1.0E-3));
        else
            h1 = powf_f_(F32C(0.01) / fmaxf(sqrtf(d1), d2), expo); /* second d2 without abs, as in the source */
        return signf_(fminf(F32C(100.0) * fabsf(h0), fminf(h1, h_max_abs)), posneg);
    }

    ODE_DEV static void init(State& s, const KernelArgsF32& a, long long t) {
        const ode_b200_config_f32& c = a.cfg;
        s.traj = t;
#pragma unroll
        for (int i = 0; i < N; ++i) s.y[i] = __ldg(a.y0 + (size_t)i * a.n_traj + t);
#pragma unroll
        for (int i = 0; i < Sys::NP; ++i)
            s.p[i] = a.params_per_traj ? __ldg(a.params + (size_t)i * a.n_traj + t) : __ldg(a.params + i);
        if constexpr (Sys::NP == 0) s.p[0] = 0.0f;
#pragma unroll
        for (int i = 0; i < N; ++i) s.ylast[i] = 0.0f;
        s.x = s.x_old = s.xd = c.x0;
        counters_init_<false>(s.c);
        s.fac_old = F32C(1.0E-4); /* controller.rs:43 */
        s.reject = false;
        s.iasti = s.non_stiff = 0;
        s.status = ODE_B200_OK;
        s.steps_left = 0;
        s.h = s.h_old = 0.0f;
        if (METHOD != ODE_B200_RK4 && c.n_stiff == 0) return; /* `accepted % n_stiff` panics in the reference */
        if constexpr (METHOD == ODE_B200_RK4) { /* rk4.rs:71-77 */
            push(a, s, s.x, s.y);
            const float ns = ceilf((c.x_end - s.x) / c.step_size);
            if (!(ns >= 0.0f) || ns > F32C(4.0e9)) { /* to_usize().unwrap() panics in the reference */
                s.status = ODE_B200_UNSUPPORTED_DOMAIN;
                s.steps_left = 0;
            } else {
                s.steps_left = (unsigned long long)ns;
            }
            s.h_old = c.step_size;
        } else {
            s.h = c.h0;
            if (s.h == 0.0f) {
                s.h = hinit(a, s.x, s.y, s.p, METHOD == ODE_B200_DOPRI5 ? 1.0f / F32C(5.0) : F32C(0.125));
                s.c.num_eval += 2;
            }
            s.h_old = s.h;
            if constexpr (METHOD == ODE_B200_DOPRI5) { /* dopri5.rs:288-296 */
                if (c.out_type == ODE_B200_OUT_SPARSE) push(a, s, s.x, s.y);
            } else { /* solution_output(y0), dop853.rs:273-274 -> :515-519 / :541 */
                push(a, s, c.x0, s.y);
                if (c.out_type == ODE_B200_OUT_DENSE) s.xd += c.dx;
            }
            rhs(s.x, s.y, s.k0, s.p);
            s.c.num_eval += 1;
        }
    }

    /* ------------------------------------------------------------------ Rk4::step, rk4.rs:79-132 */
    ODE_DEV static bool attempt_rk4(State& s, const KernelArgsF32& a) {
        if (s.steps_left == 0) return finish(a, s, s.status);
        const float step = a.cfg.step_size, half = step / F32C(2.0), x = s.x;
        float k0[N], k1[N], k2[N], k3[N], buf[N], yn[N];
        rhs(x, s.y, k0, s.p);
#pragma unroll
        for (int i = 0; i < N; ++i) buf[i] = s.y[i] + k0[i] * half;
        rhs(x + half, buf, k1, s.p);
#pragma unroll
        for (int i = 0; i < N; ++i) buf[i] = s.y[i] + k1[i] * half;
        rhs(x + half, buf, k2, s.p);
#pragma unroll
        for (int i = 0; i < N; ++i) buf[i] = s.y[i] + k2[i] * step;
        rhs(x + step, buf, k3, s.p);
        const float x_new = x + step;
        const float h6 = step / F32C(6.0);
#pragma unroll
        for (int i = 0; i < N; ++i) yn[i] = s.y[i] + (((k0[i] + k1[i] * F32C(2.0)) + k2[i] * F32C(2.0)) + k3[i]) * h6;
        bool stop = false;
        if constexpr (Sys::HAS_SOLOUT) stop = Sys::solout(x_new, yn, k0, s.p);
        s.x = x_new;
#pragma unroll
        for (int i = 0; i < N; ++i) s.y[i] = yn[i];
        push(a, s, x_new, yn);
        s.c.num_eval += 4;
        s.c.accepted += 1;
        s.c.n_step += 1;
        s.steps_left -= 1;
        if (stop) s.steps_left = 0;
        if (s.steps_left == 0) return finish(a, s, s.status);
        return false;
    }

    /* ------------------------------------------------------------------ Dopri5, dopri5.rs:299-487 */
    ODE_DEV static void dp5_y_out(const float (&rc)[5][N], float th, float th1, float (&o)[N]) {
#pragma unroll
        for (int i = 0; i < N; ++i)
            o[i] = rc[0][i] + (rc[1][i] + (rc[2][i] + (rc[3][i] + rc[4][i] * th1) * th) * th1) * th;
    }
    ODE_DEV static bool attempt_dp5(State& s, const KernelArgsF32& a) {
        const ode_b200_config_f32& c = a.cfg;
        const bool dense = c.out_type == ODE_B200_OUT_DENSE, cont = c.out_type == ODE_B200_OUT_CONTINUOUS;
        const float posneg = signf_(1.0f, c.x_end - c.x0);
        if (s.c.n_step > c.n_max) { /* '>' (Q2) */
            s.h_old = s.h;
            return finish(a, s, ODE_B200_MAX_NUM_STEP_REACHED);
        }
        if (F32C(0.1) * fabsf(s.h) <= c.uround * fabsf(s.x)) {
            s.h_old = s.h;
            return finish(a, s, ODE_B200_STEP_SIZE_UNDERFLOW);
        }
        bool last = false;
        if ((s.x + F32C(1.01) * s.h - c.x_end) * posneg > 0.0f) {
            s.h = c.x_end - s.x;
            last = true;
        }
        s.c.n_step += 1;
        const float h = s.h, x = s.x;
        float k[7][N], y_next[N], y_stiff[N], rc[5][N], tmp[N];
#pragma unroll
        for (int i = 0; i < N; ++i) k[0][i] = s.k0[i];
#pragma unroll
        for (int sg = 1; sg < 7; ++sg) { /* dopri5.rs:326-340, every term */
#pragma unroll
            for (int i = 0; i < N; ++i) y_next[i] = s.y[i];
#pragma unroll
            for (int j = 0; j < sg; ++j)
#pragma unroll
                for (int i = 0; i < N; ++i) y_next[i] = y_next[i] + (k[j][i] * h) * (float)c_dp5_a[sg][j];
            rhs(x + h * (float)c_dp5_c[sg], y_next, k[sg], s.p);
            if (sg == 5) {
#pragma unroll
                for (int i = 0; i < N; ++i) y_stiff[i] = y_next[i];
            }
        }
#pragma unroll
        for (int i = 0; i < N; ++i) k[1][i] = k[6][i];
        s.c.num_eval += 6;
        if (dense || cont) { /* every attempt (Q4) */
#pragma unroll
            for (int i = 0; i < N; ++i)
                rc[4][i] = (((((k[0][i] * (float)c_dp5_d[0] + k[2][i] * (float)c_dp5_d[2]) + k[3][i] * (float)c_dp5_d[3]) +
                              k[4][i] * (float)c_dp5_d[4]) +
                             k[5][i] * (float)c_dp5_d[5]) +
                            k[1][i] * (float)c_dp5_d[6]) *
                           h;
        }
#pragma unroll
        for (int i = 0; i < N; ++i)
            tmp[i] = (((((k[0][i] * (float)c_dp5_e[0] + k[2][i] * (float)c_dp5_e[2]) + k[3][i] * (float)c_dp5_e[3]) +
                        k[4][i] * (float)c_dp5_e[4]) +
                       k[5][i] * (float)c_dp5_e[5]) +
                      k[1][i] * (float)c_dp5_e[6]) *
                     h;
#pragma unroll
        for (int i = 0; i < N; ++i) k[3][i] = tmp[i];
        float err = 0.0f;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const float sc = c.atol + fmaxf(fabsf(s.y[i]), fabsf(y_next[i])) * c.rtol;
            err += (k[3][i] / sc) * (k[3][i] / sc);
        }
        err = sqrtf(err / (float)N);

        float h_new;
        if (ctrl_accept(a, s, err, h, h_new)) {
            s.c.accepted += 1;
            if (s.c.accepted % c.n_stiff == 0 || s.iasti > 0) { /* dopri5.rs:378-402 */
                const float num = diff_dot_f_<N>(k[1], k[5]);
                const float den = diff_dot_f_<N>(y_next, y_stiff);
                const float h_lamb = den > 0.0f ? h * sqrtf(num / den) : 0.0f;
                if (h_lamb > F32C(3.25)) {
                    s.iasti += 1;
                    s.non_stiff = 0;
                    if (s.iasti == 15) {
                        s.h_old = h;
                        return finish(a, s, ODE_B200_STIFFNESS_DETECTED);
                    }
                } else {
                    s.non_stiff += 1;
                    if (s.non_stiff == 6) s.iasti = 0;
                }
            }
            if (dense || cont) { /* dopri5.rs:405-414 */
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const float ydiff = y_next[i] - s.y[i];
                    const float bspl = k[0][i] * h - ydiff;
                    rc[0][i] = s.y[i];
                    rc[1][i] = ydiff;
                    rc[2][i] = bspl;
                    rc[3][i] = ((-k[1][i]) * h + ydiff) - bspl;
                }
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                s.k0[i] = k[1][i];
                s.y[i] = y_next[i];
            }
            s.x_old = x;
            s.x = x + h;
            s.h_old = h;
            /* solution_output, dopri5.rs:447-487 */
            if (dense) {
                bool bad = false;
                unsigned guard = 0;
                while (fabsf(s.xd) <= fabsf(s.x)) {
                    bool pushed = false;
                    if (fabsf(s.x_old) <= fabsf(s.xd) && fabsf(s.x) >= fabsf(s.xd)) {
                        const float th = (s.xd - s.x_old) / s.h_old, th1 = 1.0f - th;
                        dp5_y_out(rc, th, th1, tmp);
                        push(a, s, s.xd, tmp);
                        s.xd += c.dx;
                        pushed = true;
                    }
                    if (s.c.out_count == 0) { bad = true; break; } /* reference: unwrap() on empty */
                    if constexpr (Sys::HAS_SOLOUT) {
                        if (Sys::solout(s.xd, s.ylast, s.k0, s.p)) break;
                    }
                    if (!pushed || c.dx == 0.0f || ++guard > (1u << 26)) { bad = true; break; } /* reference never returns */
                }
                if (bad) return finish(a, s, ODE_B200_UNSUPPORTED_DOMAIN);
                if (fabsf(s.xd - c.x_end) < F32C(1.0e-9)) { /* dopri5.rs:472-478 (Q9) */
                    const float th = (c.x_end - s.x_old) / s.h_old, th1 = 1.0f - th;
                    dp5_y_out(rc, th, th1, tmp);
                    push(a, s, c.x_end, tmp);
                    s.xd += c.dx;
                }
            } else {
                push(a, s, s.x, s.y);
            }
            if (cont) add_interval(a, s, fabsf(s.x), rc, s.h_old);
            if (s.c.out_count == 0) return finish(a, s, ODE_B200_UNSUPPORTED_DOMAIN);
            if constexpr (Sys::HAS_SOLOUT) {
                if (Sys::solout(s.x, s.ylast, s.k0, s.p)) last = true;
            }
            if (last) {
                s.h_old = posneg * h_new;
                return finish(a, s, ODE_B200_OK);
            }
        } else {
            if (s.c.accepted >= 1) s.c.rejected += 1; /* Q5 */
        }
        s.h = h_new;
        return false;
    }

    /* ------------------------------------------------------------------ Dop853, dop853.rs:281-543 */
    ODE_DEV static void dp8_y_out(const float (&rc)[8][N], float th, float th1, float (&o)[N]) {
#pragma unroll
        for (int i = 0; i < N; ++i)
            o[i] = rc[0][i] +
                   (rc[1][i] +
                    (rc[2][i] +
                     (rc[3][i] + (rc[4][i] + (rc[5][i] + (rc[6][i] + rc[7][i] * th) * th1) * th) * th1) * th) *
                        th1) *
                       th;
    }
    ODE_DEV static bool attempt_dp8(State& s, const KernelArgsF32& a) {
        const ode_b200_config_f32& c = a.cfg;
        const bool dense = c.out_type == ODE_B200_OUT_DENSE;
        const float posneg = signf_(1.0f, c.x_end - c.x0);
        if (s.c.n_step >= c.n_max) { /* '>=' (Q2) */
            s.h_old = s.h;
            return finish(a, s, ODE_B200_MAX_NUM_STEP_REACHED);
        }
        if (F32C(0.1) * fabsf(s.h) <= c.uround * fabsf(s.x)) {
            s.h_old = s.h;
            return finish(a, s, ODE_B200_STEP_SIZE_UNDERFLOW);
        }
        bool last = false;
        if ((s.x + F32C(1.01) * s.h - c.x_end) * posneg > 0.0f) {
            s.h = c.x_end - s.x;
            last = true;
        }
        s.c.n_step += 1;
        const float h = s.h, x = s.x;
        float k[12][N], y_next[N], err_est[N], err_bhh[N], t1[N];
#pragma unroll
        for (int i = 0; i < N; ++i) k[0][i] = s.k0[i];
#pragma unroll
        for (int sg = 1; sg < 12; ++sg) { /* dop853.rs:307-321, every term */
#pragma unroll
            for (int i = 0; i < N; ++i) y_next[i] = s.y[i];
#pragma unroll
            for (int j = 0; j < sg; ++j) {
                const float ha = h * (float)c_dp8_a[sg][j];
#pragma unroll
                for (int i = 0; i < N; ++i) y_next[i] = y_next[i] + k[j][i] * ha;
            }
            rhs(x + (h * (float)c_dp8_c[sg]), y_next, k[sg], s.p);
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            k[1][i] = k[10][i];
            k[2][i] = k[11][i];
        }
        s.c.num_eval += 11;
#pragma unroll
        for (int i = 0; i < N; ++i)
            t1[i] = ((((((k[0][i] * (float)c_dp8_b[0] + k[5][i] * (float)c_dp8_b[5]) + k[6][i] * (float)c_dp8_b[6]) +
                        k[7][i] * (float)c_dp8_b[7]) +
                       k[8][i] * (float)c_dp8_b[8]) +
                      k[9][i] * (float)c_dp8_b[9]) +
                     k[1][i] * (float)c_dp8_b[10]) +
                    k[2][i] * (float)c_dp8_b[11];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            k[3][i] = t1[i];
            k[4][i] = s.y[i] + k[3][i] * h;
            err_est[i] = 0.0f;
        }
#pragma unroll
        for (int q = 0; q < 12; ++q) /* dop853.rs:334-337: all 12 slots, zero weights included */
#pragma unroll
            for (int i = 0; i < N; ++i) err_est[i] = err_est[i] + k[q][i] * (float)c_dp8_e[q];
        float err = 0.0f, err2 = 0.0f;
#pragma unroll
        for (int i = 0; i < N; ++i)
            err_bhh[i] = ((k[3][i] - k[0][i] * (float)c_dp8_bhh[0]) - k[8][i] * (float)c_dp8_bhh[1]) -
                         k[2][i] * (float)c_dp8_bhh[2];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const float sc = c.atol + fmaxf(fabsf(s.y[i]), fabsf(k[4][i])) * c.rtol;
            err += (err_est[i] / sc) * (err_est[i] / sc);
            err2 += (err_bhh[i] / sc) * (err_bhh[i] / sc);
        }
        float deno = err + F32C(0.01) * err2;
        if (deno <= 0.0f) deno = 1.0f;
        err = fabsf(h) * err * sqrtf(1.0f / (deno * (float)N));

        float h_new;
        if (ctrl_accept(a, s, err, h, h_new)) {
            s.c.accepted += 1;
            float y8[N];
#pragma unroll
            for (int i = 0; i < N; ++i) y8[i] = k[4][i];
            rhs(x + h, y8, k[3], s.p);
            s.c.num_eval += 1;
            if (s.c.accepted % c.n_stiff == 0 || s.iasti > 0) { /* dop853.rs:372-396 */
                const float num = diff_dot_f_<N>(k[3], k[2]);
                const float den = diff_dot_f_<N>(k[4], y_next);
                const float h_lamb = den > 0.0f ? h * sqrtf(num / den) : 0.0f;
                if (h_lamb > F32C(6.1)) {
                    s.non_stiff = 0;
                    s.iasti += 1;
                    if (s.iasti == 15) {
                        s.h_old = h;
                        return finish(a, s, ODE_B200_STIFFNESS_DETECTED);
                    }
                } else {
                    s.non_stiff += 1;
                    if (s.non_stiff == 6) s.iasti = 0;
                }
            }
            float rc[8][N];
            if (dense) { /* dop853.rs:398-480 */
                float yn[N];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    rc[0][i] = s.y[i];
                    const float y_diff = k[4][i] - s.y[i];
                    rc[1][i] = y_diff;
                    const float bspl = k[0][i] * h - y_diff;
                    rc[2][i] = bspl;
                    rc[3][i] = (y_diff - k[3][i] * h) - bspl;
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
#pragma unroll
                    for (int i = 0; i < N; ++i) rc[4 + r][i] = 0.0f;
#pragma unroll
                    for (int j = 0; j < 12; ++j)
#pragma unroll
                        for (int i = 0; i < N; ++i) rc[4 + r][i] = rc[4 + r][i] + k[j][i] * (float)c_dp8_d[r][j];
                }
#pragma unroll
                for (int i = 0; i < N; ++i)
                    yn[i] = s.y[i] + (((((((k[0][i] * (float)c_dp8_a[13][0] + k[6][i] * (float)c_dp8_a[13][6]) +
                                           k[7][i] * (float)c_dp8_a[13][7]) +
                                          k[8][i] * (float)c_dp8_a[13][8]) +
                                         k[9][i] * (float)c_dp8_a[13][9]) +
                                        k[1][i] * (float)c_dp8_a[13][10]) +
                                       k[2][i] * (float)c_dp8_a[13][11]) +
                                      k[3][i] * (float)c_dp8_a[13][12]) *
                                         h;
                rhs(x + h * (float)c_dp8_c[13], yn, k[9], s.p);
#pragma unroll
                for (int i = 0; i < N; ++i)
                    yn[i] = s.y[i] + (((((((k[0][i] * (float)c_dp8_a[14][0] + k[5][i] * (float)c_dp8_a[14][5]) +
                                           k[6][i] * (float)c_dp8_a[14][6]) +
                                          k[7][i] * (float)c_dp8_a[14][7]) +
                                         k[1][i] * (float)c_dp8_a[14][10]) +
                                        k[2][i] * (float)c_dp8_a[14][11]) +
                                       k[3][i] * (float)c_dp8_a[14][12]) +
                                      k[9][i] * (float)c_dp8_a[14][13]) *
                                         h;
                rhs(x + h * (float)c_dp8_c[14], yn, k[1], s.p);
#pragma unroll
                for (int i = 0; i < N; ++i)
                    yn[i] = s.y[i] + (((((((k[0][i] * (float)c_dp8_a[15][0] + k[5][i] * (float)c_dp8_a[15][5]) +
                                           k[6][i] * (float)c_dp8_a[15][6]) +
                                          k[7][i] * (float)c_dp8_a[15][7]) +
                                         k[8][i] * (float)c_dp8_a[15][8]) +
                                        k[3][i] * (float)c_dp8_a[15][12]) +
                                       k[9][i] * (float)c_dp8_a[15][13]) +
                                      k[1][i] * (float)c_dp8_a[15][14]) *
                                         h;
                rhs(x + h * (float)c_dp8_c[15], yn, k[2], s.p);
                s.c.num_eval += 3;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int i = 0; i < N; ++i)
                        rc[4 + r][i] = ((((rc[4 + r][i] + k[3][i] * (float)c_dp8_d[r][12]) + k[9][i] * (float)c_dp8_d[r][13]) +
                                         k[1][i] * (float)c_dp8_d[r][14]) +
                                        k[2][i] * (float)c_dp8_d[r][15]) *
                                       h;
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                s.k0[i] = k[3][i];
                s.y[i] = k[4][i];
            }
            s.x_old = x;
            s.x = x + h;
            s.h_old = h;
            /* solution_output(k[4]), dop853.rs:515-543 */
            if (dense) {
                if (fabsf(s.xd - c.x0) < F32C(2.220446049250313e-16)) {
                    push(a, s, c.x0, s.y);
                    s.xd += c.dx;
                } else {
                    bool bad = false;
                    unsigned guard = 0;
                    float yo[N];
                    while (fabsf(s.xd) <= fabsf(s.x)) {
                        if (fabsf(s.x_old) <= fabsf(s.xd) && fabsf(s.x) >= fabsf(s.xd)) {
                            const float th = (s.xd - s.x_old) / s.h_old, th1 = 1.0f - th;
                            dp8_y_out(rc, th, th1, yo);
                            push(a, s, s.xd, yo);
                            s.xd += c.dx;
                            if (++guard > (1u << 26)) { bad = true; break; }
                        } else { /* the reference never leaves this loop */
                            bad = true;
                            break;
                        }
                    }
                    if (bad) return finish(a, s, ODE_B200_UNSUPPORTED_DOMAIN);
                    if (fabsf(s.xd - c.x_end) < F32C(1.0e-9)) {
                        const float th = (c.x_end - s.x_old) / s.h_old, th1 = 1.0f - th;
                        dp8_y_out(rc, th, th1, yo);
                        push(a, s, c.x_end, yo);
                        s.xd += c.dx;
                    }
                }
            } else {
                push(a, s, c.x0, s.y); /* pushes x0, not x (Q6) */
            }
            if (s.c.out_count == 0) return finish(a, s, ODE_B200_UNSUPPORTED_DOMAIN);
            if constexpr (Sys::HAS_SOLOUT) {
                if (Sys::solout(s.x, s.ylast, s.k0, s.p)) last = true;
            }
            if (last) {
                s.h_old = posneg * h_new;
                return finish(a, s, ODE_B200_OK);
            }
        } else {
            if (s.c.accepted >= 1) s.c.rejected += 1;
        }
        s.h = h_new;
        return false;
    }

    ODE_DEV static bool attempt(State& s, const KernelArgsF32& a) {
        if (METHOD != ODE_B200_RK4 && a.cfg.n_stiff == 0) return finish(a, s, ODE_B200_UNSUPPORTED_DOMAIN);
        if constexpr (METHOD == ODE_B200_RK4)
            return attempt_rk4(s, a);
        else if constexpr (METHOD == ODE_B200_DOPRI5)
            return attempt_dp5(s, a);
        else
            return attempt_dp8(s, a);
    }
};

template <class Sys, int METHOD>
__global__ void __launch_bounds__(128, 4) ode_f32_step_loop_kernel(const __grid_constant__ KernelArgsF32 a) {
    stage_powf_tables_();
    run_lanes_<F32Stepper<Sys, METHOD>>(a);
}

#undef F32C

}  // namespace ode_b200
