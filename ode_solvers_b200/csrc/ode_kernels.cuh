/*
 * ode_kernels.cuh -- the hot path: persistent sm_100a step-loop kernels for Rk4, Dopri5 and
 * Dop853, one trajectory per thread, state and stage vectors in registers.
 *
 * What each stepper restates (reference crate ode_solvers 0.6.1):
 *   Rk4Stepper     src/rk4.rs:71-143      (integrate, step, populate_buffer)
 *   Dopri5Stepper  src/dopri5.rs:263-495  (integrate_core, solution_output, compute_y_out)
 *   Dop853Stepper  src/dop853.rs:254-559  (integrate, solution_output, compute_y_out)
 * with Controller::accept (src/controller.rs:52-77) and hinit in ode_device.cuh.
 *
 * Bit-parity rules followed everywhere:
 *  - operation order exactly as the Rust source evaluates it (left-associated nalgebra
 *    expressions, one rounding per operation, no FMA: compiled with -fmad=false);
 *  - tableau coefficients are compile-time immediates; terms whose coefficient is a
 *    structural zero are skipped (x + k*0.0 == x for finite k). To keep the reference's
 *    behaviour when a stage derivative is inf/NaN (there inf*0.0 = NaN poisons the error
 *    estimate and the step is rejected), the values that could otherwise vanish behind skipped
 *    zeros (Dopri5: k2, whose a72 = e2 = d2 = 0; Dop853: k2..k5, whose b and e weights are 0, and
 *    the 8th-order solution, which sits in a slot of the error sum with a zero weight)
 *    are tested with an integer-pipe exponent check and a hit forces err = NaN,
 *    i.e. the same reject + h/facc1 shrink. Every other stage enters the error estimate with
 *    a non-zero weight, so a non-finite value there makes err non-finite by itself;
 *  - identical sub-expressions the source evaluates twice ((e/sc)*(e/sc), k*h) are computed
 *    once: same operands, same rounding, same bits.
 *  - Signed zeros. A skipped term is +-0.0 for a finite stage value, and s + (+-0.0) == s bit for bit
 *    unless s is -0.0 and the term +0.0 (the IEEE sum is then +0.0). The recorded dense coefficients
 *    (Dop853 rcont[4..7]) are seeded with the source's +0.0 and are exact in the sign of zero too; the
 *    error estimates only enter through their squares. What remains is a stage argument: a state
 *    component that is exactly -0.0, whose every non-skipped term so far was -0.0 as well, meeting a
 *    skipped term (Dopri5: a72; Dop853: the zero columns of a) that is +0.0 -- the reference then
 *    hands +0.0 to the right-hand side where this kernel hands -0.0. No built-in system can tell the
 *    two apart (no division by a state component, no atan2/copysign), and the sign is overwritten by
 *    the next non-zero term; tests/test_gpu_edge_cases.py runs ensembles with -0.0 components, forward
 *    and backward in x, bit for bit against the oracle (which adds every term).
 */
#pragma once

#include "ode_device.cuh"

namespace ode_b200 {



/*
 * Resident CTAs per SM the register allocator must leave room for (Stepper::kMinBlocks).
 * The step loop is bound by dependent FP64 latency (ncu: "wait" is the top stall), so for small
 * states warps per scheduler matter more than a few spilled scalars. Measured on B200, 2^20
 * trajectories, kernel ms at 3/4/5/6 CTAs of 128 threads (current code):
 *   Lorenz Dopri5 32.8 / 29.7 / 28.5 / 28.7      Lorenz Dop853 20.9 / 20.4 / 20.8 / 21.3
 *   Robertson Dop853 8.8 / 8.7 / 9.1 / 9.4       Lorenz Dopri5 Dense (2^18) 10.6 / 10.6 / - / 12.3
 *   Kepler Dopri5 (2/3/4 CTAs) 91.2 / 86.8 / 87.0
 * The big steppers (lockstep_for) run one CTA per SM, see ode_kernel_args.h.
 * -DODE_B200_MIN_BLOCKS=k overrides the table for such sweeps.
 */
/*
 * Whether a stepper's per-attempt block runs through MathFast (straight-line, checked once, exact
 * redo out of line) or through the plain IEEE operators (MathExact inline: one branch per division
 * and root). Systems declare SPEC = true when their right-hand side divides or takes roots through
 * the math policy; for those the straight-line block is what lets the stage chains overlap. Dop853's
 * error norm (2n quotients by n shared reciprocals) gains from it for every system; Dopri5 with a
 * division-free right-hand side does not (measured on B200, 2^20 trajectories, kernel ms plain ->
 * MathFast: Lorenz/Dop853 22.3 -> 21.0, Robertson/Dop853 11.2 -> 10.3, Lorenz/Dopri5 28.0 -> 29.0).
 * -DODE_B200_SPEC=0/1 forces it for measurements.
 */
template <class Sys, int METHOD>
constexpr bool spec_math_for() {
#ifdef ODE_B200_SPEC
    return ODE_B200_SPEC != 0;
#else
    return Sys::SPEC || METHOD == ODE_B200_DOP853;
#endif
}

/* idle lane-passes that trigger a refill (run_lanes_): in the 256 lanes of a lockstep CTA / in a free warp */
#ifndef ODE_B200_REFILL_IDLE
#define ODE_B200_REFILL_IDLE 256
#endif
#ifndef ODE_B200_REFILL_IDLE_FREE
#define ODE_B200_REFILL_IDLE_FREE 32
#endif

template <int N, int METHOD, bool DENSE>
constexpr int min_blocks_for() {
#ifdef ODE_B200_MIN_BLOCKS
    return ODE_B200_MIN_BLOCKS;
#else
    if (lockstep_small_(N, METHOD)) return 4; /* small lockstep CTAs, occupancy as before */
    if (lockstep_for(N, METHOD)) return 1;    /* one CTA per SM */
    if (N <= 4) return (DENSE || METHOD == ODE_B200_DOP853) ? 4 : 5; /* rcont[][] / 12 stages live: 96 registers spill */
    return 4;
#endif
}

/* ============================================================================ Rk4 */
template <class Sys, bool STAGE = false>
struct Rk4Stepper {
    static constexpr int N = Sys::DIM;
    static constexpr bool kStage = STAGE;
    static constexpr int kMinBlocks = min_blocks_for<N, ODE_B200_RK4, false>();
    static constexpr bool kSpec = spec_math_for<Sys, ODE_B200_RK4>();
    static constexpr bool kLockstep = lockstep_for(N, ODE_B200_RK4);
    static constexpr int kThreads = block_threads_for(N, ODE_B200_RK4, false);
    static constexpr int kRefillIdle = kLockstep ? ODE_B200_REFILL_IDLE : ODE_B200_REFILL_IDLE_FREE;
    /* every lane of a warp records the same steps (no solout): a warp of 32 consecutive trajectories hands its
     * samples to the TMA engine with ONE tensor store per group and array (push_warp_, ode_device.cuh) */
    static constexpr bool kTensorOut = STAGE && !Sys::HAS_SOLOUT;
    struct State {
        double y[N];
        double p[Sys::NP > 0 ? Sys::NP : 1];
        double x;
        unsigned long long steps_left;
        Counters c;
        long long traj;
        int status;
        bool uni; /* kTensorOut: set by the lane scheduler for a full warp of consecutive trajectories (warp-uniform) */
    };

    ODE_DEV static void push(const KernelArgs& a, State& s, double x, const double (&y)[N]) {
        if constexpr (kTensorOut) {
            if (s.uni) {
                push_warp_<N, ODE_B200_RK4>(a, s.traj, s.c, x, y);
                return;
            }
        }
        push_<N, STAGE, 0, ODE_B200_RK4>(a, s.traj, s.c, x, y);
    }
    ODE_DEV static void finish(const KernelArgs& a, State& s, double step) {
        if constexpr (kTensorOut) {
            if (s.uni) {
                flush_warp_tail_<N, ODE_B200_RK4>(a, s.traj, s.c);
                store_final_<N, false, 1, 0, ODE_B200_RK4>(a, s.traj, s.y, s.x, step, s.c, s.status, 0.0);
                return;
            }
        }
        store_final_<N, STAGE, 1, 0, ODE_B200_RK4>(a, s.traj, s.y, s.x, step, s.c, s.status, 0.0);
    }

    /* Rk4::integrate prologue, rk4.rs:71-77 */
    ODE_DEV static void init(State& s, const KernelArgs& a, long long t) {
        s.traj = t;
        load_traj_<Sys>(a, t, s.y, s.p);
        s.x = a.cfg.x0;
        counters_init_<STAGE>(s.c);
        s.status = ODE_B200_OK;
        if constexpr (!kTensorOut) s.uni = false;
        push(a, s, s.x, s.y);
        const double ns = ceil((a.cfg.x_end - s.x) / a.cfg.step_size);
        if (!(ns >= 0.0) || ns > 4.0e9) { /* to_usize().unwrap() panics in the reference */
            s.status = ODE_B200_UNSUPPORTED_DOMAIN;
            s.steps_left = 0;
        } else {
            s.steps_left = (unsigned long long)ns;
        }
    }

    static constexpr int NPP = Sys::NP > 0 ? Sys::NP : 1;
    struct Work {
        double k0[N], yn[N];
    };
    struct ColdIn {
        double y[N], p[NPP], x, step;
    };

    /* Rk4::step, rk4.rs:103-132: the straight-line block (see MathFast) */
    template <class M>
    ODE_DEV static void core(const double (&y)[N], const double (&p)[NPP], double x, double step, Work& w, M& m) {
        const double half = step / 2.0;
        double k1[N], k2[N], k3[N], buf[N];
        Sys::rhs(x, y, w.k0, p, m);
#pragma unroll
        for (int i = 0; i < N; ++i) buf[i] = y[i] + w.k0[i] * half;
        Sys::rhs(x + half, buf, k1, p, m);
#pragma unroll
        for (int i = 0; i < N; ++i) buf[i] = y[i] + k1[i] * half;
        Sys::rhs(x + half, buf, k2, p, m);
#pragma unroll
        for (int i = 0; i < N; ++i) buf[i] = y[i] + k2[i] * step;
        Sys::rhs(x + step, buf, k3, p, m);
        const double h6 = step / 6.0;
#pragma unroll
        for (int i = 0; i < N; ++i) w.yn[i] = y[i] + (((w.k0[i] + k1[i] * 2.0) + k2[i] * 2.0) + k3[i]) * h6;
    }

    static __device__ __noinline__ void core_exact(const ColdIn* in, Work* w) {
        MathExact m;
        core<MathExact>(in->y, in->p, in->x, in->step, *w, m);
    }

    /* one Rk4::step + the loop body of integrate, rk4.rs:79-98,103-132 */
    ODE_DEV static bool attempt(State& s, const KernelArgs& a) {
        if (s.steps_left == 0) {
            finish(a, s, a.cfg.step_size);
            return true;
        }
        const double step = a.cfg.step_size, x = s.x;
        Work w;
        MathFast m;
        if constexpr (kSpec) {
            core<MathFast>(s.y, s.p, x, step, w, m);
        } else {
            MathExact mx;
            core<MathExact>(s.y, s.p, x, step, w, mx);
        }
        if (__builtin_expect(kSpec && m.bad(), 0)) {
            ColdIn in;
            Work w2;
#pragma unroll
            for (int i = 0; i < N; ++i) in.y[i] = s.y[i];
#pragma unroll
            for (int i = 0; i < NPP; ++i) in.p[i] = s.p[i];
            in.x = x;
            in.step = step;
            core_exact(&in, &w2);
            w = w2;
        }
        const double x_new = x + step;
        bool stop = false;
        if constexpr (Sys::HAS_SOLOUT) stop = Sys::solout(x_new, w.yn, w.k0, s.p);
        s.x = x_new;
#pragma unroll
        for (int i = 0; i < N; ++i) s.y[i] = w.yn[i];
        push(a, s, x_new, w.yn);
        s.c.num_eval += 4;
        s.c.accepted += 1;
        s.c.n_step += 1;
        s.steps_left -= 1;
        if (stop) s.steps_left = 0;
        if (s.steps_left == 0) {
            finish(a, s, step);
            return true;
        }
        return false;
    }
};

/* ============================================================================ Dopri5 */
template <class Sys, int OUT, bool STAGE = false>
struct Dopri5Stepper {
    static constexpr int N = Sys::DIM;
    static constexpr bool kStage = STAGE;
    static constexpr bool DENSE = OUT == ODE_B200_OUT_DENSE;
    static constexpr bool CONT = OUT == ODE_B200_OUT_CONTINUOUS;
    static constexpr bool TRACK_LAST = DENSE && Sys::HAS_SOLOUT;
    /* the staged Dense kernel (>= 1000 samples per trajectory): with the bound relaxed from 4 to 3 CTAs per SM ptxas
     * takes 128 instead of 126 registers (4 CTAs still fit) and stops squeezing the sampling loop: Lorenz, 2001
     * samples, 17.88 -> 16.01 ms (bound 2: 18.07, bound 5 / 96 registers: 20.45); Continuous and the unstaged
     * kernels measure the same either way */
#ifdef ODE_B200_MIN_BLOCKS
    static constexpr int kMinBlocks = ODE_B200_MIN_BLOCKS;
#else
    static constexpr int kMinBlocks = (STAGE && DENSE && N <= 4) ? 3 : min_blocks_for<N, ODE_B200_DOPRI5, DENSE || CONT>();
#endif
    static constexpr bool kSpec = spec_math_for<Sys, ODE_B200_DOPRI5>();
    /* shared-memory ring of ContinuousOutputModel intervals (0: written directly) */
    static constexpr int kCS = (STAGE && CONT) ? com_stage_stride(N, 5, block_threads_for(N, ODE_B200_DOPRI5, true)) : 0;
    static constexpr bool kLockstep = lockstep_for(N, ODE_B200_DOPRI5);
    static constexpr int kThreads = block_threads_for(N, ODE_B200_DOPRI5, DENSE || CONT);
    static constexpr int kRefillIdle = kLockstep ? ODE_B200_REFILL_IDLE : ODE_B200_REFILL_IDLE_FREE;

    struct State {
        double y[N], k0[N]; /* k0 = f(x, y): first-same-as-last slot k[0] */
        double p[Sys::NP > 0 ? Sys::NP : 1];
        double ylast[TRACK_LAST ? N : 1]; /* results.last() (only solout reads it) */
        double x, h, x_old, h_old, xd;
        CtrlState ctrl;
        Counters c;
        unsigned iasti, non_stiff, stiff_ctr;
        long long traj;
    };

    ODE_DEV static bool finish(State& s, const KernelArgs& a, int status) {
        store_final_<N, STAGE, 5, kCS>(a, s.traj, s.y, s.x, s.h_old, s.c, status, a.cfg.x0);
        return true;
    }

    /* integrate_core prologue, dopri5.rs:268-296 */
    ODE_DEV static void init(State& s, const KernelArgs& a, long long t) {
        s.traj = t;
        load_traj_<Sys>(a, t, s.y, s.p);
        s.x = s.x_old = s.xd = a.cfg.x0;
        counters_init_<STAGE>(s.c);
        ctrl_init_(s.ctrl, a);
        s.iasti = s.non_stiff = s.stiff_ctr = 0;
        s.h = a.cfg.h0;
        if (s.h == 0.0) {
            s.h = hinit_<Sys>(a, s.x, s.y, s.p, 1.0 / 5.0);
            s.c.num_eval += 2;
        }
        s.h_old = s.h;
        if constexpr (OUT == ODE_B200_OUT_SPARSE) push_<N, STAGE, kCS>(a, t, s.c, s.x, s.y);
        rhs_checked_<Sys>(s.x, s.y, s.k0, s.p);
        s.c.num_eval += 1;
    }

    /* compute_y_out, dopri5.rs:490-495 */
    ODE_DEV static void y_out(const double (&rc)[5][N], double th, double th1, double (&o)[N]) {
#pragma unroll
        for (int i = 0; i < N; ++i)
            o[i] = rc[0][i] + (rc[1][i] + (rc[2][i] + (rc[3][i] + rc[4][i] * th1) * th) * th1) * th;
    }

    static constexpr int NPP = Sys::NP > 0 ? Sys::NP : 1;
    /* everything one attempted step computes before the accept / reject decision */
    struct Work {
        double k[7][N]; /* k[0] = f(x, y); k[s] for s = 1..6, k[6] is the reference's new k[1] */
        double ynext[N];
        double err;
        CtrlEval ce;
    };
    struct ColdIn {
        double y[N], k0[N], p[NPP], x, h, pw_old;
    };

    /* the straight-line block of one attempt (see MathFast): 6 stages, error norm, controller */
    template <class M>
    ODE_DEV static void core(const double (&y)[N], const double (&k0)[N], const double (&p)[NPP], double pw_old,
                             const KernelArgs& a, double x, double h, Work& w, M& m) {
        const ode_b200_config& cfg = a.cfg;
        double(&k)[7][N] = w.k;
        double(&ynext)[N] = w.ynext;
        /* 6 stages, dopri5.rs:326-340 */
        bool bad = false;
#pragma unroll
        for (int i = 0; i < N; ++i) k[0][i] = k0[i];
        static_for<1, 7>([&](auto S) {
            constexpr int sg = decltype(S)::value;
#pragma unroll
            for (int i = 0; i < N; ++i) ynext[i] = y[i];
            static_for<0, sg>([&](auto J) {
                constexpr int j = decltype(J)::value;
                constexpr double aij = TabDp5::a(sg, j);
                if constexpr (aij != 0.0) {
#pragma unroll
                    for (int i = 0; i < N; ++i) ynext[i] = ynext[i] + (k[j][i] * h) * c_dp5_a[sg][j];
                }
            });
            Sys::rhs(x + h * c_dp5_c[sg], ynext, k[sg], p, m);
            if constexpr (sg == 1) { /* only k2 meets structural zeros (a72, e2, d2): see header note */
#pragma unroll
                for (int i = 0; i < N; ++i) bad |= nonfinite_(k[sg][i]);
            }
        });
        /* error estimate and norm, dopri5.rs:354-371 */
        double err = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double ee = (((((k[0][i] * c_dp5_e[0] + k[2][i] * c_dp5_e[2]) + k[3][i] * c_dp5_e[3]) +
                                 k[4][i] * c_dp5_e[4]) +
                                k[5][i] * c_dp5_e[5]) +
                               k[6][i] * c_dp5_e[6]) *
                              h;
            const double sc = cfg.atol + fmax(fabs(y[i]), fabs(ynext[i])) * cfg.rtol;
            const double q = m.div(ee, sc);
            err += q * q;
        }
        err = m.sqrt(m.div(err, (double)N));
        if (bad) err = qnan_();
        w.err = err;
        ctrl_eval_(pw_old, a, err, h, m, w.ce);
    }

    static __device__ __noinline__ void core_exact(const ColdIn* in, const KernelArgs* a, Work* w) {
        MathExact m;
        core<MathExact>(in->y, in->k0, in->p, in->pw_old, *a, in->x, in->h, *w, m);
    }

    /* one pass of `while !last`, dopri5.rs:299-443 */
    ODE_DEV static bool attempt(State& s, const KernelArgs& a) {
        const ode_b200_config& cfg = a.cfg;
        if (s.c.n_step > cfg.n_max) { /* '>' here, '>=' in dop853 (Q2) */
            s.h_old = s.h;
            return finish(s, a, ODE_B200_MAX_NUM_STEP_REACHED);
        }
        if (0.1 * fabs(s.h) <= cfg.uround * fabs(s.x)) {
            s.h_old = s.h;
            return finish(s, a, ODE_B200_STEP_SIZE_UNDERFLOW);
        }
        bool last = false;
        if ((s.x + 1.01 * s.h - cfg.x_end) * a.posneg > 0.0) {
            s.h = cfg.x_end - s.x;
            last = true;
        }
        s.c.n_step += 1;
        const double h = s.h, x = s.x;

        Work w;
        MathFast m;
        if constexpr (kSpec) {
            core<MathFast>(s.y, s.k0, s.p, s.ctrl.pw_old, a, x, h, w, m);
        } else {
            MathExact mx;
            core<MathExact>(s.y, s.k0, s.p, s.ctrl.pw_old, a, x, h, w, mx);
        }
#ifdef ODE_B200_EXPERIMENT_NOCOLD
        if (__builtin_expect(kSpec && m.bad(), 0)) __trap();
#else
        if (__builtin_expect(kSpec && m.bad(), 0)) { /* some operand left the fast paths' window: redo it exactly */
            ColdIn in;
            Work w2;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                in.y[i] = s.y[i];
                in.k0[i] = s.k0[i];
            }
#pragma unroll
            for (int i = 0; i < NPP; ++i) in.p[i] = s.p[i];
            in.x = x;
            in.h = h;
            in.pw_old = s.ctrl.pw_old;
            core_exact(&in, &a, &w2);
            w = w2;
        }
#endif
        s.c.num_eval += 6;
        double(&k)[7][N] = w.k;
        double(&ynext)[N] = w.ynext;

        double h_new;
        if (ctrl_commit_(s.ctrl, a, w.ce, w.err, h, h_new)) {
            s.c.accepted += 1;
            /* stiffness detection, dopri5.rs:378-402 (accepted % n_stiff == 0 kept as a counter) */
            if (++s.stiff_ctr == cfg.n_stiff) s.stiff_ctr = 0;
            if (s.stiff_ctr == 0 || s.iasti > 0) {
                /* y_stiff = the argument of stage 6 (dopri5.rs:335-337). It is needed once per n_stiff
                 * accepted steps, so it is rebuilt here from the still-live k1..k5 (same operations,
                 * same order) instead of occupying N register pairs in every pass. */
                double ystiff[N];
#pragma unroll
                for (int i = 0; i < N; ++i) ystiff[i] = s.y[i];
                static_for<0, 5>([&](auto J) {
                    constexpr int j = decltype(J)::value;
                    if constexpr (TabDp5::a(5, j) != 0.0) {
#pragma unroll
                        for (int i = 0; i < N; ++i) ystiff[i] = ystiff[i] + (k[j][i] * h) * c_dp5_a[5][j];
                    }
                });
                const double num = diff_dot_<N>(k[6], k[5]);
                const double den = diff_dot_<N>(ynext, ystiff);
                const double h_lamb = den > 0.0 ? h * sqrt(num / den) : 0.0;
                if (h_lamb > 3.25) {
                    s.iasti += 1;
                    s.non_stiff = 0;
                    if (s.iasti == 15) {
                        s.h_old = h;
                        return finish(s, a, ODE_B200_STIFFNESS_DETECTED);
                    }
                } else {
                    s.non_stiff += 1;
                    if (s.non_stiff == 6) s.iasti = 0;
                }
            }
            /* dense / continuous coefficients, dopri5.rs:343-351 (rcont[4], computed by the
             * reference on every attempt but only read after an accept) and :405-414 */
            double rc[5][N];
            if constexpr (DENSE || CONT) {
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    rc[4][i] = (((((k[0][i] * c_dp5_d[0] + k[2][i] * c_dp5_d[2]) + k[3][i] * c_dp5_d[3]) +
                                  k[4][i] * c_dp5_d[4]) +
                                 k[5][i] * c_dp5_d[5]) +
                                k[6][i] * c_dp5_d[6]) *
                               h;
                    const double ydiff = ynext[i] - s.y[i];
                    const double bspl = k[0][i] * h - ydiff;
                    rc[0][i] = s.y[i];
                    rc[1][i] = ydiff;
                    rc[2][i] = bspl;
                    rc[3][i] = ((-k[6][i]) * h + ydiff) - bspl;
                }
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                s.k0[i] = k[6][i];
                s.y[i] = ynext[i];
            }
            s.x_old = x;
            s.x = x + h;
            s.h_old = h;

            /* solution_output, dopri5.rs:447-487 */
            bool stop = false;
            if constexpr (DENSE) {
                bool unsupported = false;
                double yo[N];
                /* every sample of this step divides by the same h_old: one refined reciprocal, exact
                 * quotients (div_with_rcp_, ode_device.cuh) */
                const double rh = div_rcp_(s.h_old);
                const bool h_ok = div_in_range_(s.h_old);
                while (fabs(s.xd) <= fabs(s.x)) {
                    bool pushed = false;
                    if (fabs(s.x_old) <= fabs(s.xd) && fabs(s.x) >= fabs(s.xd)) {
                        const double th = div_with_rcp_(s.xd - s.x_old, s.h_old, rh, h_ok), th1 = 1.0 - th;
                        y_out(rc, th, th1, yo);
                        push_<N, STAGE, kCS>(a, s.traj, s.c, s.xd, yo);
                        if constexpr (TRACK_LAST) {
#pragma unroll
                            for (int i = 0; i < N; ++i) s.ylast[i] = yo[i];
                        }
                        s.xd += cfg.dx;
                        pushed = true;
                    }
                    if (s.c.out_count == 0) { unsupported = true; break; } /* reference: unwrap() on empty */
                    if constexpr (Sys::HAS_SOLOUT) {
                        if (Sys::solout(s.xd, s.ylast, s.k0, s.p)) break;
                    }
                    /* the reference never returns: nothing changes between passes once no sample is
                     * pushed, or when dx == 0 leaves xd where it was */
                    if (!pushed || cfg.dx == 0.0 || s.c.out_count > (1u << 26)) { unsupported = true; break; }
                }
                if (unsupported) return finish(s, a, ODE_B200_UNSUPPORTED_DOMAIN);
                if (fabs(s.xd - cfg.x_end) < 1.0e-9) { /* endpoint rule, dopri5.rs:472-478 (Q9) */
                    const double th = div_with_rcp_(cfg.x_end - s.x_old, s.h_old, rh, h_ok), th1 = 1.0 - th;
                    y_out(rc, th, th1, yo);
                    push_<N, STAGE, kCS>(a, s.traj, s.c, cfg.x_end, yo);
                    if constexpr (TRACK_LAST) {
#pragma unroll
                        for (int i = 0; i < N; ++i) s.ylast[i] = yo[i];
                    }
                    s.xd += cfg.dx;
                }
                if (s.c.out_count == 0) return finish(s, a, ODE_B200_UNSUPPORTED_DOMAIN);
                if constexpr (Sys::HAS_SOLOUT) stop = Sys::solout(s.x, s.ylast, s.k0, s.p);
            } else {
                push_<N, STAGE, kCS>(a, s.traj, s.c, s.x, s.y);
                if constexpr (CONT) /* add_interval(|x|, rcont, h_old), dopri5.rs:484-486 */
                    push_com_<N, 5, kCS>(a, s.traj, s.c, fabs(s.x), s.h_old, rc);
                if constexpr (Sys::HAS_SOLOUT) stop = Sys::solout(s.x, s.y, s.k0, s.p);
            }
            if (stop) last = true;
            if (last) { /* normal exit, dopri5.rs:432-435 */
                s.h_old = a.posneg * h_new;
                return finish(s, a, ODE_B200_OK);
            }
        } else {
            if (s.c.accepted >= 1) s.c.rejected += 1; /* Q5 */
        }
        s.h = h_new;
        return false;
    }
};

/* ============================================================================ Dop853 */
/* OUT: Dense, Sparse (the reference treats Continuous like Sparse: dop853.rs:398,540-542), or the
 * ODE_B200_OUT_CONTINUOUS8 extension (rcont[0..7] recorded as ContinuousOutputModel intervals) */
template <class Sys, int OUT, bool STAGE = false>
struct Dop853Stepper {
    static constexpr int N = Sys::DIM;
    static constexpr bool kStage = STAGE;
    static constexpr bool DENSE = OUT == ODE_B200_OUT_DENSE;
    static constexpr bool CONT8 = OUT == ODE_B200_OUT_CONTINUOUS8;
    static constexpr bool TRACK_LAST = DENSE && Sys::HAS_SOLOUT;
    static constexpr int kMinBlocks = min_blocks_for<N, ODE_B200_DOP853, DENSE || CONT8>();
    static constexpr bool kSpec = spec_math_for<Sys, ODE_B200_DOP853>();
    static constexpr int kCS = (STAGE && CONT8) ? com_stage_stride(N, 8, block_threads_for(N, ODE_B200_DOP853, true)) : 0;
    static constexpr bool kLockstep = lockstep_for(N, ODE_B200_DOP853);
    static constexpr int kThreads = block_threads_for(N, ODE_B200_DOP853, DENSE || CONT8);
    static constexpr int kRefillIdle = kLockstep ? ODE_B200_REFILL_IDLE : ODE_B200_REFILL_IDLE_FREE;

    struct State {
        double y[N], k0[N]; /* k0 = reference k[0] = f(x, y) */
        double p[Sys::NP > 0 ? Sys::NP : 1];
        double ylast[TRACK_LAST ? N : 1];
        double x, h, x_old, h_old, xd;
        CtrlState ctrl;
        Counters c;
        unsigned iasti, non_stiff, stiff_ctr;
        long long traj;
    };

    ODE_DEV static bool finish(State& s, const KernelArgs& a, int status) {
        store_final_<N, STAGE, 8, kCS>(a, s.traj, s.y, s.x, s.h_old, s.c, status, CONT8 ? a.cfg.x0 : 0.0);
        return true;
    }

    /* compute_y_out, dop853.rs:546-559 */
    ODE_DEV static void y_out(const double (&rc)[8][N], double th, double th1, double (&o)[N]) {
#pragma unroll
        for (int i = 0; i < N; ++i)
            o[i] = rc[0][i] +
                   (rc[1][i] +
                    (rc[2][i] +
                     (rc[3][i] + (rc[4][i] + (rc[5][i] + (rc[6][i] + rc[7][i] * th) * th1) * th) * th1) * th) *
                        th1) *
                       th;
    }

    /* integrate prologue, dop853.rs:255-278 */
    ODE_DEV static void init(State& s, const KernelArgs& a, long long t) {
        s.traj = t;
        load_traj_<Sys>(a, t, s.y, s.p);
        s.x = s.x_old = s.xd = a.cfg.x0;
        counters_init_<STAGE>(s.c);
        ctrl_init_(s.ctrl, a);
        s.iasti = s.non_stiff = s.stiff_ctr = 0;
        s.h = a.cfg.h0;
        if (s.h == 0.0) {
            s.h = hinit_<Sys>(a, s.x, s.y, s.p, 0.125);
            s.c.num_eval += 2;
        }
        s.h_old = s.h;
        /* solution_output(y0), dop853.rs:273-274 -> :515-519 / :541 */
        if constexpr (DENSE) {
            /* first call: |xd - x0| = 0 < EPSILON always */
            push_<N, STAGE, kCS>(a, t, s.c, a.cfg.x0, s.y);
            if constexpr (TRACK_LAST) {
#pragma unroll
                for (int i = 0; i < N; ++i) s.ylast[i] = s.y[i];
            }
            s.xd += a.cfg.dx;
        } else {
            push_<N, STAGE, kCS>(a, t, s.c, a.cfg.x0, s.y);
        }
        rhs_checked_<Sys>(s.x, s.y, s.k0, s.p);
        s.c.num_eval += 1;
    }

    static constexpr int NPP = Sys::NP > 0 ? Sys::NP : 1;
    /* everything one attempted step computes before the accept / reject decision */
    struct Work {
        double k[12][N]; /* k[0] = f(x, y); k[s] = stage s+1. Reference slots after the stage loop:
                            k[1] = k11 (our k[10]), k[2] = k12 (our k[11]) */
        double ynext[N]; /* argument of stage 12 (stiffness test) */
        double y8[N];    /* 8th-order solution, reference k[4] */
        double err;
        CtrlEval ce;
    };
    struct ColdIn {
        double y[N], k0[N], p[NPP], x, h, pw_old;
    };

    /* the straight-line block of one attempt (see MathFast): 11 stages, 8th-order solution, the
     * two error estimators, controller */
    template <class M>
    ODE_DEV static void core(const double (&y)[N], const double (&k0)[N], const double (&p)[NPP], double pw_old,
                             const KernelArgs& a, double x, double h, Work& w, M& m) {
        const ode_b200_config& cfg = a.cfg;
        double(&k)[12][N] = w.k;
        double(&ynext)[N] = w.ynext;
        double(&y8)[N] = w.y8;
        /* stages 2..12, dop853.rs:307-321: k[s] = f(x + h*c(s+1), y + sum_j k[j]*(h*a(s+1,j+1))) */
        bool bad = false;
#pragma unroll
        for (int i = 0; i < N; ++i) k[0][i] = k0[i];
        static_for<1, 12>([&](auto S) {
            constexpr int sg = decltype(S)::value;
#pragma unroll
            for (int i = 0; i < N; ++i) ynext[i] = y[i];
            /* For the big steppers the step size is made opaque per stage: the compiler then forms the
             * products h * a(s, j) where the stage uses them instead of hoisting all 66 of them to the top
             * of the block, where they compete with the stage vectors for registers (3-body-18: 12.45 ->
             * 11.4 ms, Kepler 37.6 -> 37.1; n = 3: no difference). Same operands, same products. */
            double hs = h;
            if constexpr (N >= 6) asm volatile("" : "+d"(hs));
            static_for<0, sg>([&](auto J) {
                constexpr int j = decltype(J)::value;
                constexpr double aij = TabDp8::a(sg, j);
                if constexpr (aij != 0.0) {
                    const double ha = hs * c_dp8_a[sg][j];
#pragma unroll
                    for (int i = 0; i < N; ++i) ynext[i] = ynext[i] + k[j][i] * ha;
                }
            });
            Sys::rhs(x + (h * c_dp8_c[sg]), ynext, k[sg], p, m); /* c(12) = 0.0 (Q1) */
            if constexpr (sg >= 1 && sg <= 4) { /* only k2..k5 have b = e = 0: see header note */
#pragma unroll
                for (int i = 0; i < N; ++i) bad |= nonfinite_(k[sg][i]);
            }
        });

        /* 8th-order solution, dop853.rs:323-331 */
        double bsum[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            bsum[i] = ((((((k[0][i] * c_dp8_b[0] + k[5][i] * c_dp8_b[5]) + k[6][i] * c_dp8_b[6]) +
                          k[7][i] * c_dp8_b[7]) +
                         k[8][i] * c_dp8_b[8]) +
                        k[9][i] * c_dp8_b[9]) +
                       k[10][i] * c_dp8_b[10]) +
                      k[11][i] * c_dp8_b[11];
            y8[i] = y[i] + bsum[i] * h;
        }
        /* error estimators and norm, dop853.rs:334-362. err_est runs over the 12 slots; slots
         * 1..4 (k11, k12, bsum, y8) meet e2..e5 = 0 and are skipped. */
        double err = 0.0, err2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double ee = k[0][i] * c_dp8_e[0]; /* 0.0 + v == v */
            ee = ee + k[5][i] * c_dp8_e[5];
            ee = ee + k[6][i] * c_dp8_e[6];
            ee = ee + k[7][i] * c_dp8_e[7];
            ee = ee + k[8][i] * c_dp8_e[8];
            ee = ee + k[9][i] * c_dp8_e[9];
            ee = ee + k[10][i] * c_dp8_e[10];
            ee = ee + k[11][i] * c_dp8_e[11];
            const double eb = ((bsum[i] - k[0][i] * c_dp8_bhh[0]) - k[8][i] * c_dp8_bhh[1]) -
                              k[11][i] * c_dp8_bhh[2];
            const double sc = cfg.atol + fmax(fabs(y[i]), fabs(y8[i])) * cfg.rtol;
            /* the reference's err_est runs over its 12 slots, of which slots 3 and 4 hold bsum and y8 by
             * now, times e4 = e5 = 0 (dop853.rs:334-347): a non-finite y8 (which every non-finite bsum
             * implies) poisons the error estimate there, e.g. for an infinite state component */
            bad |= nonfinite_(y8[i]);
            /* err_est_i / sc_i and err_bhh_i / sc_i: one refined reciprocal, two exact quotients */
            const double rsc = m.rcp(sc);
            const double q1 = m.div_by(ee, sc, rsc), q2 = m.div_by(eb, sc, rsc);
            err += q1 * q1;
            err2 += q2 * q2;
        }
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = fabs(h) * err * m.sqrt_nz(m.div_nz(1.0, deno * (double)N));
        if (bad) err = qnan_();
        w.err = err;
        ctrl_eval_(pw_old, a, err, h, m, w.ce);
    }

    static __device__ __noinline__ void core_exact(const ColdIn* in, const KernelArgs* a, Work* w) {
        MathExact m;
        core<MathExact>(in->y, in->k0, in->p, in->pw_old, *a, in->x, in->h, *w, m);
    }

    /* one pass of `while !last`, dop853.rs:281-510 */
    ODE_DEV static bool attempt(State& s, const KernelArgs& a) {
        const ode_b200_config& cfg = a.cfg;
        if (s.c.n_step >= cfg.n_max) { /* '>=' (Q2) */
            s.h_old = s.h;
            return finish(s, a, ODE_B200_MAX_NUM_STEP_REACHED);
        }
        if (0.1 * fabs(s.h) <= cfg.uround * fabs(s.x)) {
            s.h_old = s.h;
            return finish(s, a, ODE_B200_STEP_SIZE_UNDERFLOW);
        }
        bool last = false;
        if ((s.x + 1.01 * s.h - cfg.x_end) * a.posneg > 0.0) {
            s.h = cfg.x_end - s.x;
            last = true;
        }
        s.c.n_step += 1;
        const double h = s.h, x = s.x;

        Work w;
        MathFast m;
        if constexpr (kSpec) {
            core<MathFast>(s.y, s.k0, s.p, s.ctrl.pw_old, a, x, h, w, m);
        } else {
            MathExact mx;
            core<MathExact>(s.y, s.k0, s.p, s.ctrl.pw_old, a, x, h, w, mx);
        }
        if (__builtin_expect(kSpec && m.bad(), 0)) { /* some operand left the fast paths' window: redo it exactly */
            ColdIn in;
            Work w2;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                in.y[i] = s.y[i];
                in.k0[i] = s.k0[i];
            }
#pragma unroll
            for (int i = 0; i < NPP; ++i) in.p[i] = s.p[i];
            in.x = x;
            in.h = h;
            in.pw_old = s.ctrl.pw_old;
            core_exact(&in, &a, &w2);
            w = w2;
        }
        s.c.num_eval += 11;
        double(&k)[12][N] = w.k;
        double(&ynext)[N] = w.ynext;
        double(&y8)[N] = w.y8;

        double h_new;
        if (ctrl_commit_(s.ctrl, a, w.ce, w.err, h, h_new)) {
            s.c.accepted += 1;
            double knew[N]; /* reference k[3] = f(x + h, k[4]); no FSAL reuse (Q11) */
            rhs_checked_<Sys>(x + h, y8, knew, s.p);
            s.c.num_eval += 1;

            /* stiffness detection, dop853.rs:372-396 */
            if (++s.stiff_ctr == cfg.n_stiff) s.stiff_ctr = 0;
            if (s.stiff_ctr == 0 || s.iasti > 0) {
                const double num = diff_dot_<N>(knew, k[11]);
                const double den = diff_dot_<N>(y8, ynext); /* ynext = argument of stage 12 */
                const double h_lamb = den > 0.0 ? h * sqrt(num / den) : 0.0;
                if (h_lamb > 6.1) {
                    s.non_stiff = 0;
                    s.iasti += 1;
                    if (s.iasti == 15) {
                        s.h_old = h;
                        return finish(s, a, ODE_B200_STIFFNESS_DETECTED);
                    }
                } else {
                    s.non_stiff += 1;
                    if (s.non_stiff == 6) s.iasti = 0;
                }
            }

            double rc[8][N];
            if constexpr (DENSE || CONT8) { /* dop853.rs:398-480 */
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    rc[0][i] = s.y[i];
                    const double ydiff = y8[i] - s.y[i];
                    rc[1][i] = ydiff;
                    const double bspl = k[0][i] * h - ydiff;
                    rc[2][i] = bspl;
                    rc[3][i] = (ydiff - knew[i] * h) - bspl;
                }
                /* One pass per component, so that every stage value is read once (for n = 18 they live in
                 * local memory) for all of:
                 *  - rcont[4..7] = sum_{j=1..12} slot[j-1]*d(i,j): columns 2..5 are zero; slot 10/11 = k11/k12;
                 *  - the arguments of the three extra stages 14, 15, 16 (dop853.rs:414-454): the left-
                 *    associated sums of stages 15 and 16 end with the terms of k14 (and k15), which do not
                 *    exist yet -- p15 / p16 are those sums up to the knew term, completed below in order. */
                double yy[N], p15[N], p16[N], k14[N], k15[N], k16[N];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double k0 = k[0][i], k5 = k[5][i], k6 = k[6][i], k7 = k[7][i], k8 = k[8][i], k9 = k[9][i],
                                 k10 = k[10][i], k11 = k[11][i], kn = knew[i];
                    static_for<0, 4>([&](auto R) {
                        constexpr int r = decltype(R)::value;
                        /* seeded with the source's zero vector (dop853.rs:411): (+0.0) + (-0.0) = +0.0, and
                         * once the sum is not -0.0 the skipped (+-0.0) terms of columns 2..5 cannot change it */
                        double v = 0.0 + k0 * c_dp8_d[r][0];
                        v = v + k5 * c_dp8_d[r][5];
                        v = v + k6 * c_dp8_d[r][6];
                        v = v + k7 * c_dp8_d[r][7];
                        v = v + k8 * c_dp8_d[r][8];
                        v = v + k9 * c_dp8_d[r][9];
                        v = v + k10 * c_dp8_d[r][10];
                        v = v + k11 * c_dp8_d[r][11];
                        rc[4 + r][i] = v;
                    });
                    yy[i] = s.y[i] + (((((((k0 * c_dp8_a[13][0] + k6 * c_dp8_a[13][6]) + k7 * c_dp8_a[13][7]) +
                                           k8 * c_dp8_a[13][8]) +
                                          k9 * c_dp8_a[13][9]) +
                                         k10 * c_dp8_a[13][10]) +
                                        k11 * c_dp8_a[13][11]) +
                                       kn * c_dp8_a[13][12]) *
                                          h;
                    p15[i] = (((((k0 * c_dp8_a[14][0] + k5 * c_dp8_a[14][5]) + k6 * c_dp8_a[14][6]) +
                                k7 * c_dp8_a[14][7]) +
                               k10 * c_dp8_a[14][10]) +
                              k11 * c_dp8_a[14][11]) +
                             kn * c_dp8_a[14][12];
                    p16[i] = ((((k0 * c_dp8_a[15][0] + k5 * c_dp8_a[15][5]) + k6 * c_dp8_a[15][6]) +
                               k7 * c_dp8_a[15][7]) +
                              k8 * c_dp8_a[15][8]) +
                             kn * c_dp8_a[15][12];
                }
                rhs_checked_<Sys>(x + h * c_dp8_c[13], yy, k14, s.p);
#pragma unroll
                for (int i = 0; i < N; ++i) yy[i] = s.y[i] + (p15[i] + k14[i] * c_dp8_a[14][13]) * h;
                rhs_checked_<Sys>(x + h * c_dp8_c[14], yy, k15, s.p);
#pragma unroll
                for (int i = 0; i < N; ++i)
                    yy[i] = s.y[i] + ((p16[i] + k14[i] * c_dp8_a[15][13]) + k15[i] * c_dp8_a[15][14]) * h;
                rhs_checked_<Sys>(x + h * c_dp8_c[15], yy, k16, s.p);
                s.c.num_eval += 3;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double kn = knew[i], ka = k14[i], kb = k15[i], kc = k16[i];
                    static_for<0, 4>([&](auto R) {
                        constexpr int r = decltype(R)::value;
                        rc[4 + r][i] = ((((rc[4 + r][i] + kn * c_dp8_d[r][12]) + ka * c_dp8_d[r][13]) +
                                         kb * c_dp8_d[r][14]) +
                                        kc * c_dp8_d[r][15]) *
                                       h;
                    });
                }
            }
#pragma unroll
            for (int i = 0; i < N; ++i) {
                s.k0[i] = knew[i];
                s.y[i] = y8[i];
            }
            s.x_old = x;
            s.x = x + h;
            s.h_old = h;

            /* solution_output(k[4]), dop853.rs:515-543 */
            if constexpr (DENSE) {
                double yo[N];
                if (fabs(s.xd - cfg.x0) < 2.220446049250313e-16) {
                    push_<N, STAGE, kCS>(a, s.traj, s.c, cfg.x0, s.y);
                    if constexpr (TRACK_LAST) {
#pragma unroll
                        for (int i = 0; i < N; ++i) s.ylast[i] = s.y[i];
                    }
                    s.xd += cfg.dx;
                } else {
                    bool unsupported = false;
                    const double rh = div_rcp_(s.h_old); /* one reciprocal for every sample of the step */
                    const bool h_ok = div_in_range_(s.h_old);
                    while (fabs(s.xd) <= fabs(s.x)) {
                        if (fabs(s.x_old) <= fabs(s.xd) && fabs(s.x) >= fabs(s.xd)) {
                            const double th = div_with_rcp_(s.xd - s.x_old, s.h_old, rh, h_ok), th1 = 1.0 - th;
                            y_out(rc, th, th1, yo);
                            push_<N, STAGE, kCS>(a, s.traj, s.c, s.xd, yo);
                            if constexpr (TRACK_LAST) {
#pragma unroll
                                for (int i = 0; i < N; ++i) s.ylast[i] = yo[i];
                            }
                            s.xd += cfg.dx;
                            if (s.c.out_count > (1u << 26)) { unsupported = true; break; }
                        } else { /* the reference never leaves this loop */
                            unsupported = true;
                            break;
                        }
                    }
                    if (unsupported) return finish(s, a, ODE_B200_UNSUPPORTED_DOMAIN);
                    if (fabs(s.xd - cfg.x_end) < 1.0e-9) {
                        const double th = div_with_rcp_(cfg.x_end - s.x_old, s.h_old, rh, h_ok), th1 = 1.0 - th;
                        y_out(rc, th, th1, yo);
                        push_<N, STAGE, kCS>(a, s.traj, s.c, cfg.x_end, yo);
                        if constexpr (TRACK_LAST) {
#pragma unroll
                            for (int i = 0; i < N; ++i) s.ylast[i] = yo[i];
                        }
                        s.xd += cfg.dx;
                    }
                }
            } else if constexpr (CONT8) {
                /* extension: (x, y) like Dopri5's Continuous push (dopri5.rs:479-481) and
                 * add_interval(|x|, rcont[0..7], h_old) like dopri5.rs:484-486 */
                push_<N, STAGE, kCS>(a, s.traj, s.c, s.x, s.y);
                push_com_<N, 8, kCS>(a, s.traj, s.c, fabs(s.x), s.h_old, rc);
            } else {
                push_<N, STAGE, kCS>(a, s.traj, s.c, cfg.x0, s.y); /* pushes x0, not x (Q6) */
            }

            bool stop = false;
            if constexpr (Sys::HAS_SOLOUT) {
                if constexpr (DENSE)
                    stop = Sys::solout(s.x, s.ylast, s.k0, s.p);
                else
                    stop = Sys::solout(s.x, s.y, s.k0, s.p);
            }
            if (stop) last = true;
            if (last) { /* dop853.rs:499-502 */
                s.h_old = a.posneg * h_new;
                return finish(s, a, ODE_B200_OK);
            }
        } else {
            if (s.c.accepted >= 1) s.c.rejected += 1;
        }
        s.h = h_new;
        return false;
    }
};

/* ============================================================================ kernels */
/* PIPE: the instantiation the host pipeline launches (ode_kernel_args.h; hooks in run_lanes_) */
template <class Stepper, bool PIPE = false>
__global__ void __launch_bounds__(Stepper::kThreads, Stepper::kMinBlocks) ode_step_loop_kernel(const __grid_constant__ KernelArgs a) {
    run_lanes_<Stepper, KernelArgs, PIPE>(a);
}

}  // namespace ode_b200
