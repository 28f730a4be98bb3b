/*
 * ode_split.cuh -- Dopri5 and Dop853 for LARGE states (n = 18): one trajectory per GROUP OF THREE LANES.
 *
 * One trajectory per thread keeps k[12][n] in registers; for n = 18 that is 432 registers of stage
 * values alone, and the round-1 kernel streamed them through local memory (ncu: 20 GB of DRAM writes
 * per launch for 85 MB of results, FP64 pipe 30 % busy). Every component of every stage sum, of the
 * 8th-order solution, of the error estimators and of the dense-output coefficients depends on that
 * component only (dop853.rs:307-347, 398-480), so the components of a trajectory are dealt out to three
 * adjacent lanes (six components each: exactly the register footprint of the n = 6 kernels, no local
 * memory) and the lanes meet only where the reference couples components:
 *   - the right-hand side: a split system (SplitThreeBody18 below) evaluates its own part of f and
 *     exchanges what the others need with warp shuffles;
 *   - the three norms of hinit, the two error norms and the stiffness quotients, which the reference
 *     accumulates over the components IN INDEX ORDER (dop853.rs:349-356): the running sum is handed
 *     from lane to lane in that order, so every lane of the group ends up with the reference's bits
 *     and takes the same accept / reject decision on its own.
 * Lanes 30 and 31 of a warp stay idle (10 trajectories per warp).
 *
 * Shuffles need every named lane, so all code that contains one is executed by the WHOLE warp
 * (full-mask shuffles compile to a bare SHFL; a per-group mask costs a MATCH + VOTE + branch each and
 * splits the straight-line block): a pass is split into per-lane phases without shuffles (the
 * controller's commit, result stores, the dense sampling loop) and warp-uniform phases (the attempt
 * itself, the prologue of newly claimed trajectories, the three extra dense stages) in which lanes
 * without work compute on stale values and commit nothing.
 *
 * Bit parity: per component the operations and their order are those of Dopri5Stepper / Dop853Stepper
 * (ode_kernels.cuh), which the parity tests hold bit for bit against the oracle; the only new
 * arithmetic is the order of the cross-component sums, reproduced exactly as described above.
 */
#pragma once

#include "ode_kernels.cuh"
#include "ode_systems.cuh"

namespace ode_b200 {

constexpr unsigned kFullMask = 0xffffffffu;

/* a lane's place in its group of three */
struct Lane3 {
    unsigned role, base, next, prev;
    bool valid;
    ODE_DEV void init() {
        const unsigned lane = threadIdx.x & 31u;
        role = lane % 3u;
        base = lane - role;
        valid = lane < 30u;
        next = base + (role + 1u) % 3u; /* lanes 30/31 name lanes that do not exist: SHFL wraps, nobody reads it */
        prev = base + (role + 2u) % 3u;
    }
};

ODE_DEV double shfl_d_(double v, unsigned src) { return __shfl_sync(kFullMask, v, (int)src); }
/* true for every lane of a group in which any lane holds true */
ODE_DEV bool group_any_(const Lane3& L, bool b) { return ((__ballot_sync(kFullMask, b) >> L.base) & 7u) != 0u; }

/* sum of squares in nalgebra's `dot` order (see diff_dot_, ode_device.cuh) of an already formed difference */
template <int N>
ODE_DEV double sq_dot_(const double (&d)[N]) {
    static_assert(N > 4, "small vectors use diff_dot_");
    double res = 0.0, acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
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

/*
 * SysThreeBody18 (ode_systems.cuh; order of operations as defined there and in the test oracle,
 * threebody18_rhs) dealt out to three lanes. Lane `role` owns body `role`: local components
 * 0..2 = its position (global 3 role + c), 3..5 = its velocity (global 9 + 3 role + c), and evaluates
 * ONE of the three pairs, the one of its own body and the NEXT lane's: role 0 the pair (0,1), role 1
 * the pair (1,2), role 2 the pair (2,0). The oracle visits the pairs (a,b) in the order (0,1),(0,2),(1,2)
 * with d = r_b - r_a, inv = 1 / (|d|^2 |d|), A = (p[b] inv) d, B = (p[a] inv) d and accumulates, per body,
 *     body 0: (0 + A01) + A02      body 1: (0 - B01) + A12      body 2: (0 - B02) - B12.
 * With e = r_next - r_own (= d for roles 0 and 1, = -d for role 2: squares, |d| and inv do not see the
 * sign, and (f inv)(-d) = -((f inv) d) exactly) every lane computes
 *     own  = (p[role + 1] inv) e      ( A01 | A12 | -B02 )   the term of its own body
 *     send = (p[role]     inv) e      ( B01 | B12 | -A02 )   the term the NEXT lane's body needs
 * receives the previous lane's `send` as recv, and every body's acceleration is (0 + own) - recv:
 *     body 0: (0 + A01) - (-A02)     body 1: (0 + A12) - B01     body 2: (0 + (-B02)) - B12.
 * Bit for bit the oracle's values: a - b is a + (-b); and (0 + u) + w == (0 + w) + u for all u, w (IEEE
 * addition commutes, and adding +0 first only turns a -0 operand into +0, after which either order gives
 * the same sum, also when both are zeros), which covers body 1's swapped order. No lane-dependent
 * selection is left: the lanes differ only in the two parameters they were handed at the prologue.
 */
struct SplitThreeBody18 {
    static constexpr int G = 3, NS = 6, FULL = 18, NP = 3;
    static constexpr bool HAS_SOLOUT = false;
    ODE_DEV static int gcomp(unsigned role, int c) { return c < 3 ? 3 * (int)role + c : 9 + 3 * (int)role + (c - 3); }
    /* the index order of the full state as (owner, first local component) runs of three */
    static constexpr int NSEG = 6, SEGLEN = 3;
    ODE_DEV static constexpr unsigned seg_owner(int s) { return (unsigned)(s % 3); }
    ODE_DEV static constexpr int seg_lo(int s) { return (s / 3) * 3; }
    /* the two parameters a lane needs: pl[0] = p[role + 1] (own term), pl[1] = p[role] (the term it hands on) */
    ODE_DEV static int par_a(unsigned role) { return (int)((role + 1u) % 3u); }
    ODE_DEV static int par_b(unsigned role) { return (int)role; }

    template <class M>
    ODE_DEV static void rhs(const Lane3& L, double, const double (&y)[NS], double (&dy)[NS], const double (&pl)[2], M& m) {
        const double ex = shfl_d_(y[0], L.next) - y[0];
        const double ey = shfl_d_(y[1], L.next) - y[1];
        const double ez = shfl_d_(y[2], L.next) - y[2];
        const double rr = ex * ex + ey * ey + ez * ez;
        const double r = m.sqrt_nz(rr);
        const double inv = m.div_nz(1.0, rr * r);
        const double fo = pl[0] * inv, fs = pl[1] * inv;
        const double e[3] = {ex, ey, ez};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double own = fo * e[c];
            const double recv = shfl_d_(fs * e[c], L.prev);
            dy[c] = y[3 + c];
            dy[3 + c] = (0.0 + own) - recv;
        }
    }
};

/* sum over the FULL state in index order of per-component terms held by their owners: s = (..((0 + t_0) + t_1)..)
 * handed from owner to owner; every lane of the group returns the same bits */
template <class SS>
ODE_DEV void ordered_sum2_(const Lane3& L, const double (&t1)[SS::NS], const double (&t2)[SS::NS], double& s1, double& s2) {
    s1 = 0.0;
    s2 = 0.0;
    static_for<0, SS::NSEG>([&](auto S) {
        constexpr int sg = decltype(S)::value;
        constexpr int lo = SS::seg_lo(sg);
        double a = s1, b = s2;
#pragma unroll
        for (int i = 0; i < SS::SEGLEN; ++i) {
            a += t1[lo + i];
            b += t2[lo + i];
        }
        s1 = shfl_d_(a, L.base + SS::seg_owner(sg));
        s2 = shfl_d_(b, L.base + SS::seg_owner(sg));
    });
}
template <class SS>
ODE_DEV double ordered_sum_(const Lane3& L, const double (&t)[SS::NS]) {
    double s = 0.0;
    static_for<0, SS::NSEG>([&](auto S) {
        constexpr int sg = decltype(S)::value;
        constexpr int lo = SS::seg_lo(sg);
        double a = s;
#pragma unroll
        for (int i = 0; i < SS::SEGLEN; ++i) a += t[lo + i];
        s = shfl_d_(a, L.base + SS::seg_owner(sg));
    });
    return s;
}
/* every lane gets the full vector (index order) of a distributed one: the rare stiffness test only */
template <class SS>
ODE_DEV void gather_full_(const Lane3& L, const double (&loc)[SS::NS], double (&full)[SS::FULL]) {
    static_for<0, SS::NSEG>([&](auto S) {
        constexpr int sg = decltype(S)::value;
#pragma unroll
        for (int i = 0; i < SS::SEGLEN; ++i)
            full[sg * SS::SEGLEN + i] = shfl_d_(loc[SS::seg_lo(sg) + i], L.base + SS::seg_owner(sg));
    });
}

/* ============================================================================ shared by the split steppers */
/* State, result stores (direct or through the trajectory's shared-memory slot and the TMA engine), the checked
 * right-hand side and hinit; M = coefficient rows per ContinuousOutputModel interval (Dopri5: 5, Dop853: 8) */
template <class SS, bool STAGE, int M>
struct SplitCommon {
    static constexpr int N = SS::NS, NF = SS::FULL;
    static constexpr bool kStage = STAGE;
    /* STAGE: recorded results leave through shared memory and the TMA engine. Every trajectory slot of
     * the CTA owns M + 1 rows of NF doubles (the coefficient rows of an interval + 1 sample row): the three lanes
     * write their components into the rows (the layout of the result arrays), and the group's first lane
     * hands each block to the TMA engine as ONE bulk copy (Dop853: 1152 B of coefficients, 144 B of state) instead
     * of 54 scattered 8-byte stores per lane (ncu, direct stores: 16.3 GB written at 750 GB/s with
     * long-scoreboard and LSU-queue stalls on top). */
    static constexpr int kSlotDoubles = (M + 1) * NF;
    static constexpr int kSlotsPerWarp = 10;
    static_assert(!SS::HAS_SOLOUT, "solout of a split system is not implemented");

    struct State {
        double y[N], k0[N];
        double pl[2];
        double x, h, x_old, h_old, xd;
        CtrlState ctrl;
        Counters c;
        unsigned iasti, non_stiff, stiff_ctr;
        long long traj;
    };

    /* ---- result stores (per lane; the three lanes of a group always come here together) */
    ODE_DEV static double* slot_(const Lane3& L) {
        double* b = s_stage + ((threadIdx.x >> 5) * kSlotsPerWarp + L.base / 3u) * kSlotDoubles;
        chk_smem_(b, kSlotDoubles);
        return b;
    }
    /* the slot's previous bulk copies have been read by the TMA engine (the leader issued them, so it waits),
     * and the other two lanes learn it at the group barrier */
    ODE_DEV static void slot_acquire_(const Lane3& L) {
        if (L.role == 0u) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp(7u << L.base);
    }
    ODE_DEV static void slot_release_(const Lane3& L) { /* all three lanes have written their components */
        __syncwarp(7u << L.base);
        if (L.role == 0u) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    ODE_DEV static void bulk_(double* dst, const double* src_smem, int bytes) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr_(src_smem)),
                     "r"(bytes)
                     : "memory");
    }
    ODE_DEV static void push(const KernelArgs& a, State& s, const Lane3& L, double x, const double (&y)[N]) {
        if (a.out.out_x != nullptr && (long long)s.c.out_count < a.out.out_cap) {
            const size_t q = (size_t)s.traj * (size_t)a.out_pitch + s.c.out_count;
            if (L.role == 0u) a.out.out_x[q] = x;
            if constexpr (STAGE) {
                double* b = slot_(L) + M * NF;
                slot_acquire_(L);
#pragma unroll
                for (int c = 0; c < N; ++c) b[SS::gcomp(L.role, c)] = y[c];
                slot_release_(L);
                if (L.role == 0u) {
                    chk_row_((long long)s.c.out_count, 1, a.out_pitch);
                    bulk_(a.out.out_y + q * NF, b, NF * 8);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else {
                double* dst = a.out.out_y + q * NF;
#pragma unroll
                for (int c = 0; c < N; ++c) dst[SS::gcomp(L.role, c)] = y[c];
            }
        }
        ++s.c.out_count;
    }
    /* Continuous8: the sample (x, y) and the interval (|x|, rcont[0..7], h_old) of one accepted step */
    ODE_DEV static void push_step(const KernelArgs& a, State& s, const Lane3& L, double x, const double (&y)[N],
                                  double brk, double step, const double (&rc)[M][N]) {
        const ode_b200_outputs& o = a.out;
        const bool rec_y = o.out_x != nullptr && (long long)s.c.out_count < o.out_cap;
        const bool rec_c = o.com_break != nullptr && (long long)s.c.com_count < o.com_cap;
        const size_t qy = (size_t)s.traj * (size_t)a.out_pitch + s.c.out_count;
        const size_t qc = (size_t)s.traj * (size_t)a.com_pitch + s.c.com_count;
        if (L.role == 0u) {
            if (rec_y) o.out_x[qy] = x;
            if (rec_c) {
                o.com_break[qc] = brk;
                o.com_step[qc] = step;
            }
        }
        if constexpr (STAGE) {
            if (rec_y || rec_c) {
                double* b = slot_(L);
                slot_acquire_(L);
#pragma unroll
                for (int c = 0; c < N; ++c) b[M * NF + SS::gcomp(L.role, c)] = y[c];
                if (rec_c) {
#pragma unroll
                    for (int r = 0; r < M; ++r)
#pragma unroll
                        for (int c = 0; c < N; ++c) b[r * NF + SS::gcomp(L.role, c)] = rc[r][c];
                }
                slot_release_(L);
                if (L.role == 0u) {
                    if (rec_c) bulk_(o.com_coef + qc * (M * NF), b, M * NF * 8);
                    if (rec_y) bulk_(o.out_y + qy * NF, b + M * NF, NF * 8);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        } else {
            if (rec_y) {
#pragma unroll
                for (int c = 0; c < N; ++c) o.out_y[qy * NF + SS::gcomp(L.role, c)] = y[c];
            }
            if (rec_c) {
#pragma unroll
                for (int r = 0; r < M; ++r)
#pragma unroll
                    for (int c = 0; c < N; ++c) o.com_coef[(qc * M + r) * NF + SS::gcomp(L.role, c)] = rc[r][c];
            }
        }
        s.c.out_count += 1;
        s.c.com_count += 1;
    }
    ODE_DEV static void finish_(State& s, const KernelArgs& a, const Lane3& L, int status, double com_lower) {
        const ode_b200_outputs& o = a.out;
        const long long t = s.traj;
        if (status == ODE_B200_OK && ((o.out_cap > 0 && (long long)s.c.out_count > o.out_cap) ||
                                      (o.com_cap > 0 && (long long)s.c.com_count > o.com_cap)))
            status = ODE_B200_OUTPUT_CAPACITY_EXCEEDED;
        if (o.y_final)
#pragma unroll
            for (int c = 0; c < N; ++c) o.y_final[(size_t)SS::gcomp(L.role, c) * a.n_traj + t] = s.y[c];
        if (L.role == 0u) {
            if (o.x_final) o.x_final[t] = s.x;
            if (o.h_next) o.h_next[t] = s.h_old;
            if (o.num_eval) o.num_eval[t] = s.c.num_eval;
            if (o.accepted) o.accepted[t] = s.c.accepted;
            if (o.rejected) o.rejected[t] = s.c.rejected;
            if (o.n_step) o.n_step[t] = s.c.n_step;
            if (o.status) o.status[t] = status;
            if (o.out_count) o.out_count[t] = s.c.out_count;
            if (o.com_count) o.com_count[t] = s.c.com_count;
            if (o.com_lower) o.com_lower[t] = com_lower;
            /* the slot goes to the next trajectory only after the TMA engine has read it */
            if constexpr (STAGE) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }

    /* ---- right-hand side with its own check, executed by the whole warp: `use` marks the lanes whose
     * result will be used (only their flags may force the exact recomputation) */
    struct RhsIn {
        double y[N], pl[2], x;
    };
    static __device__ __noinline__ void rhs_exact(const Lane3* L, const RhsIn* in, double* dy) {
        MathExact m;
        double d[N];
        SS::rhs(*L, in->x, in->y, d, in->pl, m);
#pragma unroll
        for (int i = 0; i < N; ++i) dy[i] = d[i];
    }
    ODE_DEV static void rhs_checked(const Lane3& L, bool use, double x, const double (&y)[N], double (&dy)[N],
                                    const double (&pl)[2]) {
        MathFast m;
        SS::rhs(L, x, y, dy, pl, m);
        if (__builtin_expect(__any_sync(kFullMask, use && m.bad()), 0)) {
            RhsIn in;
            double d[N];
#pragma unroll
            for (int i = 0; i < N; ++i) in.y[i] = y[i];
            in.pl[0] = pl[0];
            in.pl[1] = pl[1];
            in.x = x;
            rhs_exact(&L, &in, d);
#pragma unroll
            for (int i = 0; i < N; ++i) dy[i] = d[i];
        }
    }

    /* ---- hinit (dopri5.rs:187-243, dop853.rs:195-251), norms in index order over the full state; whole warp */
    ODE_DEV static double hinit(const KernelArgs& a, const Lane3& L, double x, const double (&y)[N], const double (&pl)[2],
                                double expo) {
        double f0[N], f1[N], y1[N], t0[N], t1[N];
        MathExact mx;
        SS::rhs(L, x, y, f0, pl, mx);
        const double posneg = sign_(1.0, a.cfg.x_end - x);
        const double rtol = a.cfg.rtol, atol = a.cfg.atol;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sci = atol + fabs(y[i]) * rtol;
            const double qy = y[i] / sci, qf = f0[i] / sci;
            t0[i] = qy * qy;
            t1[i] = qf * qf;
        }
        double d0, d1;
        ordered_sum2_<SS>(L, t0, t1, d0, d1);
        double h0 = (d0 < 1.0e-10 || d1 < 1.0e-10) ? 1.0e-6 : 0.01 * sqrt(d0 / d1);
        h0 = fmin(h0, a.h_max_abs);
        h0 = sign_(h0, posneg);
#pragma unroll
        for (int i = 0; i < N; ++i) y1[i] = y[i] + f0[i] * h0;
        SS::rhs(L, x + h0, y1, f1, pl, mx);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double sci = atol + fabs(y[i]) * rtol;
            const double q = (f1[i] - f0[i]) / sci;
            t0[i] = q * q;
        }
        double d2 = ordered_sum_<SS>(L, t0);
        d2 = sqrt(d2) / h0;
        double h1;
        if (fmax(sqrt(d1), fabs(d2)) <= 1.0E-15)
            h1 = fmax(1.0E-6, fabs(h0) * 1.0E-3);
        else
            h1 = pow_d(0.01 / fmax(sqrt(d1), d2), expo); /* second d2 without abs, as in the source */
        return sign_(fmin(100.0 * fabs(h0), fmin(h1, a.h_max_abs)), posneg);
    }

    /* ---- what both prologues share (dopri5.rs:268-296, dop853.rs:255-278): the claimed trajectory's inputs,
     * hinit when h0 == 0, k0 = f(x0, y0); whole warp, only lanes with `ok` commit */
    ODE_DEV static void prologue_(State& s, const KernelArgs& a, long long t, bool ok, const Lane3& L, double expo) {
        const long long tt = ok ? t : 0; /* lanes that commit nothing still load from a valid row */
        double y[N], pl[2], k0[N];
#pragma unroll
        for (int c = 0; c < N; ++c) y[c] = __ldcg(a.y0 + (size_t)SS::gcomp(L.role, c) * a.n_traj + tt);
        const int ia = SS::par_a(L.role), ib = SS::par_b(L.role);
        pl[0] = a.params_per_traj ? __ldcg(a.params + (size_t)ia * a.n_traj + tt) : __ldg(a.params + ia);
        pl[1] = a.params_per_traj ? __ldcg(a.params + (size_t)ib * a.n_traj + tt) : __ldg(a.params + ib);
        const double x = a.cfg.x0;
        double h = a.cfg.h0;
        unsigned num_eval = 0;
        if (h == 0.0) { /* batch-uniform */
            h = hinit(a, L, x, y, pl, expo);
            num_eval += 2;
        }
        rhs_checked(L, ok, x, y, k0, pl);
        num_eval += 1;
        if (ok) {
            s.traj = t;
#pragma unroll
            for (int c = 0; c < N; ++c) {
                s.y[c] = y[c];
                s.k0[c] = k0[c];
            }
            s.pl[0] = pl[0];
            s.pl[1] = pl[1];
            s.x = s.x_old = s.xd = x;
            counters_init_<false>(s.c);
            s.c.num_eval = num_eval;
            ctrl_init_(s.ctrl, a);
            s.iasti = s.non_stiff = s.stiff_ctr = 0;
            s.h = h;
            s.h_old = h;
        }
    }
};

/* ============================================================================ Dop853, three lanes per trajectory */
template <class SS, int OUT, bool STAGE = false>
struct Dop853SplitStepper : SplitCommon<SS, STAGE, 8> {
    using C = SplitCommon<SS, STAGE, 8>;
    using typename C::State;
    using C::N;
    using C::NF;
    using C::kStage;
    using C::kSlotDoubles;
    using C::kSlotsPerWarp;
    using C::slot_;
    using C::slot_acquire_;
    using C::slot_release_;
    using C::bulk_;
    using C::push;
    using C::push_step;
    using C::rhs_checked;
    static constexpr bool DENSE = OUT == ODE_B200_OUT_DENSE;
    static constexpr bool CONT8 = OUT == ODE_B200_OUT_CONTINUOUS8;
    /* one lockstep CTA per SM like the n = 6 kernels: 12 warps at 168 registers, 8 warps with all 255 when the
     * dense-output coefficients are live as well (-D overrides for measurements) */
#ifndef ODE_B200_SPLIT_THREADS_SPARSE
#define ODE_B200_SPLIT_THREADS_SPARSE ODE_B200_LOCK_THREADS
#endif
#ifndef ODE_B200_SPLIT_THREADS_DENSE
#define ODE_B200_SPLIT_THREADS_DENSE 256
#endif
    static constexpr int kThreads = (DENSE || CONT8) ? ODE_B200_SPLIT_THREADS_DENSE : ODE_B200_SPLIT_THREADS_SPARSE;
    static constexpr int kMinBlocks = 1; /* 3 x 128 threads: 6.75 (Sparse) / 17.1 ms (Continuous8: 168 registers spill) */
    static constexpr int kRefillIdle = ODE_B200_REFILL_IDLE;

    ODE_DEV static void finish(State& s, const KernelArgs& a, const Lane3& L, int status) {
        C::finish_(s, a, L, status, CONT8 ? a.cfg.x0 : 0.0);
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

    /* ---- prologue of the trajectories the lanes with `ok` have just claimed (dop853.rs:255-278); whole warp */
    ODE_DEV static void init_all(State& s, const KernelArgs& a, long long t, bool ok, const Lane3& L) {
        C::prologue_(s, a, t, ok, L, 0.125);
        if (ok) {
            /* solution_output(y0), dop853.rs:273-274 -> :515-519 / :541 */
            push(a, s, L, s.x, s.y);
            if constexpr (DENSE) s.xd += a.cfg.dx; /* first call: |xd - x0| = 0 < EPSILON always */
        }
    }

    struct Work {
        double k[12][N]; /* k[0] = f(x, y); k[s] = stage s + 1 */
        double ynext[N]; /* argument of stage 12 (stiffness test) */
        double y8[N];    /* 8th-order solution */
        double knew[N];  /* f(x + h, y8): the reference evaluates it after accepting (dop853.rs:366); here it is
                            part of the warp-uniform block and simply unused by a rejected attempt */
        double err;
        CtrlEval ce;
    };
    struct ColdIn {
        double y[N], k0[N], pl[2], x, h, pw_old;
    };

    /* the straight-line block of one attempt; whole warp. Same statements per component as
     * Dop853Stepper::core (ode_kernels.cuh) */
    template <class M>
    ODE_DEV static void core(const Lane3& L, const double (&y)[N], const double (&k0)[N], const double (&pl)[2],
                             double pw_old, const KernelArgs& a, double x, double h, Work& w, M& m) {
        const ode_b200_config& cfg = a.cfg;
        double(&k)[12][N] = w.k;
        double(&ynext)[N] = w.ynext;
        double(&y8)[N] = w.y8;
        bool bad = false;
#pragma unroll
        for (int i = 0; i < N; ++i) k[0][i] = k0[i];
        static_for<1, 12>([&](auto S) {
            constexpr int sg = decltype(S)::value;
#pragma unroll
            for (int i = 0; i < N; ++i) ynext[i] = y[i];
            double hs = h;
            asm volatile("" : "+d"(hs)); /* see Dop853Stepper::core */
            static_for<0, sg>([&](auto J) {
                constexpr int j = decltype(J)::value;
                constexpr double aij = TabDp8::a(sg, j);
                if constexpr (aij != 0.0) {
                    const double ha = hs * c_dp8_a[sg][j];
#pragma unroll
                    for (int i = 0; i < N; ++i) ynext[i] = ynext[i] + k[j][i] * ha;
                }
            });
            SS::rhs(L, x + (h * c_dp8_c[sg]), ynext, k[sg], pl, m);
            if constexpr (sg >= 1 && sg <= 4) {
#pragma unroll
                for (int i = 0; i < N; ++i) bad |= nonfinite_(k[sg][i]);
            }
        });
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
        SS::rhs(L, x + h, y8, w.knew, pl, m);
        double t1[N], t2[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double ee = k[0][i] * c_dp8_e[0];
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
            bad |= nonfinite_(y8[i]);
            const double rsc = m.rcp(sc);
            const double q1 = m.div_by(ee, sc, rsc), q2 = m.div_by(eb, sc, rsc);
            t1[i] = q1 * q1;
            t2[i] = q2 * q2;
        }
        double err, err2;
        ordered_sum2_<SS>(L, t1, t2, err, err2); /* err += ..., err2 += ... over i = 0..n-1 (dop853.rs:349-356) */
        double deno = err + 0.01 * err2;
        if (deno <= 0.0) deno = 1.0;
        err = fabs(h) * err * m.sqrt_nz(m.div_nz(1.0, deno * (double)NF));
        if (group_any_(L, bad)) err = qnan_();
        w.err = err;
        ctrl_eval_(pw_old, a, err, h, m, w.ce);
    }

    static __device__ __noinline__ void core_exact(const Lane3* L, const ColdIn* in, const KernelArgs* a, Work* w) {
        MathExact m;
        core<MathExact>(*L, in->y, in->k0, in->pl, in->pw_old, *a, in->x, in->h, *w, m);
    }

    /* the three extra stages of the dense output (dop853.rs:414-454); whole warp */
    struct DenseIn {
        double y[N], pl[2], p15[N], p16[N], yy[N], x, h;
    };
    template <class M>
    ODE_DEV static void dense_stages(const Lane3& L, const double (&y)[N], const double (&pl)[2], const double (&yy0)[N],
                                     const double (&p15)[N], const double (&p16)[N], double x, double h, double (&k14)[N],
                                     double (&k15)[N], double (&k16)[N], M& m) {
        double yy[N];
        SS::rhs(L, x + h * c_dp8_c[13], yy0, k14, pl, m);
#pragma unroll
        for (int i = 0; i < N; ++i) yy[i] = y[i] + (p15[i] + k14[i] * c_dp8_a[14][13]) * h;
        SS::rhs(L, x + h * c_dp8_c[14], yy, k15, pl, m);
#pragma unroll
        for (int i = 0; i < N; ++i) yy[i] = y[i] + ((p16[i] + k14[i] * c_dp8_a[15][13]) + k15[i] * c_dp8_a[15][14]) * h;
        SS::rhs(L, x + h * c_dp8_c[15], yy, k16, pl, m);
    }
    static __device__ __noinline__ void dense_stages_exact(const Lane3* L, const DenseIn* in, double* k14, double* k15,
                                                           double* k16) {
        MathExact m;
        double a14[N], a15[N], a16[N];
        dense_stages<MathExact>(*L, in->y, in->pl, in->yy, in->p15, in->p16, in->x, in->h, a14, a15, a16, m);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            k14[i] = a14[i];
            k15[i] = a15[i];
            k16[i] = a16[i];
        }
    }

    /* one pass of `while !last` (dop853.rs:281-510) for every lane of the warp; `active` lanes hold a trajectory */
    ODE_DEV static void attempt_all(State& s, const KernelArgs& a, bool& active, const Lane3& L) {
        const ode_b200_config& cfg = a.cfg;
        /* ---- per lane: the two abort tests and the clamp of the last step */
        bool go = active, last = false;
        if (active) {
            int fin = -1;
            if (s.c.n_step >= cfg.n_max)
                fin = ODE_B200_MAX_NUM_STEP_REACHED;
            else if (0.1 * fabs(s.h) <= cfg.uround * fabs(s.x))
                fin = ODE_B200_STEP_SIZE_UNDERFLOW;
            if (fin >= 0) {
                s.h_old = s.h;
                finish(s, a, L, fin);
                active = go = false;
            } else {
                if ((s.x + 1.01 * s.h - cfg.x_end) * a.posneg > 0.0) {
                    s.h = cfg.x_end - s.x;
                    last = true;
                }
                s.c.n_step += 1;
            }
        }
        if (!__any_sync(kFullMask, go)) return;
        const double h = s.h, x = s.x;

        /* ---- whole warp: the attempt */
        Work w;
        MathFast m;
        core<MathFast>(L, s.y, s.k0, s.pl, s.ctrl.pw_old, a, x, h, w, m);
        if (__builtin_expect(__any_sync(kFullMask, go && m.bad()), 0)) {
            ColdIn in;
            Work w2;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                in.y[i] = s.y[i];
                in.k0[i] = s.k0[i];
            }
            in.pl[0] = s.pl[0];
            in.pl[1] = s.pl[1];
            in.x = x;
            in.h = h;
            in.pw_old = s.ctrl.pw_old;
            core_exact(&L, &in, &a, &w2);
            w = w2;
        }
        double(&k)[12][N] = w.k;
        double(&y8)[N] = w.y8;
        double(&knew)[N] = w.knew;

        /* ---- per lane: controller */
        double h_new = 0.0;
        bool acc = false;
        if (go) {
            s.c.num_eval += 11;
            acc = ctrl_commit_(s.ctrl, a, w.ce, w.err, h, h_new);
            if (acc) {
                s.c.accepted += 1;
                s.c.num_eval += 1; /* knew */
            } else {
                if (s.c.accepted >= 1) s.c.rejected += 1;
                s.h = h_new;
            }
        }
        /* ---- stiffness detection (dop853.rs:372-396): rare, whole warp when any lane is due */
        bool due = false;
        if (acc) {
            if (++s.stiff_ctr == cfg.n_stiff) s.stiff_ctr = 0;
            due = s.stiff_ctr == 0 || s.iasti > 0;
        }
        if (__builtin_expect(__any_sync(kFullMask, due), 0)) {
            double d1[N], d2[N], f1[NF], f2[NF];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                d1[i] = knew[i] - k[11][i];
                d2[i] = y8[i] - w.ynext[i];
            }
            gather_full_<SS>(L, d1, f1);
            gather_full_<SS>(L, d2, f2);
            if (due) {
                const double num = sq_dot_<NF>(f1), den = sq_dot_<NF>(f2);
                const double h_lamb = den > 0.0 ? h * sqrt(num / den) : 0.0;
                if (h_lamb > 6.1) {
                    s.non_stiff = 0;
                    s.iasti += 1;
                    if (s.iasti == 15) {
                        s.h_old = h;
                        finish(s, a, L, ODE_B200_STIFFNESS_DETECTED);
                        active = acc = false;
                    }
                } else {
                    s.non_stiff += 1;
                    if (s.non_stiff == 6) s.iasti = 0;
                }
            }
        }

        /* ---- dense-output coefficients (dop853.rs:398-480): whole warp when any lane accepted */
        /* Continuous8 with staged results: the coefficient rows go straight into the trajectory's shared-memory
         * slot as they are produced instead of living in 96 registers until the end of the pass (ncu, rows kept
         * in registers: 80 local-memory loads and 66 stores per pass, "long scoreboard" the second largest stall) */
        constexpr bool kStream = CONT8 && STAGE;
        double rc[kStream ? 4 : 8][N]; /* kStream: only the partial sums of rows 4..7 (as rc[0..3]) */
        bool rec_y = false, rec_c = false;
        double* slot = nullptr;
        if constexpr (kStream) {
            if (acc) {
                rec_y = a.out.out_x != nullptr && (long long)s.c.out_count < a.out.out_cap;
                rec_c = a.out.com_break != nullptr && (long long)s.c.com_count < a.out.com_cap;
                if (rec_y || rec_c) {
                    slot = slot_(L);
                    slot_acquire_(L);
                }
            }
        }
        constexpr int kHi = kStream ? 0 : 4; /* where rows 4..7 live in rc */
        if constexpr (DENSE || CONT8) {
            if (__any_sync(kFullMask, acc)) {
                double yy[N], p15[N], p16[N], k14[N], k15[N], k16[N];
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double ydiff = y8[i] - s.y[i];
                    const double bspl = k[0][i] * h - ydiff;
                    const double c3 = (ydiff - knew[i] * h) - bspl;
                    if constexpr (kStream) {
                        if (rec_c) {
                            double* b = slot + SS::gcomp(L.role, i);
                            b[0 * NF] = s.y[i];
                            b[1 * NF] = ydiff;
                            b[2 * NF] = bspl;
                            b[3 * NF] = c3;
                        }
                    } else {
                        rc[0][i] = s.y[i];
                        rc[1][i] = ydiff;
                        rc[2][i] = bspl;
                        rc[3][i] = c3;
                    }
                    const double k0 = k[0][i], k5 = k[5][i], k6 = k[6][i], k7 = k[7][i], k8 = k[8][i], k9 = k[9][i],
                                 k10 = k[10][i], k11 = k[11][i], kn = knew[i];
                    static_for<0, 4>([&](auto R) {
                        constexpr int r = decltype(R)::value;
                        double v = 0.0 + k0 * c_dp8_d[r][0];
                        v = v + k5 * c_dp8_d[r][5];
                        v = v + k6 * c_dp8_d[r][6];
                        v = v + k7 * c_dp8_d[r][7];
                        v = v + k8 * c_dp8_d[r][8];
                        v = v + k9 * c_dp8_d[r][9];
                        v = v + k10 * c_dp8_d[r][10];
                        v = v + k11 * c_dp8_d[r][11];
                        rc[kHi + r][i] = v;
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
                MathFast md;
                dense_stages<MathFast>(L, s.y, s.pl, yy, p15, p16, x, h, k14, k15, k16, md);
                if (__builtin_expect(__any_sync(kFullMask, acc && md.bad()), 0)) {
                    DenseIn in;
                    double e14[N], e15[N], e16[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        in.y[i] = s.y[i];
                        in.p15[i] = p15[i];
                        in.p16[i] = p16[i];
                        in.yy[i] = yy[i];
                    }
                    in.pl[0] = s.pl[0];
                    in.pl[1] = s.pl[1];
                    in.x = x;
                    in.h = h;
                    dense_stages_exact(&L, &in, e14, e15, e16);
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        k14[i] = e14[i];
                        k15[i] = e15[i];
                        k16[i] = e16[i];
                    }
                }
                if (acc) s.c.num_eval += 3;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double kn = knew[i], ka = k14[i], kb = k15[i], kc = k16[i];
                    static_for<0, 4>([&](auto R) {
                        constexpr int r = decltype(R)::value;
                        const double v = ((((rc[kHi + r][i] + kn * c_dp8_d[r][12]) + ka * c_dp8_d[r][13]) +
                                           kb * c_dp8_d[r][14]) +
                                          kc * c_dp8_d[r][15]) *
                                         h;
                        if constexpr (kStream) {
                            if (rec_c) slot[(4 + r) * NF + SS::gcomp(L.role, i)] = v;
                        } else {
                            rc[4 + r][i] = v;
                        }
                    });
                }
            }
        }

        /* ---- per lane: commit the accepted step, outputs, exit */
        if (acc) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                s.k0[i] = knew[i];
                s.y[i] = y8[i];
            }
            s.x_old = x;
            s.x = x + h;
            s.h_old = h;
            if constexpr (DENSE) { /* solution_output(k[4]), dop853.rs:515-539 */
                double yo[N];
                if (fabs(s.xd - cfg.x0) < 2.220446049250313e-16) {
                    push(a, s, L, cfg.x0, s.y);
                    s.xd += cfg.dx;
                } else {
                    bool unsupported = false;
                    const double rh = div_rcp_(s.h_old);
                    const bool h_ok = div_in_range_(s.h_old);
                    while (fabs(s.xd) <= fabs(s.x)) {
                        if (fabs(s.x_old) <= fabs(s.xd) && fabs(s.x) >= fabs(s.xd)) {
                            const double th = div_with_rcp_(s.xd - s.x_old, s.h_old, rh, h_ok), th1 = 1.0 - th;
                            y_out(rc, th, th1, yo);
                            push(a, s, L, s.xd, yo);
                            s.xd += cfg.dx;
                            if (s.c.out_count > (1u << 26)) { unsupported = true; break; }
                        } else { /* the reference never leaves this loop */
                            unsupported = true;
                            break;
                        }
                    }
                    if (unsupported) {
                        finish(s, a, L, ODE_B200_UNSUPPORTED_DOMAIN);
                        active = false;
                        return;
                    }
                    if (fabs(s.xd - cfg.x_end) < 1.0e-9) {
                        const double th = div_with_rcp_(cfg.x_end - s.x_old, s.h_old, rh, h_ok), th1 = 1.0 - th;
                        y_out(rc, th, th1, yo);
                        push(a, s, L, cfg.x_end, yo);
                        s.xd += cfg.dx;
                    }
                }
            } else if constexpr (kStream) { /* the sample row, then the slot leaves (see push_step) */
                const ode_b200_outputs& o = a.out;
                const size_t qy = (size_t)s.traj * (size_t)a.out_pitch + s.c.out_count;
                const size_t qc = (size_t)s.traj * (size_t)a.com_pitch + s.c.com_count;
                if (rec_y || rec_c) {
#pragma unroll
                    for (int c = 0; c < N; ++c) slot[8 * NF + SS::gcomp(L.role, c)] = s.y[c];
                    slot_release_(L);
                    if (L.role == 0u) {
                        if (rec_y) o.out_x[qy] = s.x;
                        chk_row_((long long)s.c.out_count, rec_y ? 1 : 0, a.out_pitch);
                        chk_row_((long long)s.c.com_count, rec_c ? 1 : 0, a.com_pitch);
                        if (rec_c) {
                            o.com_break[qc] = fabs(s.x);
                            o.com_step[qc] = s.h_old;
                            bulk_(o.com_coef + qc * (8 * NF), slot, 8 * NF * 8);
                        }
                        if (rec_y) bulk_(o.out_y + qy * NF, slot + 8 * NF, NF * 8);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                s.c.out_count += 1;
                s.c.com_count += 1;
            } else if constexpr (CONT8) {
                push_step(a, s, L, s.x, s.y, fabs(s.x), s.h_old, rc);
            } else {
                push(a, s, L, cfg.x0, s.y); /* pushes x0, not x (Q6) */
            }
            if (last) { /* dop853.rs:499-502 */
                s.h_old = a.posneg * h_new;
                finish(s, a, L, ODE_B200_OK);
                active = false;
                return;
            }
            s.h = h_new;
        }
    }
};

/* ============================================================================ Dopri5, three lanes per trajectory */
/* Per component the statements of Dopri5Stepper (ode_kernels.cuh); only the error norm and the stiffness quotient
 * cross the lanes. Dopri5 has no extra dense-output stages, so everything after the attempt's straight-line block
 * except the (rare) stiffness test runs per lane. */
template <class SS, int OUT, bool STAGE = false>
struct Dopri5SplitStepper : SplitCommon<SS, STAGE, 5> {
    using C = SplitCommon<SS, STAGE, 5>;
    using typename C::State;
    using C::N;
    using C::NF;
    using C::kStage;
    using C::kSlotDoubles;
    using C::kSlotsPerWarp;
    using C::push;
    using C::push_step;
    static constexpr bool DENSE = OUT == ODE_B200_OUT_DENSE;
    static constexpr bool CONT = OUT == ODE_B200_OUT_CONTINUOUS;
    /* three lockstep CTAs of 128 threads per SM (168 registers) instead of one of 384: the same 12 warps, barriers
     * among 4 of them (3-body, 2^18: Sparse 10.96 -> 10.68 ms, Continuous 16.43 -> 15.22; 2 x 192: 10.90 / 15.77) */
#ifndef ODE_B200_SPLIT5_THREADS
#define ODE_B200_SPLIT5_THREADS 128
#endif
#ifndef ODE_B200_SPLIT5_MIN_BLOCKS
#define ODE_B200_SPLIT5_MIN_BLOCKS 3
#endif
    static constexpr int kThreads = ODE_B200_SPLIT5_THREADS;
    static constexpr int kMinBlocks = ODE_B200_SPLIT5_MIN_BLOCKS;
    static constexpr int kRefillIdle = ODE_B200_REFILL_IDLE;

    ODE_DEV static void finish(State& s, const KernelArgs& a, const Lane3& L, int status) {
        C::finish_(s, a, L, status, a.cfg.x0);
    }

    /* compute_y_out, dopri5.rs:490-495 */
    ODE_DEV static void y_out(const double (&rc)[5][N], double th, double th1, double (&o)[N]) {
#pragma unroll
        for (int i = 0; i < N; ++i)
            o[i] = rc[0][i] + (rc[1][i] + (rc[2][i] + (rc[3][i] + rc[4][i] * th1) * th) * th1) * th;
    }

    /* ---- prologue (dopri5.rs:268-296): only Sparse records y0 here; whole warp */
    ODE_DEV static void init_all(State& s, const KernelArgs& a, long long t, bool ok, const Lane3& L) {
        C::prologue_(s, a, t, ok, L, 1.0 / 5.0);
        if constexpr (OUT == ODE_B200_OUT_SPARSE) {
            if (ok) push(a, s, L, s.x, s.y);
        }
    }

    struct Work {
        double k[7][N]; /* k[0] = f(x, y); k[s] for s = 1..6, k[6] is the reference's new k[1] */
        double ynext[N];
        double err;
        CtrlEval ce;
    };
    struct ColdIn {
        double y[N], k0[N], pl[2], x, h, pw_old;
    };

    /* the straight-line block of one attempt; whole warp. Same statements per component as Dopri5Stepper::core */
    template <class M>
    ODE_DEV static void core(const Lane3& L, const double (&y)[N], const double (&k0)[N], const double (&pl)[2],
                             double pw_old, const KernelArgs& a, double x, double h, Work& w, M& m) {
        const ode_b200_config& cfg = a.cfg;
        double(&k)[7][N] = w.k;
        double(&ynext)[N] = w.ynext;
        bool bad = false;
#pragma unroll
        for (int i = 0; i < N; ++i) k[0][i] = k0[i];
        static_for<1, 7>([&](auto S) { /* 6 stages, dopri5.rs:326-340 */
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
            SS::rhs(L, x + h * c_dp5_c[sg], ynext, k[sg], pl, m);
            if constexpr (sg == 1) { /* only k2 meets structural zeros (see the header note of ode_kernels.cuh) */
#pragma unroll
                for (int i = 0; i < N; ++i) bad |= nonfinite_(k[sg][i]);
            }
        });
        /* error estimate and norm, dopri5.rs:354-371: err += (e_i / sc_i)^2 over i = 0..n-1 in index order */
        double t[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double ee = (((((k[0][i] * c_dp5_e[0] + k[2][i] * c_dp5_e[2]) + k[3][i] * c_dp5_e[3]) +
                                 k[4][i] * c_dp5_e[4]) +
                                k[5][i] * c_dp5_e[5]) +
                               k[6][i] * c_dp5_e[6]) *
                              h;
            const double sc = cfg.atol + fmax(fabs(y[i]), fabs(ynext[i])) * cfg.rtol;
            const double q = m.div(ee, sc);
            t[i] = q * q;
        }
        double err = ordered_sum_<SS>(L, t);
        err = m.sqrt(m.div(err, (double)NF));
        if (group_any_(L, bad)) err = qnan_();
        w.err = err;
        ctrl_eval_(pw_old, a, err, h, m, w.ce);
    }

    static __device__ __noinline__ void core_exact(const Lane3* L, const ColdIn* in, const KernelArgs* a, Work* w) {
        MathExact m;
        core<MathExact>(*L, in->y, in->k0, in->pl, in->pw_old, *a, in->x, in->h, *w, m);
    }

    /* one pass of `while !last` (dopri5.rs:299-443) for every lane of the warp; `active` lanes hold a trajectory */
    ODE_DEV static void attempt_all(State& s, const KernelArgs& a, bool& active, const Lane3& L) {
        const ode_b200_config& cfg = a.cfg;
        /* ---- per lane: the two abort tests and the clamp of the last step */
        bool go = active, last = false;
        if (active) {
            int fin = -1;
            if (s.c.n_step > cfg.n_max) /* '>' here, '>=' in dop853 (Q2) */
                fin = ODE_B200_MAX_NUM_STEP_REACHED;
            else if (0.1 * fabs(s.h) <= cfg.uround * fabs(s.x))
                fin = ODE_B200_STEP_SIZE_UNDERFLOW;
            if (fin >= 0) {
                s.h_old = s.h;
                finish(s, a, L, fin);
                active = go = false;
            } else {
                if ((s.x + 1.01 * s.h - cfg.x_end) * a.posneg > 0.0) {
                    s.h = cfg.x_end - s.x;
                    last = true;
                }
                s.c.n_step += 1;
            }
        }
        if (!__any_sync(kFullMask, go)) return;
        const double h = s.h, x = s.x;

        /* ---- whole warp: the attempt */
        Work w;
        MathFast m;
        core<MathFast>(L, s.y, s.k0, s.pl, s.ctrl.pw_old, a, x, h, w, m);
        if (__builtin_expect(__any_sync(kFullMask, go && m.bad()), 0)) {
            ColdIn in;
            Work w2;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                in.y[i] = s.y[i];
                in.k0[i] = s.k0[i];
            }
            in.pl[0] = s.pl[0];
            in.pl[1] = s.pl[1];
            in.x = x;
            in.h = h;
            in.pw_old = s.ctrl.pw_old;
            core_exact(&L, &in, &a, &w2);
            w = w2;
        }
        double(&k)[7][N] = w.k;
        double(&ynext)[N] = w.ynext;

        /* ---- per lane: controller */
        double h_new = 0.0;
        bool acc = false;
        if (go) {
            s.c.num_eval += 6;
            acc = ctrl_commit_(s.ctrl, a, w.ce, w.err, h, h_new);
            if (acc) {
                s.c.accepted += 1;
            } else {
                if (s.c.accepted >= 1) s.c.rejected += 1; /* Q5 */
                s.h = h_new;
            }
        }
        /* ---- stiffness detection (dopri5.rs:378-402): rare, whole warp when any lane is due */
        bool due = false;
        if (acc) {
            if (++s.stiff_ctr == cfg.n_stiff) s.stiff_ctr = 0;
            due = s.stiff_ctr == 0 || s.iasti > 0;
        }
        if (__builtin_expect(__any_sync(kFullMask, due), 0)) {
            /* y_stiff = the argument of stage 6, rebuilt from the still-live k1..k5 (see Dopri5Stepper::attempt) */
            double ystiff[N], d1[N], d2[N], f1[NF], f2[NF];
#pragma unroll
            for (int i = 0; i < N; ++i) ystiff[i] = s.y[i];
            static_for<0, 5>([&](auto J) {
                constexpr int j = decltype(J)::value;
                if constexpr (TabDp5::a(5, j) != 0.0) {
#pragma unroll
                    for (int i = 0; i < N; ++i) ystiff[i] = ystiff[i] + (k[j][i] * h) * c_dp5_a[5][j];
                }
            });
#pragma unroll
            for (int i = 0; i < N; ++i) {
                d1[i] = k[6][i] - k[5][i];
                d2[i] = ynext[i] - ystiff[i];
            }
            gather_full_<SS>(L, d1, f1);
            gather_full_<SS>(L, d2, f2);
            if (due) {
                const double num = sq_dot_<NF>(f1), den = sq_dot_<NF>(f2);
                const double h_lamb = den > 0.0 ? h * sqrt(num / den) : 0.0;
                if (h_lamb > 3.25) {
                    s.iasti += 1;
                    s.non_stiff = 0;
                    if (s.iasti == 15) {
                        s.h_old = h;
                        finish(s, a, L, ODE_B200_STIFFNESS_DETECTED);
                        active = acc = false;
                    }
                } else {
                    s.non_stiff += 1;
                    if (s.non_stiff == 6) s.iasti = 0;
                }
            }
        }

        /* ---- per lane: dense coefficients, commit of the accepted step, outputs, exit */
        if (acc) {
            double rc[5][N];
            if constexpr (DENSE || CONT) { /* dopri5.rs:343-351 and :405-414 */
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
            if constexpr (DENSE) { /* solution_output, dopri5.rs:447-487 */
                bool unsupported = false;
                double yo[N];
                const double rh = div_rcp_(s.h_old);
                const bool h_ok = div_in_range_(s.h_old);
                while (fabs(s.xd) <= fabs(s.x)) {
                    bool pushed = false;
                    if (fabs(s.x_old) <= fabs(s.xd) && fabs(s.x) >= fabs(s.xd)) {
                        const double th = div_with_rcp_(s.xd - s.x_old, s.h_old, rh, h_ok), th1 = 1.0 - th;
                        y_out(rc, th, th1, yo);
                        push(a, s, L, s.xd, yo);
                        s.xd += cfg.dx;
                        pushed = true;
                    }
                    if (s.c.out_count == 0) { unsupported = true; break; } /* reference: unwrap() on empty */
                    /* the reference never returns (see Dopri5Stepper::attempt) */
                    if (!pushed || cfg.dx == 0.0 || s.c.out_count > (1u << 26)) { unsupported = true; break; }
                }
                if (!unsupported && fabs(s.xd - cfg.x_end) < 1.0e-9) { /* endpoint rule, dopri5.rs:472-478 (Q9) */
                    const double th = div_with_rcp_(cfg.x_end - s.x_old, s.h_old, rh, h_ok), th1 = 1.0 - th;
                    y_out(rc, th, th1, yo);
                    push(a, s, L, cfg.x_end, yo);
                    s.xd += cfg.dx;
                }
                if (unsupported || s.c.out_count == 0) {
                    finish(s, a, L, ODE_B200_UNSUPPORTED_DOMAIN);
                    active = false;
                    return;
                }
            } else if constexpr (CONT) { /* the sample, then add_interval(|x|, rcont, h_old), dopri5.rs:484-486 */
                push_step(a, s, L, s.x, s.y, fabs(s.x), s.h_old, rc);
            } else {
                push(a, s, L, s.x, s.y);
            }
            if (last) { /* normal exit, dopri5.rs:432-435 */
                s.h_old = a.posneg * h_new;
                finish(s, a, L, ODE_B200_OK);
                active = false;
                return;
            }
            s.h = h_new;
        }
    }
};

/* ---- lane scheduler of the split steppers: run_lanes_ (ode_device.cuh) with groups of three lanes
 * claiming one trajectory, and every shuffle-bearing phase executed by whole warps */
template <class Stepper, bool PIPE = false>
ODE_DEV void run_lanes_split_(const KernelArgs& a) {
    stage_tables_();
    Lane3 L;
    L.init();
    typename Stepper::State st;
    st.traj = -1;
    st.c.out_count = 0;
    st.x = st.h = st.x_old = st.h_old = st.xd = 0.0;
    st.pl[0] = st.pl[1] = 1.0;
#pragma unroll
    for (int i = 0; i < Stepper::N; ++i) st.y[i] = st.k0[i] = 1.0 + (double)i + (double)L.role; /* never read as results */
    ctrl_init_(st.ctrl, a);
    counters_init_<false>(st.c);
    st.iasti = st.non_stiff = st.stiff_ctr = 0;
    bool active = false, exhausted = !L.valid, first = true, refilled = false;
    int idle_acc = 0, idle_prev = 0;
    unsigned pass = 0;
    constexpr int kRefillRamp = ODE_B200_REFILL_RAMP;
    constexpr int group = Stepper::kThreads / 32 * 30; /* working lanes of the CTA */
    constexpr unsigned sync_every = ODE_B200_SYNC_EVERY;
    constexpr unsigned kLeaders = 0x09249249u; /* lanes 0, 3, ..., 27: one per group */
    const unsigned warp = threadIdx.x >> 5, slot = L.base / 3u;
    for (;;) {
        if (pass % sync_every == 0) {
            const int n_active = __syncthreads_count(active);
            if (n_active == 0 && refilled) break; /* the last refill handed out nothing (see run_lanes_) */
            const int idle = group - n_active;
            idle_acc += idle * (int)sync_every;
            refilled = n_active == 0 || (idle_acc >= Stepper::kRefillIdle && kRefillRamp * (idle - idle_prev) <= idle);
            idle_prev = refilled ? 0 : idle;
            if (refilled) idle_acc = 0;
            const bool want = refilled && !active && !exhausted;
            const unsigned need = idle == 0 ? 0u : __ballot_sync(kFullMask, want);
            bool claimed = false;
            if (need) {
                const unsigned need0 = need & kLeaders;
                long long idx;
                if (a.use_queue) {
                    const int leader = __ffs(need0) - 1;
                    unsigned long long b = 0;
                    if ((int)(threadIdx.x & 31u) == leader) b = atomicAdd(a.queue, (unsigned long long)__popc(need0));
                    b = __shfl_sync(kFullMask, b, leader);
                    idx = (long long)b + __popc(need0 & ((1u << L.base) - 1u));
                } else {
                    idx = first ? ((long long)blockIdx.x * (Stepper::kThreads / 32) + warp) * Stepper::kSlotsPerWarp + slot
                                : a.n_traj;
                }
                const bool ok = want && idx < a.n_traj;
                if (want && !ok) exhausted = true;
                if constexpr (PIPE) { /* host pipeline: report the trajectory stored last, wait for the next one's inputs */
                    if (want && st.traj >= 0) {
                        signal_done_<3>(a, st.traj);
                        st.traj = -1;
                    }
                    if (ok) wait_upload_(a, idx);
                }
                Stepper::init_all(st, a, idx, ok, L);
                if (ok) active = claimed = true;
                first = false;
            }
            if (refilled && __syncthreads_or(claimed)) refilled = false; /* see run_lanes_ */
        }
        if (__any_sync(kFullMask, active)) Stepper::attempt_all(st, a, active, L);
        ++pass;
    }
}

template <class Stepper, bool PIPE = false>
__global__ void __launch_bounds__(Stepper::kThreads, Stepper::kMinBlocks) ode_split_step_loop_kernel(const __grid_constant__ KernelArgs a) {
    run_lanes_split_<Stepper, PIPE>(a);
}

}  // namespace ode_b200
