/*
 * ode_kernel_args.h -- the one POD a step-loop kernel receives (by value, __grid_constant__).
 * Shared by the host launcher (ode_b200.cu), the AOT kernels and NVRTC-compiled kernels.
 */
#pragma once
#include "ode_b200.h"

namespace ode_b200 {

/* Threads per CTA of a step-loop kernel (the persistent grid is SMs x resident CTAs). Small
 * states run 4-warp CTAs, several per SM. Steppers whose unrolled per-pass code exceeds the SM's
 * instruction cache (Dop853 with n >= 6, everything with n > 8) run ONE 8-warp CTA per SM with all
 * 255 registers per thread and CTA-lockstep passes (run_lanes_, ode_device.cuh). Measured on B200:
 * CR3BP/Dop853/Dense 50.7 ms (4 x 128 threads) -> 38.4 (1 x 256) -> 33.5 (lockstep);
 * 3-body-18/Dop853 22.5 -> 16.8 -> 12.5 (128 or 192 threads: 13.2 / 12.8); Kepler/Dopri5 (n = 6, 19 KB
 * of code per pass) gains nothing from ONE big CTA (90.7 against 87.0 ms with 4 x 128 threads) -- see below.
 * -DODE_B200_LOCKSTEP=0 switches it off for measurements. */
/* Dopri5 with n = 6..8 (Kepler, CR3BP: ~19 KB of code per pass, ncu: instruction-cache hit rate 87 %, "no
 * instruction" 1.0 stalls per issue): lockstep too, but in SMALL CTAs at unchanged occupancy -- 4 CTAs of 128 threads
 * per SM, each keeping its 4 warps within a pass of each other (round 2, Kepler/Dopri5 2^20: 83.6 ms free-running,
 * 78.6 with 4 x 128 lockstep, 80.0 with 2 x 256, 83.3 with 1 x 512, 85.7 with 1 x 384 at 168 registers; barrier every
 * 2 / 4 / 8 passes: 79.6 / 78.6 / 78.7). The n = 3 kernels fit the cache and lose 2 % with it. */
#ifndef ODE_B200_LOCKSTEP_DP5
#define ODE_B200_LOCKSTEP_DP5 1 /* 0: Dopri5 with n = 6..8 free-running (measurement switch) */
#endif
#ifndef ODE_B200_LOCKSTEP_SMALL_MIN_DIM
#define ODE_B200_LOCKSTEP_SMALL_MIN_DIM 6
#endif
__host__ __device__ constexpr bool lockstep_small_(int dim, int method) {
    return ODE_B200_LOCKSTEP_DP5 && method == ODE_B200_DOPRI5 && dim >= ODE_B200_LOCKSTEP_SMALL_MIN_DIM && dim <= 8;
}
__host__ __device__ constexpr bool lockstep_for(int dim, int method) {
#ifdef ODE_B200_LOCKSTEP_ALL /* measurement switch: every adaptive stepper */
    return method != ODE_B200_RK4;
#elif defined(ODE_B200_LOCKSTEP)
    return ODE_B200_LOCKSTEP != 0 && (dim > 8 || (dim >= 6 && method == ODE_B200_DOP853));
#else
    return dim > 8 || (dim >= 6 && method == ODE_B200_DOP853) || lockstep_small_(dim, method);
#endif
}
/* `records_dense`: the stepper keeps dense-output coefficients live (Dense, Continuous, Continuous8).
 * A lockstep CTA is 12 warps at 168 registers where the state allows it (n = 6 Sparse: Kepler/Dop853
 * 39.4 ms with 8 warps / 255 registers, 37.6 with 12 / 168, 39.1 with 16 / 128, 44.2 with 10), and 8
 * warps with all 255 registers otherwise (CR3BP Dense: 32.1 against 35.7 ms with 12 warps; CR3BP Sparse
 * 2^18: 23.2 -> 20.5 with 12; 3-body-18: 12.3 with 8 warps, 21.0 with 12, 30.1 with 16). */
#ifndef ODE_B200_LOCK_THREADS
#define ODE_B200_LOCK_THREADS 384
#endif
#ifndef ODE_B200_LOCK_THREADS_DENSE
#define ODE_B200_LOCK_THREADS_DENSE 256
#endif
#ifndef ODE_B200_FREE_THREADS
#define ODE_B200_FREE_THREADS 128 /* CTA size of the steppers that do not run in lockstep (-D: tuning experiments) */
#endif
__host__ __device__ constexpr int block_threads_for(int dim, int method, bool records_dense) {
    return !lockstep_for(dim, method)    ? ODE_B200_FREE_THREADS
           : lockstep_small_(dim, method) ? 128
           : (dim <= 8 && !records_dense) ? ODE_B200_LOCK_THREADS
                                          : ODE_B200_LOCK_THREADS_DENSE;
}
__host__ __device__ constexpr bool records_dense(int method, int out_type) {
    return method == ODE_B200_DOPRI5 ? (out_type == ODE_B200_OUT_DENSE || out_type == ODE_B200_OUT_CONTINUOUS)
           : method == ODE_B200_DOP853 ? (out_type == ODE_B200_OUT_DENSE || out_type == ODE_B200_OUT_CONTINUOUS8)
                                       : false;
}

/*
 * Output staging. The per-step results are trajectory-major (one trajectory's samples are
 * contiguous, like the crate's Vec), so a warp whose 32 lanes each push one sample writes 32
 * different 128-byte lines with 8-byte pieces (measured: 0.45-0.64 TB/s of useful writes).
 * In the STAGE variant of every kernel a lane collects its samples in a shared-memory ring of two
 * groups of stage_group(dim) samples; the moment a group is complete the lane itself hands its y
 * vectors to the TMA engine as one bulk copy (cp.async.bulk.global.shared::cta: a contiguous,
 * 16-byte aligned range of the trajectory's slice of out_y), writes its x values with two 16-byte
 * stores, and goes on filling the other group -- no warp-level flush in the step loop. tools/micro/bulk_store.cu: 3.5-4.0 TB/s for this pattern on B200 (32 B + 96 B copies),
 * 4.3-4.6 TB/s with whole 8-sample runs, against 1.1-1.4 TB/s for per-lane vector stores.
 * Rk4 (all lanes of a warp record the same steps) goes one step further: the groups of a warp's 32 lanes leave as
 * ONE TMA tensor store per array (KernelArgs::map_y / map_x, push_warp_ in ode_device.cuh): 4.4-4.7 TB/s in the
 * micro-benchmark, 4.59 -> 4.15 ms for 2^18 Lorenz trajectories x 2001 recorded steps.
 * It needs an even row pitch (alignment) and shared memory, so the launcher picks it only for
 * output-heavy calls (rule in ode_b200.cu). Lane buffer: x[R] then y[R][dim], R = 2 groups.
 */
/* Groups of 8 samples for Rk4 (a bulk copy then moves 192 B instead of 96 B; tools/micro/bulk_store.cu: 4.3-4.6 TB/s
 * against 3.5-4.0 for the pattern alone) were measured in round 2 (-DODE_B200_RK4_GROUP=8): 528 B of ring per lane
 * leave 3 instead of 5 CTAs per SM, and the kernel that records every Rk4 step goes 4.63 -> 4.50 ms at 2^18
 * trajectories (3.63 -> 3.73 TB/s) but 2.51 -> 2.59 ms at 2^17: it is bound by instruction issue (ncu: 73 % of the
 * issue slots), not by the size of its copies. Not kept. */
#ifndef ODE_B200_RK4_GROUP
#define ODE_B200_RK4_GROUP 4
#endif
__host__ __device__ constexpr int stage_group(int dim, int method = ODE_B200_DOPRI5) {
    return dim <= 4 ? (method == ODE_B200_RK4 ? ODE_B200_RK4_GROUP : 4) : 2;
}
__host__ __device__ constexpr int stage_samples(int dim, int method = ODE_B200_DOPRI5) { return 2 * stage_group(dim, method); }
/* in doubles; even (16-byte aligned lane buffers), and not a multiple of 32 (shared-memory banks) */
__host__ __device__ constexpr int stage_stride(int dim, int method = ODE_B200_DOPRI5) {
    return stage_samples(dim, method) * (1 + dim) + 2;
}
/* ContinuousOutputModel intervals (breakpoint, step size, m coefficient vectors) are staged the same
 * way when they fit: a ring of two groups of 2 intervals per lane -- break[4], step[4], coef[4][m][dim]
 * -- each group leaving as one bulk copy (coefficients) and two 16-byte stores. 0 = written directly (large dim * m). */
__host__ __device__ constexpr int com_stage_stride(int dim, int m, int threads) {
    return (stage_stride(dim) + 4 * (2 + m * dim)) * threads * 8 <= 110 * 1024 ? 4 * (2 + m * dim) : 0;
}
/* dynamic shared memory of a staging launch; m = coefficient vectors per interval of the call's
 * continuous output model (0: none recorded) */
__host__ __device__ constexpr int stage_smem_doubles(int dim, int threads, int m, int method = ODE_B200_DOPRI5) {
    return threads * (stage_stride(dim, method) + (m > 0 ? com_stage_stride(dim, m, threads) : 0));
}

/* a CUtensorMap (128 bytes, 64-byte aligned) without <cuda.h>: the kernels only pass its address to the TMA engine */
struct alignas(64) TensorMapBlob {
    unsigned long long w[16];
};

struct KernelArgs {
    ode_b200_config cfg;        /* batch-uniform solver parameters (from_param argument list) */
    /* Controller::new's derived fields (controller.rs:29-49), computed once on the host with
     * the same IEEE divisions the device would do */
    double facc1;               /* 1 / fac_min */
    double facc2;               /* 1 / fac_max */
    double h_max_abs;           /* |h_max| */
    double posneg;              /* sign(1, x_end - x) */
    double pw_floor;            /* pow(1e-4, -beta): fac_old.powf(-beta) while fac_old sits at its floor;
                                   host libm == device clone bit for bit, 1.0 for beta == 0 */
    int pow_shared_log;         /* alpha and -beta are in pow's main exponent range (always, for sane configs) */
    long long n_traj;
    const double* y0;           /* SoA [dim][n_traj] */
    const double* params;       /* [n_params] or SoA [n_params][n_traj] */
    int params_per_traj;
    int use_queue;              /* 1 = persistent lanes refill from *queue; 0 = one thread per trajectory */
    unsigned long long* queue;  /* work-queue head (next unclaimed trajectory index) */
    ode_b200_outputs out;       /* device pointers */
    long long com_pitch;        /* intervals per trajectory row of com_break / com_step / com_coef (>= out.com_cap) */
    long long out_pitch;        /* samples per trajectory row of out_x / out_y in device memory (>= out.out_cap;
                                   the host-buffer entry pads its rows to a multiple of 4) */
    /* Host pipeline of the host-buffer entry (NULL = off, e.g. for the device-pointer entry). The batch is cut
     * into chunks of 1 << chunk_shift consecutive trajectories. Upload: the kernel is launched BEFORE its
     * inputs have arrived; a lane that claims a trajectory waits until upload_ready[chunk] is set (a 4-byte
     * copy queued behind the chunk's data on the upload stream). Drain: a lane that has stored its part of a
     * finished trajectory counts it in chunk_done[chunk]; the lane that completes a chunk raises chunk_flag
     * [chunk] in mapped host memory, and the host, which polls the flags, copies that chunk's results back
     * while the rest of the batch is still being integrated. */
    int chunk_shift;
    const unsigned* upload_ready;
    unsigned* chunk_done;
    unsigned* chunk_flag;
    /* Warp-level result stores of the Rk4 STAGE kernels (ode_device.cuh, push_warp_): when the 32 lanes of a warp
     * hold 32 consecutive trajectories and record the same steps (Rk4 without solout: always, except in the last,
     * partial warp of a batch), a finished group of samples of ALL its lanes leaves as ONE TMA tensor store per
     * array -- box = 32 trajectories x 1 group, rows a trajectory pitch apart -- instead of one bulk copy per lane.
     * map_y: out_y as [n_traj][out_pitch / G][G * dim] doubles, map_x: out_x as [n_traj][out_pitch / G][G];
     * tensor_out = 0: maps not set (row pitch not a multiple of the group, driver without the entry point). */
    int tensor_out;
    TensorMapBlob map_y, map_x;
};

/* the f32 kernels' arguments (device pointers) */
struct KernelArgsF32 {
    ode_b200_config_f32 cfg;
    long long n_traj;
    const float* y0;
    const float* params;
    int params_per_traj;
    int use_queue;
    unsigned long long* queue;
    ode_b200_outputs_f32 out;
};

}  // namespace ode_b200
