/*
 * ode_registry.h -- host-side table of compiled systems: for every (method, out_type) the
 * address of the matching ode_step_loop_kernel instantiation.
 */
#pragma once

namespace ode_b200 {

struct SystemEntry {
    const char* name;
    int dim;
    int n_params;
    int has_solout;
    /* [staged results: no, yes][method: Rk4, Dopri5, Dop853][out_type: Dense, Sparse, Continuous, Continuous8] */
    const void* kernels[2][3][4];
    /* the unstaged kernels once more with the host pipeline's device side compiled in (ode_kernel_args.h) */
    const void* pipe_kernels[3][4];
    /* the steppers with T = f32 by method (nullptr when the system's RHS is f64-only) */
    const void* f32_kernels[3];
    /* Dopri5 ([0]) and Dop853 ([1]) with one trajectory per group of three lanes (ode_split.cuh), by out_type;
     * nullptr: the system has no split form and runs one trajectory per thread */
    struct SplitKernels {
        const void* fn[2][4];   /* [results staged through shared memory: no, yes][out_type] */
        int threads[4];         /* CTA size of each of them */
        int slot_bytes;         /* shared memory per trajectory slot of a staged launch */
        const void* pipe_fn[4]; /* unstaged, with the host pipeline's device side */
    } split[2];
    int split_traj_per_warp; /* 10: lanes 30 and 31 stay idle */
};

}  // namespace ode_b200
