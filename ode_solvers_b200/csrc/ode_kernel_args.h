/*
 * ode_kernel_args.h -- the one POD a step-loop kernel receives (by value, __grid_constant__).
 * Shared by the host launcher (ode_b200.cu), the AOT kernels and NVRTC-compiled kernels.
 */
#pragma once
#include "ode_b200.h"

namespace ode_b200 {

/* threads per CTA of every step-loop kernel (4 warps; the persistent grid is SMs x resident CTAs) */
constexpr int kBlockThreads = 128;

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
};

}  // namespace ode_b200
