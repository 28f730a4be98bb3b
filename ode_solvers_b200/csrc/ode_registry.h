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
    /* Rk4 with T = f32 (nullptr when the system's RHS is f64-only) */
    const void* rk4_f32_kernel;
};

}  // namespace ode_b200
