/*
 * ode_oracle.h -- CPU ORACLE. TEST INFRASTRUCTURE ONLY, NOT A PRODUCT PATH.
 *
 * A plain-C, operation-order-faithful restatement of the reference crate's explicit RK
 * step loop (srenevey/ode-solvers v0.6.1, src/{rk4,dopri5,dop853,controller,
 * butcher_tableau,constants,continuous_output_model,dop_shared}.rs). Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product library (libode_b200.so) never links or calls anything in this directory.
 *
 * Parity pin: the Rust crate cannot be built in this image (no cargo/rustc), so there is no
 * oracle/_ref. The oracle is pinned against every golden value of the reference's own
 * tests (tests/test_oracle_golden.py lists them with file:line) -- those pin results to
 * 1e-9 (1e-5 for two Rk4 cases); bit-level agreement with the crate beyond that is
 * UNPINNED by the reference (see DESIGN.md "Oracle"). Third-party arithmetic restated from
 * memory of un-vendored dependencies: nalgebra 0.33.2 `dot` (stiffness test only).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -fopenmp; links glibc libm, whose
 * pow/exp are the functions the crate's f64::powf/exp resolve to on linux-gnu).
 */
#ifndef ODE_ORACLE_H
#define ODE_ORACLE_H

#include <stdint.h>
#include "../include/ode_b200.h" /* shares the POD config / outputs structs and enums only */

#ifdef __cplusplus
extern "C" {
#endif

/* trait System (src/dop_shared.rs:31-42): system() and solout(); struct fields -> p[] */
typedef void (*ode_oracle_rhs_fn)(double x, const double* y, double* dy, const double* p);
typedef int (*ode_oracle_solout_fn)(double x, const double* y, const double* dy, const double* p);

typedef struct ode_oracle_system {
    const char* name;
    int dim;
    int n_params;
    ode_oracle_rhs_fn rhs;
    ode_oracle_solout_fn solout; /* NULL = default `false` (dop_shared.rs:39-41) */
} ode_oracle_system;

/* the same with T = f32, for the systems whose RHS has an f32 form */
typedef void (*ode_oracle_rhs32_fn)(float x, const float* y, float* dy, const float* p);
typedef int (*ode_oracle_solout32_fn)(float x, const float* y, const float* dy, const float* p);
typedef struct ode_oracle_system32 {
    const char* name;
    int dim;
    int n_params;
    ode_oracle_rhs32_fn rhs;
    ode_oracle_solout32_fn solout;
} ode_oracle_system32;
const ode_oracle_system32* ode_oracle_system32_by_name(const char* name);

const ode_oracle_system* ode_oracle_system_by_name(const char* name);
int ode_oracle_system_count(void);
const ode_oracle_system* ode_oracle_system_at(int i);

/* constructors' defaults; identical arithmetic to the product's ode_b200_config_* helpers
 * but implemented independently here */
int ode_oracle_config_new(int method, double x, double x_end, double dx, double rtol, double atol,
                          ode_b200_config* out);
int ode_oracle_config_from_param(int method, double x, double x_end, double dx, double rtol, double atol,
                                 double safety_factor, double beta, double fac_min, double fac_max,
                                 double h_max, double h, uint32_t n_max, uint32_t n_stiff, int out_type,
                                 ode_b200_config* out);
int ode_oracle_config_rk4(double x, double x_end, double step_size, ode_b200_config* out);

/*
 * Integrate n_traj trajectories on the CPU (OpenMP over trajectories, n_threads <= 0 = all
 * cores). Same argument meaning and output layouts as ode_b200_integrate_batch.
 * Returns 0, or -1 on bad arguments.
 */
int ode_oracle_integrate_batch(const char* system, const ode_b200_config* cfg, int64_t n_traj,
                               const double* y0, const double* params, int params_per_traj,
                               const ode_b200_outputs* out, int n_threads);

/* The three steppers with T = f32 (the crate is generic over FloatNumber, src/dop_shared.rs:52-77): the same
 * source (ode_oracle_core.inc) instantiated for float, libm powf/sqrtf. Constructors like the f64 ones with
 * every literal rounded to f32 and every operation in f32 (Q3: Dopri5::new takes uround = f32::EPSILON). */
int ode_oracle_config_new_f32(int method, float x, float x_end, float dx, float rtol, float atol,
                              ode_b200_config_f32* out);
int ode_oracle_config_from_param_f32(int method, float x, float x_end, float dx, float rtol, float atol,
                                     float safety_factor, float beta, float fac_min, float fac_max, float h_max,
                                     float h, uint32_t n_max, uint32_t n_stiff, int out_type,
                                     ode_b200_config_f32* out);
int ode_oracle_config_rk4_f32(float x, float x_end, float step_size, ode_b200_config_f32* out);
int ode_oracle_integrate_batch_f32(const char* system, const ode_b200_config_f32* cfg, int64_t n_traj,
                                   const float* y0, const float* params, int params_per_traj,
                                   const ode_b200_outputs_f32* out, int n_threads);
float ode_oracle_powf(float x, float y);

/* Rk4 with T = f32 (src/rk4.rs generic over FloatNumber, src/dop_shared.rs:74), for the systems
 * whose RHS has an f32 form. Same arguments as ode_b200_rk4_f32_integrate_batch. */
int ode_oracle_rk4_f32_integrate_batch(const char* system, float x, float x_end, float step_size, int64_t n_traj,
                                       const float* y0, const float* params, int params_per_traj,
                                       const ode_b200_outputs_f32* out, int n_threads);

/* ContinuousOutputModel::evaluate (src/continuous_output_model.rs:31-56,86-104) for one model.
 * Returns 1 and fills y[dim] if Some, 0 if None. */
int ode_oracle_com_evaluate(int dim, int m, uint32_t n_intervals, double lower_bound,
                            const double* breakpoints, const double* step_sizes,
                            const double* coef /* [n_intervals][m][dim] */, double x, double* y);

/* tableau accessors, 1-based exactly like butcher_tableau.rs:57-84,380-426 */
double ode_oracle_tableau(const char* which, int i, int j);

/* number of threads the batch driver would use with n_threads <= 0 */
int ode_oracle_max_threads(void);

/* libm functions as seen by the oracle (what the crate's powf/exp call) */
double ode_oracle_pow(double x, double y);
double ode_oracle_exp(double x);

#ifdef __cplusplus
}
#endif
#endif
