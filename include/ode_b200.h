/*
 * ode_b200.h -- C ABI of the B200-native ensemble ODE integrator.
 *
 * This is the drop-in boundary for ONE path of the reference crate `ode_solvers`
 * (srenevey/ode-solvers v0.6.1): the explicit Runge-Kutta step loop of Rk4, Dopri5
 * and Dop853, run over a batch of independent trajectories on sm_100a.
 * Plain C, plain pointers and sizes, no torch / CUDA types in any signature
 * (a CUDA stream is passed as an opaque `void*`).
 *
 * What each entry point replaces in the reference (paths relative to the crate root):
 *
 *   ode_b200_config_new         Dopri5::new   src/dopri5.rs:76-105  (controller defaults :14-27)
 *                               Dop853::new   src/dop853.rs:76-109  (controller defaults :14-27)
 *   ode_b200_config_from_param  Dopri5::from_param src/dopri5.rs:129-184, Dop853::from_param src/dop853.rs:133-192
 *   ode_b200_config_rk4         Rk4::new      src/rk4.rs:42-68   (note the different argument order)
 *   ode_b200_system_*           trait System  src/dop_shared.rs:31-42 (`system` = RHS, `solout` = stop predicate);
 *                               the user's struct fields become the parameter row p[n_params]
 *   ode_b200_integrate_batch*   Dopri5::integrate src/dopri5.rs:255-260 -> integrate_core :263-445
 *                               Dopri5::integrate_with_continuous_output_model src/dopri5.rs:246-252
 *                               Dop853::integrate src/dop853.rs:254-512
 *                               Rk4::integrate    src/rk4.rs:71-100
 *   ode_b200_outputs            Stats src/dop_shared.rs:131-136; IntegrationError :120-128;
 *                               x_out()/y_out()/results() src/dopri5.rs:498-510, src/dop853.rs:562-574, src/rk4.rs:146-158;
 *                               ContinuousOutputModel fields src/continuous_output_model.rs:8-12,107-111
 *   ode_b200_com_evaluate       ContinuousOutputModel::evaluate src/continuous_output_model.rs:31-56
 *
 * The reference integrates ONE trajectory per stepper object; here a call integrates
 * n_traj trajectories that share one config and differ in y0 and parameters.
 * N = 1 reproduces the crate's single-stepper behaviour.
 *
 * There is NO CPU fallback: every integrate entry point fails (non-zero return,
 * message in ode_b200_last_error()) if no CUDA device is usable.
 */
#ifndef ODE_B200_H
#define ODE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODE_B200_ABI_VERSION 2
#define ODE_B200_MAX_DIM 32
#define ODE_B200_MAX_PARAMS 16

/* method ids */
enum { ODE_B200_RK4 = 0, ODE_B200_DOPRI5 = 1, ODE_B200_DOP853 = 2 };

/* OutputType, src/dop_shared.rs:112-117 (same order as the Rust enum) */
enum { ODE_B200_OUT_DENSE = 0, ODE_B200_OUT_SPARSE = 1, ODE_B200_OUT_CONTINUOUS = 2,
       /* EXTENSION (not in the reference, which builds a ContinuousOutputModel only for Dopri5 and
        * treats Continuous like Sparse in Dop853, dop853.rs:398,540-542): Dop853 computes its 8
        * dense-output vectors rcont[0..7] exactly as in Dense mode (dop853.rs:398-480) and records
        * them per accepted step as ContinuousOutputModel intervals (m = 8), pushing (x, y) like
        * Dopri5's Continuous mode. ContinuousOutputModel::evaluate on such an interval equals
        * Dop853::compute_y_out (dop853.rs:546-559). Dop853 only. */
       ODE_B200_OUT_CONTINUOUS8 = 3 };

/* per-trajectory status: 0 = Ok(Stats); 1..3 = IntegrationError variants (src/dop_shared.rs:120-128) */
enum {
    ODE_B200_OK = 0,
    ODE_B200_MAX_NUM_STEP_REACHED = 1, /* payload: x_final[i], n_step[i] */
    ODE_B200_STEP_SIZE_UNDERFLOW = 2,  /* payload: x_final[i] */
    ODE_B200_STIFFNESS_DETECTED = 3,   /* payload: x_final[i] */
    ODE_B200_OUTPUT_CAPACITY_EXCEEDED = 4, /* integration finished Ok but out_count > out_cap or com_count > com_cap */
    ODE_B200_UNSUPPORTED_DOMAIN = 5    /* input on which the reference panics or loops forever */
};

/* return codes of the API functions themselves */
enum {
    ODE_B200_SUCCESS = 0,
    ODE_B200_ERR_INVALID_ARGUMENT = -1,
    ODE_B200_ERR_NO_DEVICE = -2,
    ODE_B200_ERR_CUDA = -3,
    ODE_B200_ERR_UNKNOWN_SYSTEM = -4,
    ODE_B200_ERR_COMPILE = -5
};

/*
 * Batch-uniform solver configuration: the argument lists of Dopri5::from_param /
 * Dop853::from_param (16 args) and Rk4::new, as one POD. Fill it with the helpers
 * below, which reproduce the constructors' defaults and derived values bit for bit.
 */
typedef struct ode_b200_config {
    int32_t method;        /* ODE_B200_RK4 | DOPRI5 | DOP853 */
    int32_t out_type;      /* ODE_B200_OUT_*; `new` defaults to Dense (dopri5.rs:95, dop853.rs:96) */
    double x0;             /* `x` */
    double x_end;
    double dx;             /* dense-output increment (ignored unless Dense) */
    double rtol, atol;
    double safety_factor;  /* 0.9 */
    double alpha;          /* PI exponent: Dopri5 0.2-beta*0.75 (dopri5.rs:16,147); Dop853 1/8-beta*0.2 (dop853.rs:16,151) */
    double beta;           /* Dopri5 0.04, Dop853 0.0 */
    double fac_min;        /* Dopri5 0.2, Dop853 0.333 */
    double fac_max;        /* Dopri5 10,  Dop853 6 */
    double h_max;          /* x_end - x (Controller::new stores its abs, controller.rs:44) */
    double h0;             /* initial step; 0 => hinit() (dopri5.rs:283-286, dop853.rs:266-269) */
    double uround;         /* f64::EPSILON (dopri5.rs:89,160; dop853.rs:90,165) */
    double step_size;      /* Rk4 only (rk4.rs:42) */
    uint32_t n_max;        /* 100000 */
    uint32_t n_stiff;      /* 1000 */
} ode_b200_config;

/*
 * Results of a batch. Every pointer may be NULL (that output is skipped).
 * For ode_b200_integrate_batch the pointers are HOST memory, for
 * ode_b200_integrate_batch_device they are DEVICE memory on the context's GPU.
 *
 * Layouts (n = n_traj, d = system dimension):
 *   structure-of-arrays for per-trajectory scalars and the final state:
 *     y_final[c*n + i]  component c of trajectory i    (same layout as the y0 input)
 *   trajectory-major for the per-step results, so that one trajectory's
 *   x_out()/y_out() are contiguous like the crate's Vec<T> / Vec<OVector<T,D>>:
 *     out_x[i*out_cap + s], out_y[(i*out_cap + s)*d + c]
 *     com_break[i*com_cap + s], com_step[i*com_cap + s], com_coef[((i*com_cap + s)*m + r)*d + c]
 *   with m = 5 coefficient vectors (rcont[0..4]) for Dopri5 (dopri5.rs:52,485), m = 8 for the
 *   ODE_B200_OUT_CONTINUOUS8 extension of Dop853.
 */
typedef struct ode_b200_outputs {
    double*   y_final;    /* [d][n]  state when the loop exits (last accepted y)                      */
    double*   x_final;    /* [n]     x when the loop exits (also the `x` of an IntegrationError)      */
    double*   h_next;     /* [n]     the stepper's h_old on exit: posneg*h_new on Ok (dopri5.rs:433)  */
    uint32_t* num_eval;   /* [n]     Stats.num_eval                                                   */
    uint32_t* accepted;   /* [n]     Stats.accepted_steps                                             */
    uint32_t* rejected;   /* [n]     Stats.rejected_steps                                             */
    uint32_t* n_step;     /* [n]     attempted steps (MaxNumStepReached payload)                      */
    int32_t*  status;     /* [n]     ODE_B200_OK ...                                                  */

    /* results(): every (x, y) the reference pushes (dense samples, or one per accepted step) */
    int64_t   out_cap;    /* capacity in samples per trajectory (0 = record nothing)                  */
    double*   out_x;      /* [n][out_cap]                                                             */
    double*   out_y;      /* [n][out_cap][d]                                                          */
    uint32_t* out_count;  /* [n] number of pushes the reference makes (may exceed out_cap)            */

    /* ContinuousOutputModel (OutputType::Continuous only) */
    int64_t   com_cap;    /* capacity in intervals per trajectory                                     */
    double*   com_lower;  /* [n] lower_bound                                                          */
    double*   com_break;  /* [n][com_cap] breakpoints                                                 */
    double*   com_step;   /* [n][com_cap] Interval.step_size                                          */
    double*   com_coef;   /* [n][com_cap][m][d] Interval.coefficients                                 */
    uint32_t* com_count;  /* [n] number of intervals (may exceed com_cap)                             */
} ode_b200_outputs;

typedef struct ode_b200_ctx ode_b200_ctx; /* one per (host thread, GPU); owns stream + workspace */

/* ---- library / device ---------------------------------------------------------- */
int         ode_b200_abi_version(void);
int         ode_b200_device_count(void);          /* CUDA devices visible; 0 if none */
const char* ode_b200_last_error(void);            /* thread-local, never NULL */
/* measured FP64 pipe issue rate of `device` in thread-instructions/s (DFMA chain microbenchmark) */
int         ode_b200_measure_fp64_peak(int device, double* thread_inst_per_s, double* sm_clock_mhz_est);

/* ---- systems (trait System) ----------------------------------------------------- */
/* built-ins: "lorenz","kepler","cr3bp","robertson","bouncing_ball","threebody18",
 * "test_rk4_1","test_rk4_2","test_exp","test_exp_stop","test_cubic","test_cubic_stop" */
int ode_b200_system_builtin(const char* name);    /* -> sys_id >= 0, or ODE_B200_ERR_UNKNOWN_SYSTEM */
int ode_b200_system_dim(int sys_id);
int ode_b200_system_n_params(int sys_id);
const char* ode_b200_system_name(int sys_id);
/*
 * Register a user system from CUDA source (compiled with NVRTC for sm_100a, -fmad=false).
 * rhs_body is the body of
 *   __device__ void rhs(double x, const double* y, double* dy, const double* p)
 * and solout_body (NULL = never stop) the body of
 *   __device__ bool solout(double x, const double* y, const double* dy, const double* p)
 * rhs must write every component of dy. Inside rhs_body, `pow_d(x, y)` and `exp_d(x)` are the
 * host-libm-identical f64::powf / f64::exp; `div_d(a, b)` and `sqrt_d(x)` are `a / b` and `sqrt(x)`
 * (same correctly rounded results) routed through the step loop's branch-free speculative arithmetic,
 * which lets the kernel evaluate a whole attempted step as one straight-line block -- recommended for
 * right-hand sides that divide or take roots (the built-in Kepler / CR3BP / 3-body systems use it);
 * the plain operators stay valid.
 * Returns sys_id >= 0 or a negative error (compile log in ode_b200_last_error()).
 */
int ode_b200_system_from_source(const char* name, int dim, int n_params,
                                const char* rhs_body, const char* solout_body);
/* Compile the kernel of (sys_id, method, out_type) now instead of at the first launch; needs
 * NVRTC but no GPU, so a source error surfaces early. No-op for built-ins. */
int ode_b200_system_compile(int sys_id, int method, int out_type);

/* ---- configuration (constructors) ----------------------------------------------- */
/* Dopri5::new / Dop853::new defaults. method = ODE_B200_DOPRI5 | ODE_B200_DOP853. */
int ode_b200_config_new(int method, double x, double x_end, double dx, double rtol, double atol,
                        ode_b200_config* out);
/* Dopri5::from_param / Dop853::from_param. */
int ode_b200_config_from_param(int method, double x, double x_end, double dx, double rtol, double atol,
                               double safety_factor, double beta, double fac_min, double fac_max,
                               double h_max, double h, uint32_t n_max, uint32_t n_stiff, int out_type,
                               ode_b200_config* out);
/* Rk4::new(f, x, y, x_end, step_size). */
int ode_b200_config_rk4(double x, double x_end, double step_size, ode_b200_config* out);

/* ---- contexts ------------------------------------------------------------------- */
ode_b200_ctx* ode_b200_create(int device);        /* NULL on failure (see last_error) */
void          ode_b200_destroy(ode_b200_ctx*);
/* Device time (ms, CUDA events) of the most recent step-loop kernel this context launched ON ITS OWN
 * STREAM: every host-buffer entry, and ode_b200_integrate_batch_device when cuda_stream ==
 * ode_b200_stream(ctx). A device-entry launch on a caller's stream is NOT timed (the caller owns that
 * stream and times it itself); the value then still describes the last launch on the context stream
 * (0.0 if there was none). Blocks until that kernel has finished. */
double        ode_b200_last_kernel_ms(const ode_b200_ctx*);
/* number of kernel launches issued by this context since creation */
uint64_t      ode_b200_launch_count(const ode_b200_ctx*);
/* the context's own stream (cudaStream_t as void*) */
void*         ode_b200_stream(const ode_b200_ctx*);

/* ---- the hot path ---------------------------------------------------------------- */
/*
 * Host-buffer entry (what a reference-side FFI binding calls): copies y0/params to the
 * GPU, runs the persistent step-loop kernel, copies the requested outputs back,
 * synchronises. y0 is SoA [d][n]; params is [n_params] if params_per_traj == 0
 * (shared by all trajectories) else SoA [n_params][n].
 */
int ode_b200_integrate_batch(ode_b200_ctx* ctx, int sys_id, const ode_b200_config* cfg,
                             int64_t n_traj, const double* y0, const double* params,
                             int params_per_traj, const ode_b200_outputs* out);
/*
 * Device-buffer entry: same, but every pointer (y0, params, all of *out) is device memory
 * on the context's GPU and the kernel is enqueued on `cuda_stream` (a cudaStream_t passed as
 * void*; NULL = the CUDA default stream, as everywhere in CUDA) without synchronising.
 * ode_b200_stream(ctx) returns the context's own non-blocking stream.
 * Launches of one context may be in flight on different streams at the same time (each launch owns
 * one of 64 work-queue heads of the context, used round-robin): at most 64 un-finished launches per
 * context. The context itself is not thread-safe: one host thread per context.
 */
int ode_b200_integrate_batch_device(ode_b200_ctx* ctx, int sys_id, const ode_b200_config* cfg,
                                    int64_t n_traj, const double* y0, const double* params,
                                    int params_per_traj, const ode_b200_outputs* out,
                                    void* cuda_stream);
/*
 * Multi-GPU host entry: shards [0, n_traj) into n_gpus contiguous index ranges, one host
 * thread + context + stream per GPU, no collective; outputs are gathered into the caller's
 * host arrays. n_gpus <= ode_b200_device_count().
 */
int ode_b200_integrate_batch_multi(int sys_id, const ode_b200_config* cfg, int64_t n_traj,
                                   const double* y0, const double* params, int params_per_traj,
                                   int n_gpus, const ode_b200_outputs* out);

/*
 * ---- T = f32 ---------------------------------------------------------------------
 * The crate is generic over FloatNumber = {f32, f64} (src/dop_shared.rs:52-77); examples/bouncing_ball.rs
 * runs Rk4 in f32 (:15,36) and names Dopri5 in f32 as the alternative (:35). These entry points are the
 * three steppers with T = f32: every operation in single precision, every literal and tableau
 * coefficient the f64 value rounded to f32 (T::from / to_subset), powf the host libm's. Available for the
 * built-in systems whose right-hand side needs no libm call (lorenz, robertson, bouncing_ball,
 * test_rk4_1, test_rk4_2). Host pointers; layouts as for f64. A coverage path, not the hot one: the
 * kernels are generic loops over the state (one trajectory per thread), not the tuned f64 kernels.
 */
typedef struct ode_b200_config_f32 { /* ode_b200_config with T = f32, field for field */
    int32_t method;
    int32_t out_type;
    float x0, x_end, dx;
    float rtol, atol;
    float safety_factor, alpha, beta, fac_min, fac_max, h_max, h0;
    float uround;    /* Dopri5::new: f32::EPSILON (dopri5.rs:89); from_param and Dop853: f64::EPSILON as f32 (Q3) */
    float step_size;
    uint32_t n_max, n_stiff;
} ode_b200_config_f32;

typedef struct ode_b200_outputs_f32 { /* ode_b200_outputs with T = f32, field for field */
    float*    y_final;
    float*    x_final;
    float*    h_next;
    uint32_t* num_eval;
    uint32_t* accepted;
    uint32_t* rejected;
    uint32_t* n_step;
    int32_t*  status;
    int64_t   out_cap;
    float*    out_x;
    float*    out_y;
    uint32_t* out_count;
    int64_t   com_cap;
    float*    com_lower;
    float*    com_break;
    float*    com_step;
    float*    com_coef;
    uint32_t* com_count;
} ode_b200_outputs_f32;

/* Dopri5::<f32>::new / Dop853::<f32>::new, ::from_param, Rk4::<f32>::new */
int ode_b200_config_new_f32(int method, float x, float x_end, float dx, float rtol, float atol,
                            ode_b200_config_f32* out);
int ode_b200_config_from_param_f32(int method, float x, float x_end, float dx, float rtol, float atol,
                                   float safety_factor, float beta, float fac_min, float fac_max,
                                   float h_max, float h, uint32_t n_max, uint32_t n_stiff, int out_type,
                                   ode_b200_config_f32* out);
int ode_b200_config_rk4_f32(float x, float x_end, float step_size, ode_b200_config_f32* out);
/* integrate() / integrate_with_continuous_output_model() with T = f32 for a batch (all three methods;
 * ODE_B200_OUT_CONTINUOUS8 is not available in f32) */
int ode_b200_integrate_batch_f32(ode_b200_ctx* ctx, int sys_id, const ode_b200_config_f32* cfg,
                                 int64_t n_traj, const float* y0, const float* params,
                                 int params_per_traj, const ode_b200_outputs_f32* out);
/* Rk4::<f32>::new(f, x, y, x_end, step_size) + integrate(): ode_b200_config_rk4_f32 + the call above */
int ode_b200_rk4_f32_integrate_batch(ode_b200_ctx* ctx, int sys_id, float x, float x_end, float step_size,
                                     int64_t n_traj, const float* y0, const float* params,
                                     int params_per_traj, const ode_b200_outputs_f32* out);

/* kernel scheduling knob: 1 (default) = persistent kernel with lane refill from a work
 * queue; 0 = static one-thread-per-trajectory launch (kept for the ablation in DESIGN.md) */
int ode_b200_set_work_queue(ode_b200_ctx* ctx, int enabled);
/* host-buffer entry: pipeline of upload, integration and copy-back. chunk_log2 = 16 (default): the batch is
 * cut into chunks of 2^16 trajectories; the kernel starts before its inputs have arrived (lanes wait for
 * their chunk), and the per-trajectory results of a finished chunk are copied back while the rest of the
 * batch is still being integrated. Used for built-in systems when the call asks for per-trajectory results
 * only (out_cap = com_cap = 0), the batch has at least two chunks and the caller's result arrays are
 * pinned host memory (cudaHostAlloc / cudaHostRegister); 0 switches it off
 * (upload, kernel, copy-back one after the other, or zero-copy result stores: ode_b200_set_zero_copy). */
int ode_b200_set_pipeline(ode_b200_ctx* ctx, int chunk_log2);
/* The environment variable ODE_B200_PIPELINE=<chunk_log2> sets the initial value for every context of the process
 * (incl. the ones ode_b200_integrate_batch_multi creates): ODE_B200_PIPELINE=0 for runs under a profiler that
 * serialises kernels (ncu), which cannot execute a kernel that waits for concurrent host copies. */
/* large states: 1 (default) = Dopri5 and Dop853 run one trajectory per group of three lanes for the systems that
 * have that form (threebody18), the state and stage vectors dealt out component-wise so that they stay in
 * registers; 0 = one trajectory per thread like everything else (kept for the ablation; same bits) */
int ode_b200_set_lane_split(ode_b200_ctx* ctx, int enabled);
/* host-buffer entry: 1 (default) = per-trajectory results are stored by the kernel straight into the
 * caller's arrays when those are pinned, device-mapped host memory (no copy-back phase; best for one
 * GPU per host); 0 = always through device buffers and one DMA copy per array (better when several
 * GPUs share the host's PCIe) */
int ode_b200_set_zero_copy(ode_b200_ctx* ctx, int enabled);

/* ---- ContinuousOutputModel::evaluate, batched on the device ---------------------- */
/*
 * For each trajectory i and query q: y[(i*n_query + q)*d + c] and found[i*n_query + q]
 * (0 = the reference returns None). Host pointers. xq is [n_query] shared by all
 * trajectories. m = coefficient vectors per interval (5 for Dopri5).
 */
int ode_b200_com_evaluate(ode_b200_ctx* ctx, int64_t n_traj, int dim, int m, int64_t com_cap,
                          const double* com_lower, const double* com_break, const double* com_step,
                          const double* com_coef, const uint32_t* com_count,
                          int64_t n_query, const double* xq, double* y, uint8_t* found);
/* The same with DEVICE pointers throughout (the model where ode_b200_integrate_batch_device left it, xq, y and
 * found on the context's GPU), enqueued on `cuda_stream` without synchronising: resampling an ensemble on a
 * common grid without the model ever leaving HBM. Timed by ode_b200_last_kernel_ms when cuda_stream is the
 * context's own stream. Ascending queries are several times faster than unordered ones (neighbouring lanes
 * then read neighbouring intervals); the host-pointer entry above sorts unordered queries itself. */
int ode_b200_com_evaluate_device(ode_b200_ctx* ctx, int64_t n_traj, int dim, int m, int64_t com_cap,
                                 const double* com_lower, const double* com_break, const double* com_step,
                                 const double* com_coef, const uint32_t* com_count,
                                 int64_t n_query, const double* xq, double* y, uint8_t* found, void* cuda_stream);

/* ---- introspection used by the parity tests -------------------------------------- */
/* the device library's own copy of a tableau coefficient (0-based); which: "a5","c5","d5","e5",
 * "a8","b8","bhh8","c8","d8","e8" */
double ode_b200_tableau(const char* which, int i, int j);
/* evaluate the device pow()/exp() clones on the GPU for n inputs (host pointers) */
int ode_b200_debug_pow(ode_b200_ctx* ctx, int64_t n, const double* x, const double* y, double* out);
int ode_b200_debug_exp(ode_b200_ctx* ctx, int64_t n, const double* x, double* out);
/* the f32 powf clone (ode_powf.h: groundwork for the f32 adaptive steppers) evaluated on the GPU */
int ode_b200_debug_powf(ode_b200_ctx* ctx, int64_t n, const float* x, const float* y, float* out);
/* the kernels' exact division / square root building blocks (must equal the IEEE result bit for
 * bit). mode 0: a/b; 1: a/b and -a/b through one shared reciprocal; 2..6: the step loop's
 * speculative arithmetic (2: div, 3: div_nz, 4: sqrt(a), 5: sqrt_nz(a), 6: a/b and -a/b through one
 * reciprocal), which returns the NaN 0x7ff8f1a900000000 when it flags the operands for the exact
 * path instead of producing a value */
int ode_b200_debug_div(ode_b200_ctx* ctx, int64_t n, const double* a, const double* b, double* out,
                       int mode);

#ifdef __cplusplus
}
#endif
#endif /* ODE_B200_H */
