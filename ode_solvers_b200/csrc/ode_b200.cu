/*
 * ode_b200.cu -- implementation of the C ABI (include/ode_b200.h): system registry,
 * constructors, per-GPU contexts, the launchers of the persistent step-loop kernels,
 * the multi-GPU sharding driver and small device utilities (FP64 peak probe, batched
 * ContinuousOutputModel::evaluate, pow/exp clone probes).
 *
 * No CPU fallback exists anywhere in this file: without a usable CUDA device every
 * integrate entry point returns an error.
 */
#include <cuda.h> /* types of the tensor-map encoder only: the entry point comes from cudaGetDriverEntryPoint */
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ode_b200.h"
#include "ode_device.cuh"
#include "ode_powf.h"
#include "ode_jit.h"
#include "ode_registry.h"

using namespace ode_b200;

/* ------------------------------------------------------------------ errors */
static thread_local std::string g_last_error = "";

static int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(ODE_B200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));    \
    } while (0)

extern "C" const char* ode_b200_last_error(void) { return g_last_error.c_str(); }
extern "C" int ode_b200_abi_version(void) { return ODE_B200_ABI_VERSION; }
extern "C" int ode_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

/* ------------------------------------------------------------------ registry */
extern "C" {
extern const SystemEntry ode_b200_entry_lorenz, ode_b200_entry_kepler, ode_b200_entry_cr3bp,
    ode_b200_entry_robertson, ode_b200_entry_bouncing_ball, ode_b200_entry_threebody18,
    ode_b200_entry_test_rk4_1, ode_b200_entry_test_rk4_2, ode_b200_entry_test_exp,
    ode_b200_entry_test_exp_stop, ode_b200_entry_test_cubic, ode_b200_entry_test_cubic_stop;
}
static const SystemEntry* const kBuiltins[] = {
    &ode_b200_entry_lorenz,        &ode_b200_entry_kepler,      &ode_b200_entry_cr3bp,
    &ode_b200_entry_robertson,     &ode_b200_entry_bouncing_ball, &ode_b200_entry_threebody18,
    &ode_b200_entry_test_rk4_1,    &ode_b200_entry_test_rk4_2,  &ode_b200_entry_test_exp,
    &ode_b200_entry_test_exp_stop, &ode_b200_entry_test_cubic,  &ode_b200_entry_test_cubic_stop,
};
static constexpr int kNumBuiltins = sizeof(kBuiltins) / sizeof(kBuiltins[0]);

extern "C" int ode_b200_system_builtin(const char* name) {
    if (!name) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "name is NULL");
    for (int i = 0; i < kNumBuiltins; ++i)
        if (!strcmp(kBuiltins[i]->name, name)) return i;
    return fail(ODE_B200_ERR_UNKNOWN_SYSTEM, std::string("unknown built-in system: ") + name);
}

struct SysInfo {
    const char* name;
    int dim, n_params;
    bool jit;
};
static bool sys_info(int id, SysInfo* s) {
    if (id >= 0 && id < kNumBuiltins) {
        *s = {kBuiltins[id]->name, kBuiltins[id]->dim, kBuiltins[id]->n_params, false};
        return true;
    }
    const JitSystem* j = jit_system(id - kNumBuiltins);
    if (!j) return false;
    *s = {j->name.c_str(), j->dim, j->n_params, true};
    return true;
}
extern "C" int ode_b200_system_dim(int id) {
    SysInfo s;
    return sys_info(id, &s) ? s.dim : fail(ODE_B200_ERR_UNKNOWN_SYSTEM, "bad sys_id");
}
extern "C" int ode_b200_system_n_params(int id) {
    SysInfo s;
    return sys_info(id, &s) ? s.n_params : fail(ODE_B200_ERR_UNKNOWN_SYSTEM, "bad sys_id");
}
extern "C" const char* ode_b200_system_name(int id) {
    SysInfo s;
    return sys_info(id, &s) ? s.name : nullptr;
}
extern "C" int ode_b200_system_from_source(const char* name, int dim, int n_params, const char* rhs_body,
                                           const char* solout_body) {
    if (!name || !rhs_body || dim < 1 || dim > ODE_B200_MAX_DIM || n_params < 0 || n_params > ODE_B200_MAX_PARAMS)
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "system_from_source: bad arguments");
    std::string err;
    int id = jit_register(name, dim, n_params, rhs_body, solout_body, &err);
    if (id < 0) return fail(ODE_B200_ERR_COMPILE, err);
    return kNumBuiltins + id;
}

extern "C" int ode_b200_system_compile(int sys_id, int method, int out_type) {
    SysInfo s;
    if (!sys_info(sys_id, &s)) return fail(ODE_B200_ERR_UNKNOWN_SYSTEM, "bad sys_id");
    if (method < ODE_B200_RK4 || method > ODE_B200_DOP853 || out_type < ODE_B200_OUT_DENSE ||
        out_type > ODE_B200_OUT_CONTINUOUS8 || (out_type == ODE_B200_OUT_CONTINUOUS8 && method == ODE_B200_DOPRI5))
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "system_compile: bad method / out_type");
    if (!s.jit) return ODE_B200_SUCCESS; /* built-ins are compiled ahead of time */
    std::string err;
    if (!jit_compile(sys_id - kNumBuiltins, method, out_type, false, &err)) return fail(ODE_B200_ERR_COMPILE, err);
    return ODE_B200_SUCCESS;
}

/* ------------------------------------------------------------------ constructors */
static const double kEps = 2.220446049250313e-16; /* f64::EPSILON */

/* Dopri5::new (dopri5.rs:76-105, controller defaults :14-27); Dop853::new (dop853.rs:76-109, :14-27) */
extern "C" int ode_b200_config_new(int method, double x, double x_end, double dx, double rtol, double atol,
                                   ode_b200_config* c) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "config is NULL");
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
    c->uround = kEps;
    c->n_max = 100000;
    c->n_stiff = 1000;
    if (method == ODE_B200_DOPRI5) {
        volatile double b = 0.04; /* keep the two roundings of `0.2 - 0.04 * 0.75` */
        c->beta = 0.04;
        c->alpha = 0.2 - b * 0.75;
        c->fac_max = 10.0;
        c->fac_min = 0.2;
    } else if (method == ODE_B200_DOP853) {
        c->beta = 0.0;
        c->alpha = 1.0 / 8.0;
        c->fac_max = 6.0;
        c->fac_min = 0.333;
    } else {
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "config_new: method must be DOPRI5 or DOP853");
    }
    return ODE_B200_SUCCESS;
}

/* Dopri5::from_param (dopri5.rs:129-184); Dop853::from_param (dop853.rs:133-192) */
extern "C" int ode_b200_config_from_param(int method, double x, double x_end, double dx, double rtol, double atol,
                                          double safety_factor, double beta, double fac_min, double fac_max,
                                          double h_max, double h, uint32_t n_max, uint32_t n_stiff, int out_type,
                                          ode_b200_config* c) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "config is NULL");
    if (out_type < ODE_B200_OUT_DENSE || out_type > ODE_B200_OUT_CONTINUOUS8)
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "config_from_param: bad out_type");
    memset(c, 0, sizeof(*c));
    c->method = method;
    c->out_type = out_type;
    c->x0 = x;
    c->x_end = x_end;
    c->dx = dx;
    c->rtol = rtol;
    c->atol = atol;
    c->safety_factor = safety_factor;
    c->beta = beta;
    c->fac_min = fac_min;
    c->fac_max = fac_max;
    c->h_max = h_max;
    c->h0 = h;
    c->uround = kEps;
    c->n_max = n_max;
    c->n_stiff = n_stiff;
    volatile double b = beta;
    if (method == ODE_B200_DOPRI5) {
        volatile double t = b * 0.75;
        c->alpha = 0.2 - t;
    } else if (method == ODE_B200_DOP853) {
        volatile double t = b * 0.2;
        c->alpha = 1.0 / 8.0 - t;
    } else {
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "config_from_param: method must be DOPRI5 or DOP853");
    }
    return ODE_B200_SUCCESS;
}

/* Rk4::new (rk4.rs:42-68) */
extern "C" int ode_b200_config_rk4(double x, double x_end, double step_size, ode_b200_config* c) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "config is NULL");
    memset(c, 0, sizeof(*c));
    c->method = ODE_B200_RK4;
    c->out_type = ODE_B200_OUT_SPARSE;
    c->x0 = x;
    c->x_end = x_end;
    c->step_size = step_size;
    return ODE_B200_SUCCESS;
}

/* ------------------------------------------------------------------ contexts */
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

/* work-queue heads per context: launch k uses word k % kQueueSlots (more than kQueueSlots launches of
 * one context in flight at once is not supported; include/ode_b200.h says so) */
static constexpr unsigned kQueueSlots = 64;
static constexpr int kPipeMaxChunks = 4096;

struct ode_b200_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    int use_queue = 1;
    int zero_copy = 1; /* per-trajectory results straight into pinned host arrays (ode_b200_set_zero_copy) */
    int lane_split = 1; /* large states: one trajectory per group of three lanes where the system has that form */
    int tensor_stores = 1; /* Rk4 STAGE kernels: one TMA tensor store per warp and group of samples (ODE_B200_TENSOR_STORES=0: off) */
    /* host pipeline of the host-buffer entry (ode_kernel_args.h): log2 of the chunk size, 0 = off */
    int pipeline_shift = 16;
    cudaStream_t up_stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_pipe = nullptr;
    unsigned* pipe_flags_host = nullptr; /* mapped pinned: one word per chunk + one constant 1 at [kPipeMaxChunks] */
    unsigned* pipe_flags_dev = nullptr;  /* its device alias */
    DevBuf pipe_dev;                     /* upload_ready[chunks], chunk_done[chunks] */
    unsigned long long* queue = nullptr;
    uint64_t launches = 0;
    double last_ms = 0.0;
    /* grow-only workspace for the host-buffer entry points */
    DevBuf y0, params, y_final, x_final, h_next, num_eval, accepted, rejected, n_step, status, out_x, out_y,
        out_count, com_lower, com_break, com_step, com_coef, com_count, scratch0, scratch1, scratch2;
};

static int ensure(ode_b200_ctx* c, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) CUDA_TRY(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    CUDA_TRY(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    (void)c;
    return 0;
}

extern "C" ode_b200_ctx* ode_b200_create(int device) {
    int n = ode_b200_device_count();
    if (n <= 0) {
        fail(ODE_B200_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
        return nullptr;
    }
    if (device < 0 || device >= n) {
        fail(ODE_B200_ERR_INVALID_ARGUMENT, "device index out of range");
        return nullptr;
    }
    ode_b200_ctx* c = new ode_b200_ctx();
    c->device = device;
    /* ODE_B200_PIPELINE=<chunk_log2>: initial value of ode_b200_set_pipeline for every context of the process, also
     * the ones ode_b200_integrate_batch_multi makes (0 = off: a profiler that serialises kernels cannot run one
     * that waits for concurrent host copies) */
    if (const char* e = getenv("ODE_B200_TENSOR_STORES")) c->tensor_stores = atoi(e) != 0; /* ablation switch */
    if (const char* e = getenv("ODE_B200_PIPELINE")) {
        const int v = atoi(e);
        if (v == 0 || (v >= 8 && v <= 24)) c->pipeline_shift = v;
    }
    bool ok = cudaSetDevice(device) == cudaSuccess &&
              cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&c->ev0) == cudaSuccess && cudaEventCreate(&c->ev1) == cudaSuccess &&
              cudaMalloc((void**)&c->queue, kQueueSlots * sizeof(unsigned long long)) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->up_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_pipe, cudaEventDisableTiming) == cudaSuccess &&
              cudaHostAlloc((void**)&c->pipe_flags_host, (kPipeMaxChunks + 1) * sizeof(unsigned), cudaHostAllocMapped) ==
                  cudaSuccess &&
              cudaHostGetDevicePointer((void**)&c->pipe_flags_dev, c->pipe_flags_host, 0) == cudaSuccess;
    if (ok) c->pipe_flags_host[kPipeMaxChunks] = 1u;
    if (!ok) {
        fail(ODE_B200_ERR_CUDA, std::string("context creation failed: ") + cudaGetErrorString(cudaGetLastError()));
        delete c;
        return nullptr;
    }
    return c;
}

extern "C" void ode_b200_destroy(ode_b200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DevBuf* bufs[] = {&c->y0, &c->params, &c->y_final, &c->x_final, &c->h_next, &c->num_eval, &c->accepted,
                      &c->rejected, &c->n_step, &c->status, &c->out_x, &c->out_y, &c->out_count, &c->com_lower,
                      &c->com_break, &c->com_step, &c->com_coef, &c->com_count, &c->scratch0, &c->scratch1,
                      &c->scratch2};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    if (c->queue) cudaFree(c->queue);
    if (c->pipe_dev.p) cudaFree(c->pipe_dev.p);
    if (c->pipe_flags_host) cudaFreeHost(c->pipe_flags_host);
    if (c->up_stream) cudaStreamDestroy(c->up_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->ev_pipe) cudaEventDestroy(c->ev_pipe);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" double ode_b200_last_kernel_ms(const ode_b200_ctx* c) {
    if (!c || !c->timed) return 0.0;
    float ms = 0.f;
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return 0.0;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return 0.0;
    return (double)ms;
}
extern "C" uint64_t ode_b200_launch_count(const ode_b200_ctx* c) { return c ? c->launches : 0; }
extern "C" void* ode_b200_stream(const ode_b200_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int ode_b200_set_zero_copy(ode_b200_ctx* c, int enabled) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
    c->zero_copy = enabled ? 1 : 0;
    return ODE_B200_SUCCESS;
}
extern "C" int ode_b200_set_pipeline(ode_b200_ctx* c, int chunk_log2) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (chunk_log2 != 0 && (chunk_log2 < 8 || chunk_log2 > 24))
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "set_pipeline: chunk_log2 must be 0 (off) or 8..24");
    c->pipeline_shift = chunk_log2;
    return ODE_B200_SUCCESS;
}
extern "C" int ode_b200_set_lane_split(ode_b200_ctx* c, int enabled) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
    c->lane_split = enabled ? 1 : 0;
    return ODE_B200_SUCCESS;
}
extern "C" int ode_b200_set_work_queue(ode_b200_ctx* c, int enabled) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
    c->use_queue = enabled ? 1 : 0;
    return ODE_B200_SUCCESS;
}

/* ------------------------------------------------------------------ launch */
static int validate(const ode_b200_config* cfg, const SysInfo& s, int64_t n, const double* y0,
                    const double* params, const ode_b200_outputs* out) {
    if (!cfg || !out) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "cfg / out is NULL");
    if (n < 0) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "n_traj < 0");
    if (n > 0 && !y0) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "y0 is NULL");
    if (n > 0 && s.n_params > 0 && !params) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "params is NULL");
    if (cfg->method < ODE_B200_RK4 || cfg->method > ODE_B200_DOP853)
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "bad method");
    if (cfg->out_type < ODE_B200_OUT_DENSE || cfg->out_type > ODE_B200_OUT_CONTINUOUS8)
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "bad out_type");
    if (cfg->out_type == ODE_B200_OUT_CONTINUOUS8 && cfg->method == ODE_B200_DOPRI5)
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "ODE_B200_OUT_CONTINUOUS8 is a Dop853-only extension");
    if (out->out_cap < 0 || out->com_cap < 0) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "negative capacity");
    if (out->out_cap > 0 && (!out->out_x || !out->out_y))
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "out_cap > 0 needs out_x and out_y");
    if (out->com_cap > 0 && (!out->com_break || !out->com_step || !out->com_coef))
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "com_cap > 0 needs com_break, com_step and com_coef");
    return 0;
}

__global__ void fill_unsupported_kernel(KernelArgs a, int dim) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n_traj) return;
    const ode_b200_outputs& o = a.out;
    if (o.y_final)
        for (int c = 0; c < dim; ++c) o.y_final[(size_t)c * a.n_traj + t] = a.y0[(size_t)c * a.n_traj + t];
    if (o.x_final) o.x_final[t] = a.cfg.x0;
    if (o.h_next) o.h_next[t] = 0.0;
    if (o.num_eval) o.num_eval[t] = 0;
    if (o.accepted) o.accepted[t] = 0;
    if (o.rejected) o.rejected[t] = 0;
    if (o.n_step) o.n_step[t] = 0;
    if (o.status) o.status[t] = ODE_B200_UNSUPPORTED_DOMAIN;
    if (o.out_count) o.out_count[t] = 0;
    if (o.com_count) o.com_count[t] = 0;
    if (o.com_lower) o.com_lower[t] = 0.0;
}

/* enqueue the step-loop kernel for device-resident inputs/outputs on `stream` */
struct PipeArgs { /* host pipeline of this launch (ode_kernel_args.h); shift == 0: off */
    int shift = 0;
    const unsigned* upload_ready = nullptr;
    unsigned* chunk_done = nullptr;
    unsigned* chunk_flag = nullptr;
};

/* Tensor maps of the Rk4 STAGE kernels' warp-level result stores (KernelArgs::map_y / map_x, push_warp_):
 * out_y as [n_traj][out_pitch / G][G * dim] doubles and out_x as [n_traj][out_pitch / G][G], box = 32 trajectories x
 * 1 group. Returns 1 when both are set. */
static int make_result_maps(KernelArgs* ka, int dim) {
    typedef CUresult (*encode_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_t encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        (void)cudaGetLastError();
        return (encode_t)fn;
    }();
    const int G = stage_group(dim, ODE_B200_RK4);
    if (!encode || ka->out_pitch % G != 0 || ka->n_traj > 0x7fffffffLL || ka->out_pitch / G > 0x7fffffffLL) return 0;
    static_assert(sizeof(CUtensorMap) == sizeof(TensorMapBlob), "CUtensorMap is 128 bytes");
    auto make = [&](double* base, int inner, TensorMapBlob* out) {
        CUtensorMap m;
        const cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)(ka->out_pitch / G), (cuuint64_t)ka->n_traj};
        const cuuint64_t strides[2] = {(cuuint64_t)inner * 8, (cuuint64_t)(ka->out_pitch / G) * inner * 8};
        const cuuint32_t box[3] = {(cuuint32_t)inner, 1, 32}, es[3] = {1, 1, 1};
        if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
        memcpy(out, &m, sizeof(m));
        return true;
    };
    return make(ka->out.out_y, G * dim, &ka->map_y) && make(ka->out.out_x, G, &ka->map_x) ? 1 : 0;
}

static int launch_device(ode_b200_ctx* c, int sys_id, const ode_b200_config* cfg, int64_t n, const double* y0,
                         const double* params, int per_traj, const ode_b200_outputs* out, long long out_pitch,
                         long long com_pitch, cudaStream_t stream, const PipeArgs* pipe = nullptr) {
    SysInfo s;
    if (!sys_info(sys_id, &s)) return fail(ODE_B200_ERR_UNKNOWN_SYSTEM, "bad sys_id");
    if (int rc = validate(cfg, s, n, y0, params, out)) return rc;
    if (n == 0) return ODE_B200_SUCCESS;
    CUDA_TRY(cudaSetDevice(c->device));

    KernelArgs ka;
    memset(&ka, 0, sizeof(ka));
    ka.cfg = *cfg;
    /* Controller::new, controller.rs:29-49 */
    ka.facc1 = 1.0 / cfg->fac_min;
    ka.facc2 = 1.0 / cfg->fac_max;
    ka.h_max_abs = fabs(cfg->h_max);
    ka.posneg = (cfg->x_end - cfg->x0) > 0.0 ? 1.0 : -1.0; /* sign(1, x_end - x) */
#ifndef __CUDA_ARCH__ /* host pass only: the host tables do not exist in the device pass */
    {
        /* fac_old.powf(-beta) at fac_old's floor 1e-4, through the SAME clone the kernels run
         * (host build of ode_pow.h; equal to libm pow bit for bit, tests/test_pow_clone.py) */
        const ode_pow_tabs T = ode_pow_host_tabs();
        ka.pw_floor = cfg->beta == 0.0 ? 1.0 : ode_pow(1.0E-4, -cfg->beta, T);
        ka.pow_shared_log = ode_pow_y_is_plain_(cfg->alpha) && (cfg->beta == 0.0 || ode_pow_y_is_plain_(-cfg->beta));
    }
#endif
    ka.n_traj = n;
    ka.y0 = y0;
    ka.params = params;
    ka.params_per_traj = per_traj ? 1 : 0;
    ka.use_queue = c->use_queue;
    /* every launch gets its own work-queue head out of a ring of kQueueSlots words, so launches of one
     * context that are in flight on different streams never share (or re-zero) a counter */
    ka.queue = c->queue + (c->launches % kQueueSlots);
    ka.out = *out;
    ka.out_pitch = out_pitch > 0 ? out_pitch : out->out_cap;
    ka.com_pitch = com_pitch > 0 ? com_pitch : out->com_cap;
    if (ka.out.out_cap == 0) ka.out.out_x = ka.out.out_y = nullptr;
    if (ka.out.com_cap == 0) ka.out.com_break = ka.out.com_step = ka.out.com_coef = nullptr;
    if (pipe && pipe->shift > 0) {
        ka.chunk_shift = pipe->shift;
        ka.upload_ready = pipe->upload_ready;
        ka.chunk_done = pipe->chunk_done;
        ka.chunk_flag = pipe->chunk_flag;
    }

    /* large states (ode_split.cuh): Dopri5 / Dop853 with one trajectory per group of three lanes */
    const void* split_fn = nullptr;
    const SystemEntry::SplitKernels* split = nullptr;
    if (!s.jit && cfg->method != ODE_B200_RK4 && c->lane_split) {
        split = &kBuiltins[sys_id]->split[cfg->method == ODE_B200_DOP853 ? 1 : 0];
        split_fn = split->fn[0][cfg->out_type];
    }
    const int block = split_fn ? split->threads[cfg->out_type]
                               : block_threads_for(s.dim, cfg->method, records_dense(cfg->method, cfg->out_type));
    const long long traj_per_block = split_fn ? (long long)(block / 32) * kBuiltins[sys_id]->split_traj_per_warp : block;
    const long long blocks_all = (n + traj_per_block - 1) / traj_per_block;

    /* `accepted % n_stiff` panics in the reference when n_stiff == 0 */
    if (cfg->method != ODE_B200_RK4 && cfg->n_stiff == 0) {
        fill_unsupported_kernel<<<(unsigned)blocks_all, block, 0, stream>>>(ka, s.dim);
        c->launches++;
        CUDA_TRY(cudaGetLastError());
        return ODE_B200_SUCCESS;
    }

    /* Staging of recorded results through shared memory and TMA bulk copies (ode_kernel_args.h) pays
     * when a lane pushes about one sample per attempted step or more: every Rk4 step, every accepted
     * step (Sparse / Continuous), or Dense with many samples per trajectory; with a few dozen samples
     * per trajectory the shared memory costs more occupancy than the stores cost time. The bulk copies
     * need 16-byte aligned rows: an even row pitch and aligned base pointers (always true for the
     * host-buffer entry, which pads its device rows). */
    bool stage = ka.out.out_x != nullptr && split_fn == nullptr;
    if (stage && cfg->method != ODE_B200_RK4 && cfg->out_type == ODE_B200_OUT_DENSE) {
        const double samples = cfg->dx > 0.0 ? fabs(cfg->x_end - cfg->x0) / cfg->dx : 0.0;
        stage = samples >= 1000.0; /* measured: 201 samples 10.6 ms unstaged vs 12.2 staged; 2001 samples 37.4 vs 18.3 */
    }
    if (stage && ((ka.out_pitch & 1) != 0 || ((uintptr_t)ka.out.out_x & 15u) != 0 || ((uintptr_t)ka.out.out_y & 15u) != 0))
        stage = false;
    /* Coefficient vectors per interval of the continuous output model of this (method, out_type): the
     * STAGE kernels of Dopri5/Continuous and Dop853/Continuous8 carry the interval ring in their
     * compile-time shared-memory layout (Stepper::kCS) whether or not this call records the model,
     * so the ring is sized from the same rule even when com_cap == 0 (com_break == NULL). */
    const int com_m = (cfg->method == ODE_B200_DOPRI5 && cfg->out_type == ODE_B200_OUT_CONTINUOUS)   ? 5
                      : (cfg->method == ODE_B200_DOP853 && cfg->out_type == ODE_B200_OUT_CONTINUOUS8) ? 8
                                                                                                      : 0;
    if (stage && com_m > 0 && ka.out.com_break != nullptr &&
        ((ka.com_pitch & 1) != 0 || ((uintptr_t)ka.out.com_break & 15u) != 0 || ((uintptr_t)ka.out.com_step & 15u) != 0 ||
         ((uintptr_t)ka.out.com_coef & 15u) != 0))
        stage = false;
    size_t dyn_smem = stage ? (size_t)stage_smem_doubles(s.dim, block, com_m, cfg->method) * sizeof(double) : 0;
    if (split_fn && ka.out.out_x != nullptr && ((uintptr_t)ka.out.out_y & 15u) == 0 &&
        (ka.out.com_break == nullptr || ((uintptr_t)ka.out.com_coef & 15u) == 0)) {
        /* the split kernels stage recorded results per trajectory slot (whole rows of the result arrays:
         * 16-byte aligned whenever the arrays are) */
        split_fn = split->fn[1][cfg->out_type];
        dyn_smem = (size_t)(block / 32) * kBuiltins[sys_id]->split_traj_per_warp * split->slot_bytes;
    }
    ka.tensor_out = 0;
    if (stage && cfg->method == ODE_B200_RK4 && c->tensor_stores) ka.tensor_out = make_result_maps(&ka, s.dim);
    CUDA_TRY(cudaMemsetAsync(ka.queue, 0, sizeof(unsigned long long), stream));
    if (stream == c->stream) CUDA_TRY(cudaEventRecord(c->ev0, stream));
    if (!s.jit) {
        const void* fn = split_fn ? split_fn : kBuiltins[sys_id]->kernels[stage ? 1 : 0][cfg->method][cfg->out_type];
        if (ka.chunk_shift > 0) { /* host pipeline: the instantiation with its device side (never staged: see integrate_range) */
            if (stage || dyn_smem != 0) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "internal: pipelined launch of a staged kernel");
            fn = split_fn ? split->pipe_fn[cfg->out_type]
                          : kBuiltins[sys_id]->pipe_kernels[cfg->method][cfg->out_type];
        }
        if (dyn_smem > 32 * 1024) /* static tables + dynamic rings may exceed the 48 KB default together */
            CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
        int per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block, dyn_smem));
        if (per_sm < 1) return fail(ODE_B200_ERR_CUDA, "kernel does not fit on an SM");
        long long grid = blocks_all;
        if (c->use_queue) grid = std::min<long long>(blocks_all, (long long)c->num_sms * per_sm);
        void* args[] = {&ka};
        CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(block), args, dyn_smem, stream));
    } else {
        std::string err;
        if (!jit_launch(sys_id - kNumBuiltins, c->device, c->num_sms, cfg->method, cfg->out_type, stage, &ka, sizeof(ka),
                        block, blocks_all, c->use_queue, dyn_smem, (void*)stream, &err))
            return fail(ODE_B200_ERR_COMPILE, err);
    }
    if (stream == c->stream) {
        CUDA_TRY(cudaEventRecord(c->ev1, stream));
        c->timed = true;
    }
    c->launches++;
    return ODE_B200_SUCCESS;
}

extern "C" int ode_b200_integrate_batch_device(ode_b200_ctx* c, int sys_id, const ode_b200_config* cfg,
                                               int64_t n, const double* y0, const double* params, int per_traj,
                                               const ode_b200_outputs* out, void* cuda_stream) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (!cfg || !out) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "cfg / out is NULL");
    cudaStream_t st = (cudaStream_t)cuda_stream; /* NULL = the CUDA default stream */
    return launch_device(c, sys_id, cfg, n, y0, params, per_traj, out, out->out_cap, out->com_cap, st);
}

/* device alias of a pinned, device-mapped host allocation (nullptr for pageable memory) */
static void* mapped_host_ptr(const void* host) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (at.type != cudaMemoryTypeHost || at.devicePointer == nullptr) return nullptr;
    return at.devicePointer;
}

/*
 * Host-buffer path for the trajectory range [lo, hi) of a batch of n_total: stage the
 * range's column block of the SoA inputs, run, copy the range's outputs back into the
 * caller's full-size arrays. [0, n_total) on one context is ode_b200_integrate_batch;
 * one range per GPU is the multi-GPU shard.
 */
static int integrate_range_impl(ode_b200_ctx* c, int sys_id, const ode_b200_config* cfg, int64_t n_total, int64_t lo,
                                int64_t hi, const double* y0, const double* params, int per_traj,
                                const ode_b200_outputs* out);

/* Never return while copies or zero-copy kernel stores that target the caller's arrays may still be
 * in flight: an error in the middle of the enqueue sequence drains the stream first. */
static int integrate_range(ode_b200_ctx* c, int sys_id, const ode_b200_config* cfg, int64_t n_total, int64_t lo,
                           int64_t hi, const double* y0, const double* params, int per_traj,
                           const ode_b200_outputs* out) {
    const int rc = integrate_range_impl(c, sys_id, cfg, n_total, lo, hi, y0, params, per_traj, out);
    if (rc != ODE_B200_SUCCESS && c && c->stream) {
        const std::string keep = g_last_error;
        cudaStreamSynchronize(c->up_stream);
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->copy_stream);
        cudaGetLastError();
        g_last_error = keep;
    }
    return rc;
}

static int integrate_range_impl(ode_b200_ctx* c, int sys_id, const ode_b200_config* cfg, int64_t n_total, int64_t lo,
                                int64_t hi, const double* y0, const double* params, int per_traj,
                                const ode_b200_outputs* out) {
    SysInfo s;
    if (!sys_info(sys_id, &s)) return fail(ODE_B200_ERR_UNKNOWN_SYSTEM, "bad sys_id");
    if (int rc = validate(cfg, s, n_total, y0, params, out)) return rc;
    const int64_t n = hi - lo;
    if (n <= 0) return ODE_B200_SUCCESS;
    CUDA_TRY(cudaSetDevice(c->device));
    const size_t d = (size_t)s.dim, np = (size_t)s.n_params;
    const int m = cfg->method == ODE_B200_DOPRI5 ? 5 : 8;
    cudaStream_t st = c->stream;

    /* Host pipeline (ode_kernel_args.h): with at least two chunks of work and pinned per-trajectory result
     * arrays the kernel is launched first, the inputs are uploaded chunk by chunk behind it, and finished
     * chunks are copied back while the rest is still being integrated; what is left exposed is the first
     * chunk's upload and the last chunk's copy. Measured on B200, ms per call (kernel alone / one-shot copies /
     * zero-copy result stores / pipeline): see DESIGN.md section 5. */
    const int shift = c->pipeline_shift;
    const int64_t n_chunks = shift > 0 ? (n + ((int64_t)1 << shift) - 1) >> shift : 0;
    /* ... for calls that return per-trajectory results only (the per-step arrays are copied after the kernel and
     * dominate such a call anyway) of built-in systems (the PIPE instantiations are compiled ahead of time) */
    bool pipe = shift > 0 && n_chunks >= 2 && n_chunks <= kPipeMaxChunks && !s.jit && out->out_cap == 0 &&
                out->com_cap == 0 && !(cfg->method != ODE_B200_RK4 && cfg->n_stiff == 0);
    if (pipe) {
        const void* ptrs[] = {out->y_final, out->x_final, out->h_next,    out->num_eval,  out->accepted, out->rejected,
                              out->n_step,  out->status,  out->out_count, out->com_count, out->com_lower};
        bool any = false;
        for (const void* q : ptrs)
            if (q) {
                any = true;
                cudaPointerAttributes at;
                if (cudaPointerGetAttributes(&at, q) != cudaSuccess || at.type != cudaMemoryTypeHost) pipe = false;
                cudaGetLastError();
            }
        pipe = pipe && any;
    }
    PipeArgs pa;
    if (pipe) {
        if (ensure(c, c->pipe_dev, (size_t)n_chunks * 2 * sizeof(unsigned))) return ODE_B200_ERR_CUDA;
        pa.shift = shift;
        pa.upload_ready = (const unsigned*)c->pipe_dev.p;
        pa.chunk_done = (unsigned*)c->pipe_dev.p + n_chunks;
        pa.chunk_flag = c->pipe_flags_dev;
        for (int64_t k = 0; k < n_chunks; ++k) c->pipe_flags_host[k] = 0u;
        CUDA_TRY(cudaMemsetAsync(c->pipe_dev.p, 0, (size_t)n_chunks * 2 * sizeof(unsigned), st));
        CUDA_TRY(cudaEventRecord(c->ev_pipe, st));
        CUDA_TRY(cudaStreamWaitEvent(c->up_stream, c->ev_pipe, 0));
    }

    if (ensure(c, c->y0, d * n * 8)) return ODE_B200_ERR_CUDA;
    if (!pipe) CUDA_TRY(cudaMemcpy2DAsync(c->y0.p, n * 8, y0 + lo, n_total * 8, n * 8, d, cudaMemcpyHostToDevice, st));
    if (np) {
        if (per_traj) {
            if (ensure(c, c->params, np * n * 8)) return ODE_B200_ERR_CUDA;
            if (!pipe)
                CUDA_TRY(cudaMemcpy2DAsync(c->params.p, n * 8, params + lo, n_total * 8, n * 8, np,
                                           cudaMemcpyHostToDevice, st));
        } else {
            if (ensure(c, c->params, np * 8)) return ODE_B200_ERR_CUDA;
            CUDA_TRY(cudaMemcpyAsync(c->params.p, params, np * 8, cudaMemcpyHostToDevice, st));
        }
    }
    ode_b200_outputs dv;
    memset(&dv, 0, sizeof(dv));
    /* Per-trajectory results (a few stores per trajectory, written once when it retires) go
     * STRAIGHT into the caller's arrays when those are pinned, device-mapped host memory: the
     * stores trickle over PCIe while the kernel runs and no device-to-host copy phase is left
     * (bench.py: 1.2 ms of a 30.3 ms end-to-end pass). Pageable memory, sharded ranges of the SoA
     * state and the large per-step arrays take the staged copy. When several GPUs share the host's
     * PCIe the fine-grained stores lose against one DMA burst per array (8 ranks: 37.5 ms per pass
     * against 33 ms); ode_b200_set_zero_copy(ctx, 0) selects the staged copy. */
    unsigned zero_copy = 0;
    int zc_bit = 0;
#define WANT(field, buf, bytes)                                   \
    if (out->field) {                                             \
        if (ensure(c, c->buf, (bytes))) return ODE_B200_ERR_CUDA; \
        dv.field = (decltype(dv.field))c->buf.p;                  \
    }
#define WANT_PT(field, buf, elem_bytes)                                                        \
    {                                                                                          \
        ++zc_bit;                                                                              \
        if (out->field) {                                                                      \
            void* mp = (c->zero_copy && !pipe) ? mapped_host_ptr(out->field) : nullptr;        \
            if (mp) {                                                                          \
                dv.field = (decltype(dv.field))((char*)mp + (size_t)lo * (elem_bytes));        \
                zero_copy |= 1u << zc_bit;                                                     \
            } else {                                                                           \
                if (ensure(c, c->buf, (size_t)n * (elem_bytes))) return ODE_B200_ERR_CUDA;     \
                dv.field = (decltype(dv.field))c->buf.p;                                       \
            }                                                                                  \
        }                                                                                      \
    }
    if (out->y_final) {
        void* mp = (c->zero_copy && !pipe && lo == 0 && n == n_total) ? mapped_host_ptr(out->y_final) : nullptr;
        if (mp) {
            dv.y_final = (double*)mp;
            zero_copy |= 1u;
        } else {
            WANT(y_final, y_final, d * n * 8)
        }
    }
    WANT_PT(x_final, x_final, 8)
    WANT_PT(h_next, h_next, 8)
    WANT_PT(num_eval, num_eval, 4)
    WANT_PT(accepted, accepted, 4)
    WANT_PT(rejected, rejected, 4)
    WANT_PT(n_step, n_step, 4)
    WANT_PT(status, status, 4)
    WANT_PT(out_count, out_count, 4)
    WANT_PT(com_count, com_count, 4)
    WANT_PT(com_lower, com_lower, 8)
#undef WANT_PT
    /* device rows of the per-step results are padded to a multiple of 4 samples (whole 32-byte
     * sectors per staged group, 16-byte aligned bulk copies); the copy back strips the padding */
    const size_t pitch = ((size_t)out->out_cap + 3) & ~(size_t)3;
    if (out->out_cap > 0) {
        dv.out_cap = out->out_cap;
        WANT(out_x, out_x, (size_t)n * pitch * 8)
        WANT(out_y, out_y, (size_t)n * pitch * d * 8)
    }
    const size_t cpitch = ((size_t)out->com_cap + 3) & ~(size_t)3;
    if (out->com_cap > 0) {
        dv.com_cap = out->com_cap;
        WANT(com_break, com_break, (size_t)n * cpitch * 8)
        WANT(com_step, com_step, (size_t)n * cpitch * 8)
        WANT(com_coef, com_coef, (size_t)n * cpitch * m * d * 8)
    }
#undef WANT
    if (int rc = launch_device(c, sys_id, cfg, n, (const double*)c->y0.p, (const double*)c->params.p, per_traj,
                               &dv, (long long)pitch, (long long)cpitch, st, pipe ? &pa : nullptr))
        return rc;

    if (pipe) {
        /* ---- upload behind the running kernel: chunk data, then the chunk's ready word (same stream: in order) */
        const int64_t ch = (int64_t)1 << shift;
        for (int64_t k = 0; k < n_chunks; ++k) {
            const int64_t first = k << shift, cnt = std::min(ch, n - first);
            CUDA_TRY(cudaMemcpy2DAsync((double*)c->y0.p + first, n * 8, y0 + lo + first, n_total * 8, cnt * 8, d,
                                       cudaMemcpyHostToDevice, c->up_stream));
            if (np && per_traj)
                CUDA_TRY(cudaMemcpy2DAsync((double*)c->params.p + first, n * 8, params + lo + first, n_total * 8, cnt * 8,
                                           np, cudaMemcpyHostToDevice, c->up_stream));
            CUDA_TRY(cudaMemcpyAsync((unsigned*)c->pipe_dev.p + k, c->pipe_flags_host + kPipeMaxChunks, sizeof(unsigned),
                                     cudaMemcpyHostToDevice, c->up_stream));
        }
        /* ---- drain: copy a chunk's per-trajectory results back as soon as the kernel has flagged it */
        auto drain = [&](int64_t k) -> int {
            const int64_t first = k << shift, cnt = std::min(ch, n - first);
            cudaStream_t cs = c->copy_stream;
            if (out->y_final)
                CUDA_TRY(cudaMemcpy2DAsync(out->y_final + lo + first, n_total * 8, dv.y_final + first, n * 8, cnt * 8, d,
                                           cudaMemcpyDeviceToHost, cs));
#define DRAIN(field, elem_bytes)                                                                               \
    if (out->field)                                                                                            \
        CUDA_TRY(cudaMemcpyAsync((char*)out->field + (size_t)(lo + first) * (elem_bytes),                      \
                                 (const char*)dv.field + (size_t)first * (elem_bytes), (size_t)cnt * (elem_bytes), \
                                 cudaMemcpyDeviceToHost, cs));
            DRAIN(x_final, 8)
            DRAIN(h_next, 8)
            DRAIN(num_eval, 4)
            DRAIN(accepted, 4)
            DRAIN(rejected, 4)
            DRAIN(n_step, 4)
            DRAIN(status, 4)
            DRAIN(out_count, 4)
            DRAIN(com_count, 4)
            DRAIN(com_lower, 8)
#undef DRAIN
            return 0;
        };
        volatile unsigned* flags = c->pipe_flags_host;
        int64_t next = 0;
        unsigned spins = 0;
        while (next < n_chunks) {
            if (flags[next] != 0u) {
                if (int rc = drain(next)) return rc;
                ++next;
                continue;
            }
            if ((++spins & 1023u) == 0u) {
                const cudaError_t q = cudaStreamQuery(st);
                if (q == cudaSuccess) { /* the kernel is done: whatever is left is complete */
                    for (; next < n_chunks; ++next)
                        if (int rc = drain(next)) return rc;
                    break;
                }
                if (q != cudaErrorNotReady) return fail(ODE_B200_ERR_CUDA, std::string("step-loop kernel: ") + cudaGetErrorString(q));
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        CUDA_TRY(cudaStreamSynchronize(c->up_stream));
    }

    if (out->y_final && !(zero_copy & 1u) && !pipe)
        CUDA_TRY(cudaMemcpy2DAsync(out->y_final + lo, n_total * 8, dv.y_final, n * 8, n * 8, d,
                                   cudaMemcpyDeviceToHost, st));
#define BACK(field, elem_bytes, per)                                                                      \
    if (out->field)                                                                                       \
        CUDA_TRY(cudaMemcpyAsync((char*)out->field + (size_t)lo * (per) * (elem_bytes), dv.field,         \
                                 (size_t)n * (per) * (elem_bytes), cudaMemcpyDeviceToHost, st));
    zc_bit = 0;
#define BACK_PT(field, elem_bytes)                 \
    {                                              \
        ++zc_bit;                                  \
        if (!(zero_copy & (1u << zc_bit))) {       \
            BACK(field, elem_bytes, 1)             \
        }                                          \
    }
    if (!pipe) {
    BACK_PT(x_final, 8)
    BACK_PT(h_next, 8)
    BACK_PT(num_eval, 4)
    BACK_PT(accepted, 4)
    BACK_PT(rejected, 4)
    BACK_PT(n_step, 4)
    BACK_PT(status, 4)
    BACK_PT(out_count, 4)
    BACK_PT(com_count, 4)
    BACK_PT(com_lower, 8)
    }
#undef BACK_PT
    if (out->out_cap > 0) {
        const size_t cap = (size_t)out->out_cap;
        if (pitch == cap) {
            BACK(out_x, 8, cap)
            BACK(out_y, 8, cap * d)
        } else {
            CUDA_TRY(cudaMemcpy2DAsync(out->out_x + (size_t)lo * cap, cap * 8, dv.out_x, pitch * 8, cap * 8, (size_t)n,
                                       cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpy2DAsync(out->out_y + (size_t)lo * cap * d, cap * d * 8, dv.out_y, pitch * d * 8,
                                       cap * d * 8, (size_t)n, cudaMemcpyDeviceToHost, st));
        }
    }
    if (out->com_cap > 0) {
        const size_t cap = (size_t)out->com_cap;
        if (cpitch == cap) {
            BACK(com_break, 8, cap)
            BACK(com_step, 8, cap)
            BACK(com_coef, 8, cap * m * d)
        } else {
            CUDA_TRY(cudaMemcpy2DAsync(out->com_break + (size_t)lo * cap, cap * 8, dv.com_break, cpitch * 8, cap * 8,
                                       (size_t)n, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpy2DAsync(out->com_step + (size_t)lo * cap, cap * 8, dv.com_step, cpitch * 8, cap * 8,
                                       (size_t)n, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaMemcpy2DAsync(out->com_coef + (size_t)lo * cap * m * d, cap * m * d * 8, dv.com_coef,
                                       cpitch * m * d * 8, cap * m * d * 8, (size_t)n, cudaMemcpyDeviceToHost, st));
        }
    }
#undef BACK
    CUDA_TRY(cudaStreamSynchronize(st));
    if (pipe) CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
    return ODE_B200_SUCCESS;
}

extern "C" int ode_b200_integrate_batch(ode_b200_ctx* c, int sys_id, const ode_b200_config* cfg, int64_t n,
                                        const double* y0, const double* params, int per_traj,
                                        const ode_b200_outputs* out) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
    return integrate_range(c, sys_id, cfg, n, 0, n, y0, params, per_traj, out);
}

/* Rk4 with T = f32 (see include/ode_b200.h) */
/* ---- T = f32: constructors (every literal through T::from, every operation in f32) ------------------ */
static const float kEpsF64AsF32 = (float)2.220446049250313e-16; /* T::from(f64::EPSILON) */

/* Dopri5::<f32>::new (dopri5.rs:14-27,76-105: uround = T::epsilon(), Q3); Dop853::<f32>::new (dop853.rs:14-27,76-109) */
extern "C" int ode_b200_config_new_f32(int method, float x, float x_end, float dx, float rtol, float atol,
                                       ode_b200_config_f32* c) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "config is NULL");
    memset(c, 0, sizeof(*c));
    c->method = method;
    c->out_type = ODE_B200_OUT_DENSE;
    c->x0 = x;
    c->x_end = x_end;
    c->dx = dx;
    c->rtol = rtol;
    c->atol = atol;
    c->safety_factor = (float)0.9;
    c->h_max = x_end - x;
    c->h0 = 0.0f;
    c->n_max = 100000;
    c->n_stiff = 1000;
    if (method == ODE_B200_DOPRI5) {
        c->alpha = (float)(0.2 - 0.04 * 0.75); /* T::from(0.2 - 0.04 * 0.75): the f64 value, then rounded */
        c->beta = (float)0.04;
        c->fac_max = (float)10.0;
        c->fac_min = (float)0.2;
        c->uround = 1.1920929e-07f; /* f32::EPSILON = 2^-23 */
    } else if (method == ODE_B200_DOP853) {
        c->alpha = 1.0f / (float)8.0;
        c->beta = 0.0f;
        c->fac_max = (float)6.0;
        c->fac_min = (float)0.333;
        c->uround = kEpsF64AsF32;
    } else {
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "config_new_f32: method must be DOPRI5 or DOP853");
    }
    return ODE_B200_SUCCESS;
}

/* Dopri5::<f32>::from_param (dopri5.rs:129-184); Dop853::<f32>::from_param (dop853.rs:133-192) */
extern "C" int ode_b200_config_from_param_f32(int method, float x, float x_end, float dx, float rtol, float atol,
                                              float safety_factor, float beta, float fac_min, float fac_max,
                                              float h_max, float h, uint32_t n_max, uint32_t n_stiff, int out_type,
                                              ode_b200_config_f32* c) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "config is NULL");
    if (out_type < ODE_B200_OUT_DENSE || out_type > ODE_B200_OUT_CONTINUOUS)
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "config_from_param_f32: bad out_type");
    memset(c, 0, sizeof(*c));
    c->method = method;
    c->out_type = out_type;
    c->x0 = x;
    c->x_end = x_end;
    c->dx = dx;
    c->rtol = rtol;
    c->atol = atol;
    c->safety_factor = safety_factor;
    c->beta = beta;
    c->fac_min = fac_min;
    c->fac_max = fac_max;
    c->h_max = h_max;
    c->h0 = h;
    c->uround = kEpsF64AsF32;
    c->n_max = n_max;
    c->n_stiff = n_stiff;
    volatile float b = beta; /* two roundings: the product, then the difference */
    if (method == ODE_B200_DOPRI5) {
        volatile float t = b * (float)0.75;
        c->alpha = (float)0.2 - t;
    } else if (method == ODE_B200_DOP853) {
        volatile float t = b * (float)0.2;
        c->alpha = 1.0f / (float)8.0 - t;
    } else {
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "config_from_param_f32: method must be DOPRI5 or DOP853");
    }
    return ODE_B200_SUCCESS;
}

/* Rk4::<f32>::new (rk4.rs:42-68) */
extern "C" int ode_b200_config_rk4_f32(float x, float x_end, float step_size, ode_b200_config_f32* c) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "config is NULL");
    memset(c, 0, sizeof(*c));
    c->method = ODE_B200_RK4;
    c->out_type = ODE_B200_OUT_SPARSE;
    c->x0 = x;
    c->x_end = x_end;
    c->step_size = step_size;
    return ODE_B200_SUCCESS;
}

/* ---- T = f32: the batched entry (see include/ode_b200.h) ------------------------------------------- */
static int integrate_f32_impl(ode_b200_ctx* c, int sys_id, const ode_b200_config_f32* cfg, int64_t n, const float* y0,
                              const float* params, int per_traj, const ode_b200_outputs_f32* out);

extern "C" int ode_b200_integrate_batch_f32(ode_b200_ctx* c, int sys_id, const ode_b200_config_f32* cfg, int64_t n,
                                            const float* y0, const float* params, int per_traj,
                                            const ode_b200_outputs_f32* out) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
    const int rc = integrate_f32_impl(c, sys_id, cfg, n, y0, params, per_traj, out);
    if (rc != ODE_B200_SUCCESS && c->stream) { /* see integrate_range */
        const std::string keep = g_last_error;
        cudaStreamSynchronize(c->stream);
        cudaGetLastError();
        g_last_error = keep;
    }
    return rc;
}

extern "C" int ode_b200_rk4_f32_integrate_batch(ode_b200_ctx* c, int sys_id, float x, float x_end, float step_size,
                                                int64_t n, const float* y0, const float* params, int per_traj,
                                                const ode_b200_outputs_f32* out) {
    ode_b200_config_f32 cfg;
    ode_b200_config_rk4_f32(x, x_end, step_size, &cfg);
    return ode_b200_integrate_batch_f32(c, sys_id, &cfg, n, y0, params, per_traj, out);
}

static int integrate_f32_impl(ode_b200_ctx* c, int sys_id, const ode_b200_config_f32* cfg, int64_t n, const float* y0,
                              const float* params, int per_traj, const ode_b200_outputs_f32* out) {
    if (sys_id < 0 || sys_id >= kNumBuiltins) return fail(ODE_B200_ERR_UNKNOWN_SYSTEM, "f32 needs a built-in system");
    if (!cfg || !out) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "f32: cfg / out is NULL");
    if (cfg->method < ODE_B200_RK4 || cfg->method > ODE_B200_DOP853) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "bad method");
    if (cfg->out_type < ODE_B200_OUT_DENSE || cfg->out_type > ODE_B200_OUT_CONTINUOUS)
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "f32: out_type must be Dense, Sparse or Continuous");
    const SystemEntry* e = kBuiltins[sys_id];
    const void* fn = e->f32_kernels[cfg->method];
    if (!fn) return fail(ODE_B200_ERR_UNKNOWN_SYSTEM, std::string("system '") + e->name + "' has no f32 instantiation");
    if (n < 0 || (n > 0 && !y0) || (n > 0 && e->n_params > 0 && !params) || out->out_cap < 0 || out->com_cap < 0 ||
        (out->out_cap > 0 && (!out->out_x || !out->out_y)) ||
        (out->com_cap > 0 && (!out->com_break || !out->com_step || !out->com_coef)))
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "integrate_batch_f32: bad arguments");
    if (n == 0) return ODE_B200_SUCCESS;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t d = (size_t)e->dim, np = (size_t)e->n_params;
    if (ensure(c, c->y0, d * n * 4)) return ODE_B200_ERR_CUDA;
    CUDA_TRY(cudaMemcpyAsync(c->y0.p, y0, d * n * 4, cudaMemcpyHostToDevice, st));
    if (np) {
        const size_t pb = (per_traj ? np * n : np) * 4;
        if (ensure(c, c->params, pb)) return ODE_B200_ERR_CUDA;
        CUDA_TRY(cudaMemcpyAsync(c->params.p, params, pb, cudaMemcpyHostToDevice, st));
    }
    KernelArgsF32 ka;
    memset(&ka, 0, sizeof(ka));
    ka.cfg = *cfg;
    ka.n_traj = n;
    ka.y0 = (const float*)c->y0.p;
    ka.params = (const float*)c->params.p;
    ka.params_per_traj = per_traj ? 1 : 0;
    ka.use_queue = c->use_queue;
    ka.queue = c->queue + (c->launches % kQueueSlots);
    ode_b200_outputs_f32& dv = ka.out;
    const size_t cap = (size_t)out->out_cap, ccap = (size_t)out->com_cap;
#define WANT32(field, buf, bytes)                                 \
    if (out->field) {                                             \
        if (ensure(c, c->buf, (bytes))) return ODE_B200_ERR_CUDA; \
        dv.field = (decltype(dv.field))c->buf.p;                  \
    }
    WANT32(y_final, y_final, d * n * 4)
    WANT32(x_final, x_final, n * 4)
    WANT32(h_next, h_next, n * 4)
    WANT32(num_eval, num_eval, n * 4)
    WANT32(accepted, accepted, n * 4)
    WANT32(rejected, rejected, n * 4)
    WANT32(n_step, n_step, n * 4)
    WANT32(status, status, n * 4)
    WANT32(out_count, out_count, n * 4)
    WANT32(com_count, com_count, n * 4)
    WANT32(com_lower, com_lower, n * 4)
    if (cap > 0) {
        dv.out_cap = out->out_cap;
        WANT32(out_x, out_x, (size_t)n * cap * 4)
        WANT32(out_y, out_y, (size_t)n * cap * d * 4)
    }
    if (ccap > 0) {
        dv.com_cap = out->com_cap;
        WANT32(com_break, com_break, (size_t)n * ccap * 4)
        WANT32(com_step, com_step, (size_t)n * ccap * 4)
        WANT32(com_coef, com_coef, (size_t)n * ccap * 5 * d * 4)
    }
#undef WANT32
    const int block = 128; /* __launch_bounds__ of ode_f32_step_loop_kernel */
    const long long blocks_all = (n + block - 1) / block;
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block, 0));
    long long grid = blocks_all;
    if (c->use_queue) grid = std::min<long long>(blocks_all, (long long)c->num_sms * std::max(per_sm, 1));
    CUDA_TRY(cudaMemsetAsync(ka.queue, 0, sizeof(unsigned long long), st));
    CUDA_TRY(cudaEventRecord(c->ev0, st));
    void* args[] = {&ka};
    CUDA_TRY(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(block), args, 0, st));
    CUDA_TRY(cudaEventRecord(c->ev1, st));
    c->timed = true;
    c->launches++;
#define BACK32(field, bytes) \
    if (out->field) CUDA_TRY(cudaMemcpyAsync(out->field, dv.field, (bytes), cudaMemcpyDeviceToHost, st));
    BACK32(y_final, d * n * 4)
    BACK32(x_final, n * 4)
    BACK32(h_next, n * 4)
    BACK32(num_eval, n * 4)
    BACK32(accepted, n * 4)
    BACK32(rejected, n * 4)
    BACK32(n_step, n * 4)
    BACK32(status, n * 4)
    BACK32(out_count, n * 4)
    BACK32(com_count, n * 4)
    BACK32(com_lower, n * 4)
    if (cap > 0) {
        BACK32(out_x, (size_t)n * cap * 4)
        BACK32(out_y, (size_t)n * cap * d * 4)
    }
    if (ccap > 0) {
        BACK32(com_break, (size_t)n * ccap * 4)
        BACK32(com_step, (size_t)n * ccap * 4)
        BACK32(com_coef, (size_t)n * ccap * 5 * d * 4)
    }
#undef BACK32
    CUDA_TRY(cudaStreamSynchronize(st));
    return ODE_B200_SUCCESS;
}

/* contiguous index ranges, one host thread + context + stream per GPU, no collective */
extern "C" int ode_b200_integrate_batch_multi(int sys_id, const ode_b200_config* cfg, int64_t n,
                                              const double* y0, const double* params, int per_traj, int n_gpus,
                                              const ode_b200_outputs* out) {
    const int avail = ode_b200_device_count();
    if (avail <= 0) return fail(ODE_B200_ERR_NO_DEVICE, "no CUDA device available (no CPU fallback)");
    if (n_gpus < 1 || n_gpus > avail) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "n_gpus out of range");
    /* this entry owns one internal context per GPU: concurrent callers take turns (callers that want
     * to overlap use their own contexts and ode_b200_integrate_batch per GPU) */
    static std::mutex mu;
    static std::vector<ode_b200_ctx*> ctxs;
    std::lock_guard<std::mutex> whole_call(mu);
    while ((int)ctxs.size() < n_gpus) {
        ode_b200_ctx* c = ode_b200_create((int)ctxs.size());
        if (!c) return ODE_B200_ERR_CUDA;
        ctxs.push_back(c);
    }
    /* up to 4 GPUs store their per-trajectory results straight into pinned host arrays; with more of
     * them sharing the host's PCIe one DMA copy per array is faster (measured with 8 ranks: 33.0 against
     * 37.5 ms per pass; with 4: 32.6 against 30.0) */
    for (int g = 0; g < n_gpus; ++g) ctxs[g]->zero_copy = n_gpus <= 4 ? 1 : 0;
    const int64_t per = (n + n_gpus - 1) / n_gpus;
    std::vector<int> rcs(n_gpus, 0);
    std::vector<std::string> errs(n_gpus);
    std::vector<std::thread> th;
    for (int g = 0; g < n_gpus; ++g) {
        th.emplace_back([&, g]() {
            const int64_t lo = std::min<int64_t>(n, per * g), hi = std::min<int64_t>(n, per * (g + 1));
            rcs[g] = integrate_range(ctxs[g], sys_id, cfg, n, lo, hi, y0, params, per_traj, out);
            if (rcs[g]) errs[g] = g_last_error;
        });
    }
    for (auto& t : th) t.join();
    for (int g = 0; g < n_gpus; ++g)
        if (rcs[g]) return fail(rcs[g], "gpu " + std::to_string(g) + ": " + errs[g]);
    return ODE_B200_SUCCESS;
}

/* ------------------------------------------------------------------ FP64 peak probe */
/* 8 independent DFMA chains per thread; counts FP64-pipe thread-instructions per second */
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6,
           a7 = seed + 7;
    const double m = 0.9999999, c = 1e-9 * threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fma(a0, m, c);
            a1 = fma(a1, m, c);
            a2 = fma(a2, m, c);
            a3 = fma(a3, m, c);
            a4 = fma(a4, m, c);
            a5 = fma(a5, m, c);
            a6 = fma(a6, m, c);
            a7 = fma(a7, m, c);
        }
    }
    double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 123.456) sink[0] = r;
}

extern "C" int ode_b200_measure_fp64_peak(int device, double* inst_per_s, double* sm_mhz_est) {
    if (ode_b200_device_count() <= 0) return fail(ODE_B200_ERR_NO_DEVICE, "no CUDA device");
    CUDA_TRY(cudaSetDevice(device));
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    double* sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const int iters = 4096, grid = sms * 8, block = 256;
    double best = 0.0;
    cudaError_t err = cudaMalloc(&sink, 8);
    if (err == cudaSuccess) err = cudaEventCreate(&e0);
    if (err == cudaSuccess) err = cudaEventCreate(&e1);
    for (int rep = 0; rep < 5 && err == cudaSuccess; ++rep) {
        float ms = 0.f;
        err = cudaEventRecord(e0);
        if (err == cudaSuccess) {
            fp64_peak_kernel<<<grid, block>>>(sink, iters, 1.0 + rep);
            err = cudaGetLastError();
        }
        if (err == cudaSuccess) err = cudaEventRecord(e1);
        if (err == cudaSuccess) err = cudaEventSynchronize(e1);
        if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
        const double inst = (double)grid * block * (double)iters * 16.0 * 8.0;
        if (err == cudaSuccess && rep > 0) best = std::max(best, inst / (ms * 1e-3));
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (sink) cudaFree(sink);
    if (err != cudaSuccess) return fail(ODE_B200_ERR_CUDA, std::string("fp64 peak probe: ") + cudaGetErrorString(err));
    if (inst_per_s) *inst_per_s = best;
    /* 64 FP64 lanes per SM: implied SM clock if the pipe was saturated */
    if (sm_mhz_est) *sm_mhz_est = best / (64.0 * sms) * 1e-6;
    return ODE_B200_SUCCESS;
}

/* ------------------------------------------------------------------ COM::evaluate */
/*
 * ContinuousOutputModel::evaluate (continuous_output_model.rs:31-56) with get_interval_index (:86-104), batched:
 * n models x nq queries. HBM-bound: per (model, query) it reads the interval's m x dim coefficients + 2 scalars and
 * writes dim + 1/8 doubles. One WARP per model (the round-1 kernel ran one thread per (model, query) and
 * binary-searched global memory with a stride of `cap` doubles between neighbouring threads): every lane takes the
 * queries lane, lane + 32, ..., finds its interval through a coarse index of the model's breakpoints in shared
 * memory, and the 32 results of a round are written as ONE contiguous block of 32 x dim doubles. Neighbouring queries fall into the same or neighbouring intervals, so the coefficient loads of a round
 * hit the same few cache lines. M and D are compile-time for the shapes the steppers produce (Horner fully unrolled);
 * other shapes take the generic instantiation (M = D = 0: runtime loops).
 * Bit parity: theta = (x - lower) / step, then r = c[m-1]; r = c[q] + r * (q even ? theta : 1 - theta), as the source.
 */
constexpr int kComWarps = 8;     /* warps (= models) per CTA */
#ifndef ODE_B200_COM_COARSE
#define ODE_B200_COM_COARSE 192
#endif
#ifndef ODE_B200_COM_BLOCKS
#define ODE_B200_COM_BLOCKS 6
#endif
constexpr int kComCoarse = ODE_B200_COM_COARSE; /* entries of a model's coarse breakpoint index in shared memory */
constexpr int kComBlocks = ODE_B200_COM_BLOCKS; /* CTAs per SM */

/* first index in [lo, hi) whose breakpoint is >= x (hi if none): the match of the reference's binary_search_by
 * (breakpoints are strictly increasing), or its insertion point */
template <class Load>
__device__ __forceinline__ unsigned com_lower_bound_(unsigned lo, unsigned hi, double x, Load load) {
    while (lo < hi) {
        const unsigned mid = lo + (hi - lo) / 2;
        if (load(mid) < x)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

template <int M, int D>
__global__ void __launch_bounds__(kComWarps * 32, kComBlocks) com_evaluate_kernel(long long n, int dim_rt, int m_rt, long long cap,
                                                                         const double* __restrict__ lower,
                                                                         const double* __restrict__ brk,
                                                                         const double* __restrict__ stp,
                                                                         const double* __restrict__ coef,
                                                                         const unsigned* __restrict__ count, long long nq,
                                                                         const double* __restrict__ xq, double* __restrict__ y,
                                                                         unsigned char* __restrict__ found,
                                                                         const unsigned* __restrict__ perm) {
    /* Two-level search. A full copy of the breakpoints in shared memory (the first version of this kernel: 12 KB
     * per warp) left 16 warps per SM and the kernel latency-bound (ncu: "long scoreboard" 9.9 stalls per issue,
     * 2.9 TB/s). Now every warp keeps a coarse index of its model -- every s-th breakpoint, s = ceil(count / 192),
     * 1.5 KB -- which narrows a query down to s breakpoints; those are read from global memory (cached: the
     * queries of a round are neighbours) for the last log2(s) steps. 48 warps per SM. */
    __shared__ double s_idx[kComWarps][kComCoarse];
    const int dim = D > 0 ? D : dim_rt, m = M > 0 ? M : m_rt;
    const unsigned lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    for (long long t = (long long)blockIdx.x * kComWarps + w; t < n; t += (long long)gridDim.x * kComWarps) {
        const unsigned cnt = (unsigned)min((long long)count[t], cap);
        const double* bp = brk + (size_t)t * cap;
        const unsigned s = (cnt + kComCoarse - 1) / kComCoarse;                 /* stride of the coarse index */
        const unsigned nc = s ? (cnt + s - 1) / s : 0;                          /* its length: entry j = bp[min((j + 1) s, cnt) - 1] */
        __syncwarp();
        for (unsigned j = lane; j < nc; j += 32u) s_idx[w][j] = __ldg(bp + min((j + 1u) * s, cnt) - 1u);
        __syncwarp();
        const double lb0 = lower[t];
        for (long long q0 = 0; q0 < nq; q0 += 32) {
            const long long qi = q0 + lane;
            const bool have = qi < nq;
            const double x = have ? xq[qi] : 0.0;
            const long long qo = have ? (perm ? (long long)perm[qi] : qi) : 0; /* host entry: xq sorted, results in the caller's order */
            unsigned lo = cnt;
            if (have && !(x < lb0) && m > 0 && x == x) { /* x < lower_bound -> None; NaN: partial_cmp().unwrap() panics */
                /* block j holds breakpoints [j s, min((j + 1) s, cnt)); its last one is s_idx[j]: the first block whose
                 * last breakpoint is >= x contains the answer */
                const unsigned j = com_lower_bound_(0u, nc, x, [&](unsigned k) { return s_idx[w][k]; });
                if (j < nc) lo = com_lower_bound_(j * s, min((j + 1u) * s, cnt) - 1u, x, [&](unsigned k) { return __ldg(bp + k); });
            }
            const bool hit = have && lo < cnt;
            if (hit) {
                double* yo = y + ((size_t)t * nq + qo) * dim;
                const double lb = lo == 0 ? lb0 : __ldg(bp + lo - 1);
                const double theta = (x - lb) / __ldg(stp + (size_t)t * cap + lo), theta1 = 1.0 - theta;
                if constexpr (M > 0 && D > 0) {
                    double r[D];
                    auto horner = [&](const double* c) {
#pragma unroll
                        for (int i = 0; i < D; ++i) r[i] = __ldg(c + (M - 1) * D + i);
#pragma unroll
                        for (int k = M - 2; k >= 0; --k)
#pragma unroll
                            for (int i = 0; i < D; ++i) r[i] = __ldg(c + k * D + i) + r[i] * ((k % 2 == 0) ? theta : theta1);
                    };
                    horner(coef + ((size_t)t * cap + lo) * (M * D));
#pragma unroll
                    for (int i = 0; i < D; ++i) yo[i] = r[i];
                } else {
                    const double* c = coef + ((size_t)t * cap + lo) * m * dim;
                    for (int i = 0; i < dim; ++i) {
                        double r = __ldg(c + (size_t)(m - 1) * dim + i);
                        for (int k = m - 2; k >= 0; --k)
                            r = __ldg(c + (size_t)k * dim + i) + r * ((k % 2 == 0) ? theta : theta1);
                        yo[i] = r;
                    }
                }
            }
            if (have) found[(size_t)t * nq + qo] = hit ? 1 : 0;
        }
    }
}

typedef void (*com_kernel_t)(long long, int, int, long long, const double*, const double*, const double*, const double*,
                             const unsigned*, long long, const double*, double*, unsigned char*, const unsigned*);
static com_kernel_t com_kernel_for(int m, int dim) {
#define COM_CASE(MM, DD) \
    if (m == MM && dim == DD) return com_evaluate_kernel<MM, DD>;
    COM_CASE(5, 1) COM_CASE(5, 2) COM_CASE(5, 3) COM_CASE(5, 4) COM_CASE(5, 6) COM_CASE(5, 18)
    COM_CASE(8, 1) COM_CASE(8, 2) COM_CASE(8, 3) COM_CASE(8, 6) COM_CASE(8, 18)
#undef COM_CASE
    return com_evaluate_kernel<0, 0>;
}

static int com_validate(ode_b200_ctx* c, int64_t n, int dim, int m, int64_t cap, const void* lower, const void* brk,
                        const void* stp, const void* coef, const void* count, int64_t nq, const void* xq, const void* y,
                        const void* found) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (n < 0 || nq < 0 || dim < 1 || m < 0 || cap < 1 || !lower || !brk || !stp || !coef || !count || !xq || !y || !found)
        return fail(ODE_B200_ERR_INVALID_ARGUMENT, "com_evaluate: bad arguments");
    return 0;
}

static int com_launch(ode_b200_ctx* c, int64_t n, int dim, int m, int64_t cap, const double* lower, const double* brk,
                      const double* stp, const double* coef, const uint32_t* count, int64_t nq, const double* xq, double* y,
                      uint8_t* found, const unsigned* perm, cudaStream_t st) {
    const long long blocks = std::min<long long>((n + kComWarps - 1) / kComWarps, (long long)c->num_sms * kComBlocks);
    com_kernel_for(m, dim)<<<(unsigned)blocks, kComWarps * 32, 0, st>>>(n, dim, m, cap, lower, brk, stp, coef, count, nq, xq, y,
                                                                      found, perm);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    return ODE_B200_SUCCESS;
}

/* device pointers: the model is where ode_b200_integrate_batch_device left it; enqueued on `cuda_stream`, not synchronised */
extern "C" int ode_b200_com_evaluate_device(ode_b200_ctx* c, int64_t n, int dim, int m, int64_t cap, const double* lower,
                                            const double* brk, const double* stp, const double* coef,
                                            const uint32_t* count, int64_t nq, const double* xq, double* y, uint8_t* found,
                                            void* cuda_stream) {
    if (int rc = com_validate(c, n, dim, m, cap, lower, brk, stp, coef, count, nq, xq, y, found)) return rc;
    if (n == 0 || nq == 0) return ODE_B200_SUCCESS;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (st == c->stream) CUDA_TRY(cudaEventRecord(c->ev0, st));
    if (int rc = com_launch(c, n, dim, m, cap, lower, brk, stp, coef, count, nq, xq, y, found, nullptr, st)) return rc;
    if (st == c->stream) {
        CUDA_TRY(cudaEventRecord(c->ev1, st));
        c->timed = true;
    }
    return ODE_B200_SUCCESS;
}

extern "C" int ode_b200_com_evaluate(ode_b200_ctx* c, int64_t n, int dim, int m, int64_t cap,
                                     const double* lower, const double* brk, const double* stp,
                                     const double* coef, const uint32_t* count, int64_t nq, const double* xq,
                                     double* y, uint8_t* found) {
    if (int rc = com_validate(c, n, dim, m, cap, lower, brk, stp, coef, count, nq, xq, y, found)) return rc;
    if (n == 0 || nq == 0) return ODE_B200_SUCCESS;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t nb = (size_t)n * cap * 8, nc = (size_t)n * cap * m * dim * 8;
    if (ensure(c, c->com_lower, n * 8) || ensure(c, c->com_break, nb) || ensure(c, c->com_step, nb) ||
        ensure(c, c->com_coef, nc) || ensure(c, c->com_count, n * 4) || ensure(c, c->scratch0, nq * 8) ||
        ensure(c, c->scratch1, (size_t)n * nq * dim * 8) || ensure(c, c->scratch2, (size_t)n * nq))
        return ODE_B200_ERR_CUDA;
    int rc = ODE_B200_SUCCESS;
    /* Neighbouring lanes should ask for neighbouring intervals (their coefficient loads then share cache lines:
     * 4.8 ms sorted against 18.5 ms in random order for 2^17 models x 2001 queries): unsorted queries are sorted
     * here and the kernel writes each result to its place in the caller's order. */
    std::vector<double> xs;
    std::vector<unsigned> perm;
    bool sorted = true;
    for (int64_t q = 1; q < nq && sorted; ++q) sorted = !(xq[q] < xq[q - 1]);
    if (!sorted && nq <= 0xffffffffll) {
        perm.resize((size_t)nq);
        for (int64_t q = 0; q < nq; ++q) perm[(size_t)q] = (unsigned)q;
        auto key = [&](unsigned i) { const double v = xq[i]; return v != v ? INFINITY : v; }; /* NaN last: a total order */
        std::stable_sort(perm.begin(), perm.end(), [&](unsigned a, unsigned b) { return key(a) < key(b); });
        xs.resize((size_t)nq);
        for (int64_t q = 0; q < nq; ++q) xs[(size_t)q] = xq[perm[(size_t)q]];
        if (ensure(c, c->pipe_dev, std::max<size_t>(c->pipe_dev.cap, (size_t)nq * sizeof(unsigned)))) return ODE_B200_ERR_CUDA;
    }
    const double* xq_src = perm.empty() ? xq : xs.data();
    auto body = [&]() -> int {
        if (!perm.empty())
            CUDA_TRY(cudaMemcpyAsync(c->pipe_dev.p, perm.data(), (size_t)nq * sizeof(unsigned), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c->com_lower.p, lower, n * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c->com_break.p, brk, nb, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c->com_step.p, stp, nb, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c->com_coef.p, coef, nc, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c->com_count.p, count, n * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(c->scratch0.p, xq_src, nq * 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(c->ev0, st));
        if (int r = com_launch(c, n, dim, m, cap, (const double*)c->com_lower.p, (const double*)c->com_break.p,
                               (const double*)c->com_step.p, (const double*)c->com_coef.p, (const uint32_t*)c->com_count.p,
                               nq, (const double*)c->scratch0.p, (double*)c->scratch1.p, (uint8_t*)c->scratch2.p,
                               perm.empty() ? nullptr : (const unsigned*)c->pipe_dev.p, st))
            return r;
        CUDA_TRY(cudaEventRecord(c->ev1, st));
        c->timed = true;
        CUDA_TRY(cudaMemcpyAsync(y, c->scratch1.p, (size_t)n * nq * dim * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(found, c->scratch2.p, (size_t)n * nq, cudaMemcpyDeviceToHost, st));
        return ODE_B200_SUCCESS;
    };
    rc = body();
    const std::string keep = g_last_error;
    const cudaError_t e = cudaStreamSynchronize(st); /* also on an error path: no copy may outlive the call */
    if (rc == ODE_B200_SUCCESS && e != cudaSuccess) return fail(ODE_B200_ERR_CUDA, cudaGetErrorString(e));
    if (rc != ODE_B200_SUCCESS) g_last_error = keep;
    return rc;
}

/* ------------------------------------------------------------------ introspection */
extern "C" double ode_b200_tableau(const char* w, int i, int j) {
    if (!w) return NAN;
    auto in = [](int v, int n) { return v >= 0 && v < n; };
    if (!strcmp(w, "a5")) return in(i, 7) && in(j, 7) ? TabDp5::a(i, j) : NAN;
    if (!strcmp(w, "c5")) return in(i, 7) ? TabDp5::c(i) : NAN;
    if (!strcmp(w, "d5")) return in(i, 7) ? TabDp5::d(i) : NAN;
    if (!strcmp(w, "e5")) return in(i, 7) ? TabDp5::e(i) : NAN;
    if (!strcmp(w, "a8")) return in(i, 16) && in(j, 16) ? TabDp8::a(i, j) : NAN;
    if (!strcmp(w, "b8")) return in(i, 12) ? TabDp8::b(i) : NAN;
    if (!strcmp(w, "bhh8")) return in(i, 3) ? TabDp8::bhh(i) : NAN;
    if (!strcmp(w, "c8")) return in(i, 16) ? TabDp8::c(i) : NAN;
    if (!strcmp(w, "d8")) return in(i, 4) && in(j, 16) ? TabDp8::d(i, j) : NAN;
    if (!strcmp(w, "e8")) return in(i, 16) ? TabDp8::e(i) : NAN;
    return NAN;
}

__global__ void debug_pow_kernel(long long n, const double* x, const double* y, double* o) {
    stage_tables_();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = y ? pow_d(x[i], y[i]) : exp_d(x[i]);
}

__global__ void debug_div_kernel(long long n, const double* a, const double* b, double* o, int shared_den) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (shared_den == 1) { /* one reciprocal for two numerators, as the kernels use it */
        const double y = div_rcp_(b[i]);
        const bool ok = div_in_range_(b[i]);
        const double q1 = div_with_rcp_(a[i], b[i], y, ok), q2 = div_with_rcp_(-a[i], b[i], y, ok);
        o[i] = (q1 == -q2 || (q1 != q1 && q2 != q2)) ? q1 : __longlong_as_double(0x7ff8dead00000000LL);
    } else if (shared_den >= 2) {
        /* the step loop's speculative arithmetic (MathFast, ode_device.cuh): the value when the
         * operands were accepted, a tagged NaN when they were flagged for the exact path */
        MathFast m;
        double q;
        if (shared_den == 2) q = m.div(a[i], b[i]);
        else if (shared_den == 3) q = m.div_nz(a[i], b[i]);
        else if (shared_den == 4) q = m.sqrt(a[i]);
        else if (shared_den == 5) q = m.sqrt_nz(a[i]);
        else { /* one reciprocal, two numerators */
            const double rb = m.rcp(b[i]);
            const double q1 = m.div_by(a[i], b[i], rb), q2 = m.div_by(-a[i], b[i], rb);
            q = (q1 == -q2 || (q1 != q1 && q2 != q2)) ? q1 : __longlong_as_double(0x7ff8dead00000000LL);
            if (q1 == 0.0 && (__double2hiint(q1) ^ __double2hiint(q2)) >= 0) q = __longlong_as_double(0x7ff8dead00000000LL);
        }
        o[i] = m.bad() ? __longlong_as_double(0x7ff8f1a900000000LL) : q;
    } else {
        o[i] = div_(a[i], b[i]);
    }
}

static int debug_fn(ode_b200_ctx* c, int64_t n, const double* x, const double* y, double* o) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (n <= 0) return ODE_B200_SUCCESS;
    CUDA_TRY(cudaSetDevice(c->device));
    if (ensure(c, c->scratch0, n * 8) || ensure(c, c->scratch1, n * 8) || ensure(c, c->scratch2, n * 8))
        return ODE_B200_ERR_CUDA;
    cudaStream_t st = c->stream;
    CUDA_TRY(cudaMemcpyAsync(c->scratch0.p, x, n * 8, cudaMemcpyHostToDevice, st));
    if (y) CUDA_TRY(cudaMemcpyAsync(c->scratch1.p, y, n * 8, cudaMemcpyHostToDevice, st));
    debug_pow_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, (const double*)c->scratch0.p,
                                                                  y ? (const double*)c->scratch1.p : nullptr,
                                                                  (double*)c->scratch2.p);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(o, c->scratch2.p, n * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return ODE_B200_SUCCESS;
}
extern "C" int ode_b200_debug_pow(ode_b200_ctx* c, int64_t n, const double* x, const double* y, double* o) {
    if (!x || !y || !o) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "NULL pointer");
    return debug_fn(c, n, x, y, o);
}
extern "C" int ode_b200_debug_div(ode_b200_ctx* c, int64_t n, const double* a, const double* b, double* o,
                                  int shared_den) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (!a || !b || !o) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "NULL pointer");
    if (n <= 0) return ODE_B200_SUCCESS;
    CUDA_TRY(cudaSetDevice(c->device));
    if (ensure(c, c->scratch0, n * 8) || ensure(c, c->scratch1, n * 8) || ensure(c, c->scratch2, n * 8))
        return ODE_B200_ERR_CUDA;
    cudaStream_t st = c->stream;
    CUDA_TRY(cudaMemcpyAsync(c->scratch0.p, a, n * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c->scratch1.p, b, n * 8, cudaMemcpyHostToDevice, st));
    debug_div_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, (const double*)c->scratch0.p,
                                                                  (const double*)c->scratch1.p,
                                                                  (double*)c->scratch2.p, shared_den);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(o, c->scratch2.p, n * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return ODE_B200_SUCCESS;
}
/* the f32 powf clone (ode_powf.h) on the device: tables in shared memory like the f64 ones */
static __device__ const unsigned long long g_ode_powf_log2_tab[32] = ODE_POWF_LOG2_TAB_INIT;
static __device__ const unsigned long long g_ode_exp2f_tab[32] = ODE_EXP2F_TAB_INIT;
__global__ void debug_powf_kernel(long long n, const float* x, const float* y, float* o) {
    __shared__ unsigned long long tabs[64];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) tabs[i] = i < 32 ? g_ode_powf_log2_tab[i] : g_ode_exp2f_tab[i - 32];
    __syncthreads();
    ode_powf_tabs T;
    T.lt = tabs;
    T.et = tabs + 32;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = ode_powf(x[i], y[i], T);
}
extern "C" int ode_b200_debug_powf(ode_b200_ctx* c, int64_t n, const float* x, const float* y, float* o) {
    if (!c) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (!x || !y || !o) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "NULL pointer");
    if (n <= 0) return ODE_B200_SUCCESS;
    CUDA_TRY(cudaSetDevice(c->device));
    if (ensure(c, c->scratch0, n * 4) || ensure(c, c->scratch1, n * 4) || ensure(c, c->scratch2, n * 4))
        return ODE_B200_ERR_CUDA;
    cudaStream_t st = c->stream;
    CUDA_TRY(cudaMemcpyAsync(c->scratch0.p, x, n * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(c->scratch1.p, y, n * 4, cudaMemcpyHostToDevice, st));
    debug_powf_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, (const float*)c->scratch0.p,
                                                                   (const float*)c->scratch1.p, (float*)c->scratch2.p);
    c->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(o, c->scratch2.p, n * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return ODE_B200_SUCCESS;
}
extern "C" int ode_b200_debug_exp(ode_b200_ctx* c, int64_t n, const double* x, double* o) {
    if (!x || !o) return fail(ODE_B200_ERR_INVALID_ARGUMENT, "NULL pointer");
    return debug_fn(c, n, x, nullptr, o);
}
