/*
 * ode_jit.cpp -- NVRTC path for user-registered systems (see ode_jit.h).
 *
 * The reference lets users implement `trait System` in Rust (src/dop_shared.rs:31-42) and
 * monomorphises the solver around it. The device equivalent of that monomorphisation is to
 * compile the kernel template around the user's rhs/solout at run time: the generated
 * translation unit defines `struct SysUser` from the two function bodies and explicitly
 * instantiates ode_step_loop_kernel<Stepper<SysUser, OUT>> for the requested method and
 * output type only.
 */
#include "ode_jit.h"

#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <vector>

#include "ode_b200.h"
#include "ode_embedded_headers.inc"

namespace ode_b200 {

namespace {

struct Api {
    bool ok = false;     /* NVRTC usable (compiling needs no GPU) */
    bool drv_ok = false; /* driver API usable (loading / launching) */
    std::string why;
    /* nvrtc */
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*,
                                 const char* const*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
    nvrtcResult (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
    /* driver */
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             CUstream, void**, void**) = nullptr;
    CUresult (*Occupancy)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*GetErrorName)(CUresult, const char**) = nullptr;
};

void* open_first(const std::vector<const char*>& names) {
    for (const char* n : names)
        if (void* h = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) return h;
    return nullptr;
}

Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, []() {
        void* rtc = open_first({"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"});
        if (!rtc) {
            a.why = "cannot load libnvrtc.so.12";
            return;
        }
#define L(h, field, sym)                                         \
    a.field = (decltype(a.field))dlsym(h, sym);                  \
    if (!a.field) {                                              \
        a.why = std::string("missing symbol ") + sym;            \
        return;                                                  \
    }
        L(rtc, CreateProgram, "nvrtcCreateProgram")
        L(rtc, DestroyProgram, "nvrtcDestroyProgram")
        L(rtc, CompileProgram, "nvrtcCompileProgram")
        L(rtc, GetProgramLogSize, "nvrtcGetProgramLogSize")
        L(rtc, GetProgramLog, "nvrtcGetProgramLog")
        L(rtc, GetCUBINSize, "nvrtcGetCUBINSize")
        L(rtc, GetCUBIN, "nvrtcGetCUBIN")
        L(rtc, AddNameExpression, "nvrtcAddNameExpression")
        L(rtc, GetLoweredName, "nvrtcGetLoweredName")
        L(rtc, GetErrorString, "nvrtcGetErrorString")
        a.ok = true;
        void* drv = open_first({"libcuda.so.1", "libcuda.so"});
        if (!drv) {
            a.why = "cannot load libcuda.so.1 (no NVIDIA driver)";
            return;
        }
        L(drv, ModuleLoadData, "cuModuleLoadData")
        L(drv, ModuleGetFunction, "cuModuleGetFunction")
        L(drv, LaunchKernel, "cuLaunchKernel")
        L(drv, Occupancy, "cuOccupancyMaxActiveBlocksPerMultiprocessor")
        L(drv, FuncSetAttribute, "cuFuncSetAttribute")
        L(drv, GetErrorName, "cuGetErrorName")
#undef L
        a.drv_ok = true;
    });
    return a;
}

std::mutex g_mu;
std::vector<std::unique_ptr<JitSystem>> g_systems;

struct Compiled {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    int per_sm = 0;
};
/* (system, device, method, out_type) -> loaded kernel */
std::map<std::tuple<int, int, int, int, int>, Compiled> g_cache;

std::string stepper_expr(int method, int out_type, bool staged) {
    const std::string st = staged ? ", true>" : ", false>";
    if (method == ODE_B200_RK4) return "ode_b200::Rk4Stepper<ode_b200::SysUser" + st;
    if (method == ODE_B200_DOPRI5)
        return "ode_b200::Dopri5Stepper<ode_b200::SysUser, " + std::to_string(out_type) + st;
    /* Dop853 treats Continuous like Sparse (dop853.rs:398,540-542) */
    const int o = out_type == ODE_B200_OUT_DENSE || out_type == ODE_B200_OUT_CONTINUOUS8 ? out_type : ODE_B200_OUT_SPARSE;
    return "ode_b200::Dop853Stepper<ode_b200::SysUser, " + std::to_string(o) + st;
}

std::string make_source(const JitSystem& s, const std::string& stepper) {
    std::string src = "#include \"ode_kernels.cuh\"\nnamespace ode_b200 {\nstruct SysUser {\n";
    src += "  static constexpr int DIM = " + std::to_string(s.dim) + ", NP = " + std::to_string(s.n_params) + ";\n";
    src += std::string("  static constexpr bool HAS_SOLOUT = ") + (s.has_solout ? "true" : "false") + ";\n";
    /* div_d / sqrt_d: division and square root through the kernel's math policy (ode_device.cuh,
     * MathFast): same IEEE results as `/` and sqrt(), but branch-free inside the step loop */
    const bool spec = s.rhs_body.find("div_d(") != std::string::npos || s.rhs_body.find("sqrt_d(") != std::string::npos;
    src += std::string("  static constexpr bool SPEC = ") + (spec ? "true" : "false") + ";\n";
    src += "#define div_d(a, b) ode_m_.div((a), (b))\n#define sqrt_d(v) ode_m_.sqrt(v)\n";
    src += "  template <class M>\n  ODE_DEV static void rhs(double x, const double* y, double* dy, const double* p, M& ode_m_) {\n" +
           s.rhs_body + "\n  }\n";
    src += "#undef div_d\n#undef sqrt_d\n";
    src += "  ODE_DEV static bool solout(double x, const double* y, const double* dy, const double* p) {\n" +
           (s.has_solout ? s.solout_body : std::string("return false;")) + "\n  }\n};\n}\n";
    src += "template __global__ void ode_b200::ode_step_loop_kernel<" + stepper +
           " >(const __grid_constant__ ode_b200::KernelArgs);\n";
    return src;
}

struct Cubin {
    std::vector<char> bin;
    std::string mangled;
};
/* (system, method, out_type) -> cubin; device independent */
std::map<std::tuple<int, int, int, int>, Cubin> g_cubins;

bool compile_cubin(const JitSystem& s, int method, int out_type, bool staged, Cubin* out, std::string* err) {
    Api& a = api();
    const std::string stepper = stepper_expr(method, out_type, staged);
    const std::string src = make_source(s, stepper);
    const std::string name_expr = "ode_b200::ode_step_loop_kernel<" + stepper + " >";
    nvrtcProgram prog;
    nvrtcResult r = a.CreateProgram(&prog, src.c_str(), (s.name + ".cu").c_str(), kNumEmbedded, kEmbeddedSources,
                                    kEmbeddedNames);
    if (r != NVRTC_SUCCESS) {
        *err = std::string("nvrtcCreateProgram: ") + a.GetErrorString(r);
        return false;
    }
    a.AddNameExpression(prog, name_expr.c_str());
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo",
                          "-default-device"};
    r = a.CompileProgram(prog, 5, opts);
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        a.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) a.GetProgramLog(prog, &log[0]);
        *err = std::string("NVRTC compile of system '") + s.name + "' failed: " + a.GetErrorString(r) + "\n" + log;
        a.DestroyProgram(&prog);
        return false;
    }
    const char* lowered = nullptr;
    r = a.GetLoweredName(prog, name_expr.c_str(), &lowered);
    if (r != NVRTC_SUCCESS || !lowered) {
        *err = "nvrtcGetLoweredName failed";
        a.DestroyProgram(&prog);
        return false;
    }
    out->mangled = lowered;
    size_t nbin = 0;
    a.GetCUBINSize(prog, &nbin);
    out->bin.resize(nbin);
    r = a.GetCUBIN(prog, out->bin.data());
    a.DestroyProgram(&prog);
    if (r != NVRTC_SUCCESS || nbin == 0) {
        *err = "nvrtcGetCUBIN failed";
        return false;
    }
    return true;
}

/* g_mu held */
const Cubin* get_cubin(int idx, const JitSystem& s, int method, int out_type, bool staged, std::string* err) {
    const int o = method == ODE_B200_RK4 ? 0
                  : (method == ODE_B200_DOP853 && out_type == ODE_B200_OUT_CONTINUOUS ? ODE_B200_OUT_SPARSE : out_type);
    auto key = std::make_tuple(idx, method, o, staged ? 1 : 0);
    auto it = g_cubins.find(key);
    if (it != g_cubins.end()) return &it->second;
    Cubin c;
    if (!compile_cubin(s, method, out_type, staged, &c, err)) return nullptr;
    return &(g_cubins[key] = std::move(c));
}

bool load(const Cubin& cb, Compiled* out, std::string* err) {
    Api& a = api();
    if (!a.drv_ok) {
        *err = "JIT launch unavailable: " + a.why;
        return false;
    }
    CUresult cr = a.ModuleLoadData(&out->mod, cb.bin.data());
    if (cr == CUDA_SUCCESS) cr = a.ModuleGetFunction(&out->fn, out->mod, cb.mangled.c_str());
    if (cr != CUDA_SUCCESS) {
        const char* nm = "?";
        a.GetErrorName(cr, &nm);
        *err = std::string("loading the JIT-compiled kernel failed: ") + nm;
        return false;
    }
    return true;
}

}  // namespace

int jit_register(const char* name, int dim, int n_params, const char* rhs_body, const char* solout_body,
                 std::string* err) {
    (void)err;
    std::lock_guard<std::mutex> g(g_mu);
    std::unique_ptr<JitSystem> s(new JitSystem());
    s->name = name;
    s->dim = dim;
    s->n_params = n_params;
    s->rhs_body = rhs_body;
    s->has_solout = solout_body != nullptr && solout_body[0] != '\0';
    if (s->has_solout) s->solout_body = solout_body;
    g_systems.push_back(std::move(s));
    return (int)g_systems.size() - 1;
}

bool jit_compile(int idx, int method, int out_type, bool staged, std::string* err) {
    Api& a = api();
    if (!a.ok) {
        *err = "JIT unavailable: " + a.why;
        return false;
    }
    const JitSystem* s = jit_system(idx);
    if (!s) {
        *err = "bad JIT system index";
        return false;
    }
    std::lock_guard<std::mutex> g(g_mu);
    return get_cubin(idx, *s, method, out_type, staged, err) != nullptr;
}

const JitSystem* jit_system(int idx) {
    std::lock_guard<std::mutex> g(g_mu);
    if (idx < 0 || idx >= (int)g_systems.size()) return nullptr;
    return g_systems[idx].get();
}

bool jit_launch(int idx, int device, int num_sms, int method, int out_type, bool staged, const void* kernel_args,
                size_t kernel_args_bytes, int block, long long blocks_all, int use_queue, size_t dyn_smem, void* stream,
                std::string* err) {
    (void)kernel_args_bytes;
    Api& a = api();
    if (!a.ok) {
        *err = "JIT unavailable: " + a.why;
        return false;
    }
    const JitSystem* s = jit_system(idx);
    if (!s) {
        *err = "bad JIT system index";
        return false;
    }
    Compiled c;
    {
        std::lock_guard<std::mutex> g(g_mu);
        const int o = method == ODE_B200_RK4 ? 0 : out_type;
        auto key = std::make_tuple(idx, device, method, o, staged ? 1 : 0);
        auto it = g_cache.find(key);
        if (it == g_cache.end()) {
            const Cubin* cb = get_cubin(idx, *s, method, out_type, staged, err);
            if (!cb || !load(*cb, &c, err)) return false;
            g_cache[key] = c;
        } else {
            c = it->second;
        }
    }
    if (dyn_smem > 32 * 1024) a.FuncSetAttribute(c.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dyn_smem);
    int per_sm = 0;
    if (a.Occupancy(&per_sm, c.fn, block, dyn_smem) != CUDA_SUCCESS || per_sm < 1) {
        *err = "JIT kernel does not fit on an SM";
        return false;
    }
    long long grid = blocks_all;
    if (use_queue) grid = std::min<long long>(blocks_all, (long long)num_sms * per_sm);
    void* args[] = {const_cast<void*>(kernel_args)};
    CUresult cr = a.LaunchKernel(c.fn, (unsigned)grid, 1, 1, (unsigned)block, 1, 1, (unsigned)dyn_smem,
                                 (CUstream)stream, args, nullptr);
    if (cr != CUDA_SUCCESS) {
        const char* nm = "?";
        a.GetErrorName(cr, &nm);
        *err = std::string("cuLaunchKernel failed: ") + nm;
        return false;
    }
    return true;
}

}  // namespace ode_b200
