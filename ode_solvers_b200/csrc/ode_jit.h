/*
 * ode_jit.h -- user systems registered from CUDA source (ode_b200_system_from_source).
 * The same kernel templates as the built-ins (ode_kernels.cuh, embedded in the library as
 * text) are compiled with NVRTC for sm_100a with --fmad=false around the user's rhs/solout
 * bodies, so a registered system gets full inlining and register allocation -- no device
 * function pointers. libnvrtc and libcuda are dlopen()ed on first use, so the library loads
 * (and exports every ABI symbol) on a machine without a GPU driver.
 */
#pragma once
#include <string>

namespace ode_b200 {

struct JitSystem {
    std::string name;
    int dim = 0, n_params = 0;
    bool has_solout = false;
    std::string rhs_body, solout_body;
};

/* returns the index (>= 0) of the registered system or -1 (message in *err) */
int jit_register(const char* name, int dim, int n_params, const char* rhs_body, const char* solout_body,
                 std::string* err);
const JitSystem* jit_system(int idx);

/* NVRTC-compile one (method, out_type) kernel of a registered system to a cubin (no GPU needed) */
bool jit_compile(int idx, int method, int out_type, bool staged, std::string* err);

/* compile (once per system/device/method/out_type), size the grid and launch on `stream` */
bool jit_launch(int idx, int device, int num_sms, int method, int out_type, bool staged, const void* kernel_args,
                size_t kernel_args_bytes, int block, long long blocks_all, int use_queue, size_t dyn_smem, void* stream,
                std::string* err);

}  // namespace ode_b200
