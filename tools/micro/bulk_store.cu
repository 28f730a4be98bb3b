// Micro-benchmark: trajectory-major result stores from per-lane shared-memory runs on B200.
// Each lane owns a run of R samples (x[R], y[R][3]) and a private output slice (like one trajectory's
// out_x / out_y); per flush it writes 8R bytes to out_x and 24R bytes to out_y.
//   mode 0: per-lane TMA bulk copies (cp.async.bulk.global.shared::cta), 2 per lane per flush
//   mode 1: warp-cooperative copy of one lane's run at a time (the current flush_staged_)
//   mode 2: per-lane 16-byte vector stores
//   mode 3: per-lane TMA bulk copies of half runs (R/2 samples: 32 B + 96 B), the other half keeps
//           filling meanwhile (wait_group.read 1) -- the layout the step-loop kernels use
//   mode 4: ONE TMA tensor store per warp and half run (lanes in lockstep on consecutive trajectories, the Rk4
//           case): box = 32 trajectories x (R/2 samples x 3 doubles), rows a trajectory pitch apart
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_store bulk_store.cu -lcuda
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
constexpr int R = 8, N = 3, S = R * (1 + N) + 2;  // lane stride in doubles (16-byte aligned)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(128) k(double* out_x, double* out_y, long long cap, int flushes) {
    extern __shared__ __align__(128) double buf[];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double* mine = buf + threadIdx.x * S;
    const unsigned lane = threadIdx.x & 31;
    for (int f = 0; f < flushes; ++f) {
        if (MODE == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        for (int j = 0; j < R; ++j) {
            mine[j] = f + j;
            for (int i = 0; i < N; ++i) mine[R + j * N + i] = t + i;
        }
        const long long s0 = t * cap + (long long)f * R;
        if (MODE == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out_x + s0), "r"(smem_u32(mine)), "r"(R * 8) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out_y + s0 * N), "r"(smem_u32(mine + R)), "r"(R * N * 8) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        } else if (MODE == 1) {
            __syncwarp();
            for (int l = 0; l < 32; ++l) {
                const long long tl = __shfl_sync(0xffffffffu, t, l);
                const double* src = buf + ((threadIdx.x & ~31u) + l) * S;
                const long long b0 = tl * cap + (long long)f * R;
                for (unsigned j = lane; j < R * (1 + N); j += 32) {
                    if (j < R) out_x[b0 + j] = src[j];
                    else out_y[b0 * N + (j - R)] = src[R + (j - R)];
                }
            }
            __syncwarp();
        } else {
            const double2* s2 = reinterpret_cast<const double2*>(mine);
            double2* dx = reinterpret_cast<double2*>(out_x + s0);
            double2* dy = reinterpret_cast<double2*>(out_y + s0 * N);
#pragma unroll
            for (int j = 0; j < R / 2; ++j) dx[j] = s2[j];
#pragma unroll
            for (int j = 0; j < R * N / 2; ++j) dy[j] = s2[R / 2 + j];
        }
    }
    if (MODE == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// half-run groups: slot = sample % R, group g = (sample / (R/2)) % 2; x-run [R] then y-run [R][N]
__global__ void __launch_bounds__(128) k3(double* out_x, double* out_y, long long cap, int flushes) {
    extern __shared__ __align__(128) double buf[];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double* mine = buf + threadIdx.x * S;
    constexpr int G = R / 2;
    for (int f = 0; f < 2 * flushes; ++f) {  // one half run per iteration
        const int g = f & 1;
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        for (int j = 0; j < G; ++j) {
            mine[g * G + j] = f / 2 + g * G + j;
            for (int i = 0; i < N; ++i) mine[R + (g * G + j) * N + i] = t + i;
        }
        const long long s0 = t * cap + (long long)f * G;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out_x + s0), "r"(smem_u32(mine + g * G)), "r"(G * 8) : "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out_y + s0 * N), "r"(smem_u32(mine + R + g * G * N)), "r"(G * N * 8) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// warp-level tensor stores: per warp two group buffers of [32][G*N] (y) and [32][G] (x) doubles, dense
__global__ void __launch_bounds__(128) k4(const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_x,
                                          int flushes) {
    extern __shared__ __align__(128) double buf[];
    constexpr int G = R / 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double* wy = buf + warp * (2 * 32 * G * N + 2 * 32 * G);  // [2][32][G*N] then [2][32][G]
    double* wx = wy + 2 * 32 * G * N;
    for (int f = 0; f < 2 * flushes; ++f) {
        const int g = f & 1;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        double* my = wy + (g * 32 + lane) * G * N;
        double* mx = wx + (g * 32 + lane) * G;
        for (int j = 0; j < G; ++j) {
            mx[j] = f / 2 + g * G + j;
            for (int i = 0; i < N; ++i) my[j * N + i] = t + i;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            const int traj0 = (int)(t);  // lane 0's trajectory = the warp's first
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&map_y), "r"(0),
                         "r"(f), "r"(traj0), "r"(smem_u32(wy + g * 32 * G * N))
                         : "memory");
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(&map_x), "r"(0),
                         "r"(f), "r"(traj0), "r"(smem_u32(wx + g * 32 * G))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static CUtensorMap make_map(double* base, long long cap, long long nt, int inner) {  // [nt][cap / G][inner] doubles
    CUtensorMap m;
    constexpr int G = R / 2;
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)(cap / G), (cuuint64_t)nt};
    cuuint64_t strides[2] = {(cuuint64_t)inner * 8, (cuuint64_t)(cap / G) * inner * 8};
    cuuint32_t box[3] = {(cuuint32_t)inner, 1, 32};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, es,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return m;
}

template <int MODE>
void run(const char* name, int ctas_per_sm) {
    const int flushes = 250;
    const long long cap = (long long)flushes * R;
    const int grid = 148 * ctas_per_sm, block = 128;
    const long long nt = (long long)grid * block;
    double *ox, *oy;
    cudaMalloc(&ox, nt * cap * 8);
    cudaMalloc(&oy, nt * cap * 8 * N);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const size_t smem = MODE == 4 ? (size_t)(block / 32) * (2 * 32 * (R / 2) * (N + 1)) * 8 : block * S * 8;
    CUtensorMap my, mx;
    if (MODE == 4) {
        my = make_map(oy, cap, nt, (R / 2) * N);
        mx = make_map(ox, cap, nt, R / 2);
    }
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (MODE == 4) k4<<<grid, block, smem>>>(my, mx, flushes);
        else if (MODE == 3) k3<<<grid, block, smem>>>(ox, oy, cap, flushes);
        else k<(MODE == 3 || MODE == 4) ? 0 : MODE><<<grid, block, smem>>>(ox, oy, cap, flushes);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double bytes = (double)nt * cap * 8 * (1 + N);
    printf("%-28s CTAs/SM=%d: %.3f ms, %.1f GB/s (%.2f GB) %s\n", name, ctas_per_sm, best, bytes / best / 1e6, bytes / 1e9,
           cudaGetErrorString(cudaGetLastError()));
    // spot check
    double h[4];
    cudaMemcpy(h, oy + (nt - 1) * cap * N + (cap - 1) * N, 24, cudaMemcpyDeviceToHost);
    cudaMemcpy(h + 3, ox + (nt - 1) * cap + cap - 1, 8, cudaMemcpyDeviceToHost);
    if (h[0] != nt - 1 || h[2] != nt + 1 || h[3] != flushes - 1 + R - 1) printf("   WRONG DATA %g %g %g %g\n", h[0], h[1], h[2], h[3]);
    cudaFree(ox);
    cudaFree(oy);
}

int main() {
    for (int c : {2, 4, 8}) {
        run<0>("TMA bulk per lane", c);
        run<1>("warp-cooperative", c);
        run<2>("per-lane 16-byte stores", c);
        run<3>("TMA bulk, half runs", c);
        run<4>("TMA tensor store per warp", c);
    }
    return 0;
}
