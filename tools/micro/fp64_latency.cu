// FP64 pipe micro-benchmark for B200: dependent-issue latency and throughput of DFMA / DADD / DMUL
// as a function of ILP (independent chains per warp) and warps per SM sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = a + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == 0) v[i] = fma(v[i], b, a);
                if (OP == 1) v[i] = __dadd_rn(v[i], b);
                if (OP == 2) v[i] = __dmul_rn(v[i], b);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP, int OP>
void run(const char* name, int warps) {
    double* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8);
    cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    chain<ILP, OP><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 0.9999999);
    chain<ILP, OP><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 0.9999999);
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = (double)h[0] / (iters * 16.0);
    printf("%s ILP=%d warps/SM=%2d: %.2f cycles per round of %d instr/warp -> %.2f cyc/instr/warp, SM thread-instr/clk %.1f\n", name, ILP,
           warps, c, ILP, c / ILP, warps * 32.0 * ILP / c);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    const char* names[3] = {"DFMA", "DADD", "DMUL"};
    for (int w : {1, 4, 8, 16}) {
        run<1, 0>(names[0], w);
        run<2, 0>(names[0], w);
        run<4, 0>(names[0], w);
        run<8, 0>(names[0], w);
        run<1, 1>(names[1], w);
        run<4, 1>(names[1], w);
        run<1, 2>(names[2], w);
        run<4, 2>(names[2], w);
    }
    return 0;
}
