// FP64 SIMT pipe on this GPU: throughput (independent chains, many warps) and latency (one dependent chain) of DFMA / DADD / DMUL /
// F2F.F32.F64 / F2F.F64.F32, in SM cycles per warp instruction.  nvcc -arch=sm_100a -O3 -o fp64_pipe fp64_pipe.cu && ./fp64_pipe
#include <cstdio>
#include <cuda_runtime.h>

template <int OP, int CHAINS>
__global__ void k(double* out, long long* cyc, double seed, int iters) {
    double a[CHAINS];
    float f[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { a[i] = seed + i + threadIdx.x; f[i] = (float)(seed + i); }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) a[i] = fma(a[i], 1.0000001, 0.5);
            if (OP == 1) a[i] = a[i] + 0.5;
            if (OP == 2) a[i] = a[i] * 1.0000001;
            if (OP == 3) { f[i] = (float)a[i]; a[i] = __hiloint2double(__double2hiint(a[i]), __float_as_int(f[i])); }   // F2F.F32.F64 + cheap dependency
            if (OP == 4) { a[i] = (double)f[i]; f[i] = __int_as_float(__double2loint(a[i]) ^ 1); }                          // F2F.F64.F32
            if (OP == 5) f[i] = fmaf(f[i], 1.0000001f, 0.5f);                                                                // FP32 reference
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP, int CHAINS>
void run(const char* name, int warps) {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    k<OP, CHAINS><<<148, warps * 32>>>(out, cyc, 1.0, iters);
    k<OP, CHAINS><<<148, warps * 32>>>(out, cyc, 1.0, iters);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / iters / CHAINS;
    printf("%-14s chains %d warps/SM %2d: %7.2f cycles per instruction per warp -> %6.2f warp-instr/clk/SM (%5.1f lanes/clk/SM)\n", name, CHAINS, warps,
           per, warps / per, 32.0 * warps / per);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0, 1>("DFMA latency", 1); run<0, 8>("DFMA", 1); run<0, 8>("DFMA", 4); run<0, 8>("DFMA", 16);
    run<1, 8>("DADD", 16); run<2, 8>("DMUL", 16);
    run<3, 1>("F2F.F32.F64 lat", 1); run<3, 8>("F2F.F32.F64", 16);
    run<4, 1>("F2F.F64.F32 lat", 1); run<4, 8>("F2F.F64.F32", 16);
    run<5, 1>("FFMA latency", 1); run<5, 8>("FFMA", 16);
    return 0;
}
