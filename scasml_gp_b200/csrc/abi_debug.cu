// Test hooks and micro-benchmark entry points (include/scasml_b200_debug.h).  Linked into libscasml_b200_dbg.so only
// (-DSCASML_DEBUG_HOOKS); the product library does not contain this file.
#include "../../include/scasml_b200_debug.h"
#include "abi_handle.cuh"
#include "gp_fit.cuh"

#ifndef SCASML_DEBUG_HOOKS
#error "abi_debug.cu belongs to the debug build (-DSCASML_DEBUG_HOOKS)"
#endif

using namespace scasml;

namespace {

__global__ void debug_draw_kernel(PhiloxKey key, unsigned long long start, long long count, int uniform,
                                  const __half* __restrict__ tab, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t c = chunk16(start + (unsigned long long)i, key);
    out[i] = uniform ? chunk_to_uniform(c) : chunk_to_normal(tab, c);
}

}  // namespace

extern "C" {

int scasml_debug_draw(unsigned stream_id, unsigned domain, unsigned seed, long long start, long long count,
                      int uniform, double* out_dev, void* stream) {
    const __half* tab = normal_table_for_current_device();
    SC_REQUIRE(tab != nullptr, "normal table not set");
    if (count <= 0) return OK;
    debug_draw_kernel<<<(unsigned)cdiv(count, 256), 256, 0, (cudaStream_t)stream>>>(
        make_key(stream_id, domain, seed), (unsigned long long)start, count, uniform, tab, out_dev);
    SC_LAUNCH_CHECK();
    return OK;
}

int scasml_debug_spd_inverse(double* A_dev, long long n, double* P_dev, void* ws_dev, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const long nblk = cdiv(n, 64);
    const size_t need = ((size_t)n * n + (size_t)nblk * 64 * 64 + 64 * (size_t)n) * 8 + 1024;
    SC_REQUIRE(ws_bytes >= need, "debug_spd_inverse: workspace too small");
    double* X = (double*)ws_dev;
    double* invd = X + (size_t)n * n;
    double* tmp = invd + (size_t)nblk * 64 * 64;
    int* d_fail = (int*)(tmp + 64 * (size_t)n);
    SC_CUDA(cudaMemsetAsync(d_fail, 0, sizeof(int), st));
    int rc = cholesky_lower(A_dev, (long)n, invd, d_fail, st);
    if (rc != OK) return rc;
    rc = tri_inverse_lower(A_dev, (long)n, invd, X, P_dev, st);     // P_dev is written afterwards: scratch until then
    if (rc != OK) return rc;
    rc = dgemm((int)n, (int)n, (int)n, 1.0, X, 1, n, X, n, 1, 0.0, P_dev, n, 0, st);
    if (rc != OK) return rc;
    int h_fail = 0;
    SC_CUDA(cudaMemcpyAsync(&h_fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, st));
    SC_CUDA(cudaStreamSynchronize(st));
    if (h_fail) { set_error("matrix is not positive definite"); return ERR_NUMERIC; }
    return OK;
}

int scasml_debug_lu_solve(double* A_dev, long long n, double* rhs_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int* d_fail = nullptr;
    SC_CUDA(cudaMalloc(&d_fail, sizeof(int)));
    SC_CUDA(cudaMemsetAsync(d_fail, 0, sizeof(int), st));
    int rc = lu_solve_inplace(A_dev, (long)n, rhs_dev, d_fail, st);
    int h_fail = 0;
    if (rc == OK) {
        cudaMemcpyAsync(&h_fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
    }
    cudaFree(d_fail);
    if (rc != OK) return rc;
    if (h_fail) { set_error("singular matrix in LU"); return ERR_NUMERIC; }
    return OK;
}

int scasml_debug_tc_gemm(const void* A_half_dev, const void* B_half_dev, float* D_dev, int K, int N, unsigned lbo16,
                         unsigned sbo16, unsigned layout, unsigned kstep_bytes, void* stream) {
    SC_REQUIRE(A_half_dev && B_half_dev && D_dev, "debug_tc_gemm: null");
    return tc_selftest(A_half_dev, B_half_dev, D_dev, K, N, lbo16, sbo16, layout, kstep_bytes, (cudaStream_t)stream);
}

int scasml_debug_tc_timeline(const scasml_gp* g, const double* X_dev, long long R, int mode, int block,
                             long long* stamps_dev, double* scratch_dev, void* stream) {
    SC_REQUIRE(g && g->has_alpha && g->tc.images, "debug_tc_timeline: tcgen05 route unavailable");
    return tc_timeline(g->v, g->tc, X_dev, (long)R, mode, block, stamps_dev, scratch_dev, (cudaStream_t)stream);
}

int scasml_debug_tc_mma_bench(int N, int nchains, int ts_mode, int iters, long long* cycles_dev, void* stream) {
    return tc_mma_bench(N, nchains, ts_mode, iters, cycles_dev, (cudaStream_t)stream);
}

int scasml_debug_tc_pipe_bench(int mode, int N, int iters, long long* out_dev, void* stream) {
    return tc_pipe_bench(mode, N, iters, out_dev, (cudaStream_t)stream);
}

}  // extern "C"
