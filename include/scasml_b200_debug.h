/* scasml_b200_debug.h -- test hooks and micro-benchmarks.  NOT part of the product library.
 *
 * These entry points exist only in libscasml_b200_dbg.so, which scasml_gp_b200/build.py links from the same sources with
 * -DSCASML_DEBUG_HOOKS (the tcgen05 evaluation kernel then carries its clock64 stamps and experiment flags; the product
 * build compiles them out).  tests/ and tools/ load it through scasml_gp_b200/_lib.py::load_debug(); nothing under
 * scasml_gp_b200/{equations,models,solvers} does.
 */
#ifndef SCASML_B200_DEBUG_H
#define SCASML_B200_DEBUG_H

#include "scasml_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* sampler: normals / uniforms for flat indices start..start+count-1 of stream (k0, domain, seed)
 * (replaces jax.random.normal / uniform of solvers/ScaSML.py:190,229, solvers/ScaSML_full_history.py:110,144) */
SCASML_API int scasml_debug_draw(unsigned stream_id, unsigned domain, unsigned seed, long long start, long long count,
                      int uniform, double* out_dev, void* stream);

/* dense FP64 pieces of the fit (models/GP.py:260-268 factor, :533 Newton solve) */
SCASML_API int scasml_debug_spd_inverse(double* A_dev /*in: SPD, out: L*/, long long n, double* P_dev, void* ws_dev,
                             size_t ws_bytes, void* stream); /* ws >= (n*n + 64*64*ceil(n/64) + 64*n)*8 */
SCASML_API int scasml_debug_lu_solve(double* A_dev, long long n, double* rhs_dev, void* stream);
/* tcgen05 plumbing self-test: D[128][N] (f32) = A[128][K] (f16) x B[N][K]^T (f16) through the same shared-memory
 * layout, descriptors, tcgen05.mma and tcgen05.ld helpers as the evaluation kernel (descriptor fields at run time) */
SCASML_API int scasml_debug_tc_gemm(const void* A_half_dev, const void* B_half_dev, float* D_dev, int K, int N, unsigned lbo16,
                         unsigned sbo16, unsigned layout, unsigned kstep_bytes, void* stream);
/* SM-clock timeline of CTA `block & 0xFFFFFF` of one tcgen05 evaluation launch, experiment flags in `block >> 24`
 * (stamps_dev: 1024 int64; scratch_dev: 4 R doubles) */
SCASML_API int scasml_debug_tc_timeline(const scasml_gp* gp, const double* X_dev, long long R, int mode, int block,
                             long long* stamps_dev, double* scratch_dev, void* stream);
/* micro-benchmark: cycles per tcgen05.mma (M=128, K=16, f16) for N, `nchains` independent accumulators, A from smem (0) / TMEM (1);
 * cycles_dev[0] = issue span, cycles_dev[1] = span until the commit arrives */
SCASML_API int scasml_debug_tc_mma_bench(int N, int nchains, int ts_mode, int iters, long long* cycles_dev, void* stream);
/* micro-benchmark of the epilogue pipes (csrc/tc_bench.cu: TMEM load/store, MUFU, split chunk, MMA interference);
 * out_dev: 8 int64: [0] epilogue-warp cycles, [1] MMA issue cycles, [2] MMA cycles until commit, [3] MMA count */
SCASML_API int scasml_debug_tc_pipe_bench(int mode, int N, int iters, long long* out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCASML_B200_DEBUG_H */
