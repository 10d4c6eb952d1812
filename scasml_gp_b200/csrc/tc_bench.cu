// Pipe micro-benchmarks behind the tile sizing of the tcgen05 evaluation kernel (debug hooks, need a GPU).
// One CTA per SM, 16 "epilogue" warps + 1 MMA warp, SM-clock cycles of warp 0 / of the MMA warp:
//   mode 0  tcgen05.ld 32x32b.x16 stream (4 loads per wait)            -> TMEM read bytes / clk / SM
//   mode 1  tcgen05.st 32x32b.x8 stream                                -> TMEM write bytes / clk / SM
//   mode 2  ex2.approx stream (16 independent per iteration)           -> MUFU rate
//   mode 3  the evaluation epilogue chunk: ld x16 -> 16 ex2 -> f16 hi/lo split -> 2 x st x8 (in place)
//   mode 4  mode 3 with the MMA warp issuing SS-mode N = `N` MMAs concurrently (interference)
//   mode 5  MMA warp alone: SS-mode N = `N` MMAs
//   mode 6  MMA warp alone: TS-mode N = `N` MMAs (A operand in tensor memory)
//   mode 7  mode 6 with a tcgen05.commit (to a barrier nobody waits on) after every 7 MMAs
//   mode 8  mode 6 with commit + tcgen05.fence + an mbarrier poll on an already-completed barrier after every 7 MMAs
//   mode 9  mode 3 (epilogue chunk loop) with the MMA warp issuing TS-mode MMAs concurrently
//   mode 10 MMA warp alone: TS-mode, one N = 128 MMA followed by two N = 16 MMAs into another accumulator (shape switching)
//   mode 11 like 10, batched: 14 x N = 128 then 24 x N = 16
//   mode 12 / 13  the 16 warps run a DFMA loop (8 chains) with / without the MMA warp issuing TS-mode N = `N` MMAs concurrently
//   mode 14 / 15  the same with FFMA     (is the FP64 SIMT pipe slowed down while the tensor pipe is busy?)
#include "common.cuh"
#include "gp_tc.cuh"
#include "tc_ptx.cuh"

namespace scasml {
namespace tc {

constexpr int TM = 128;
constexpr int A_BLK = TM * 128;      // bytes of one [128 x 64] f16 block

// ---- self test: D[128 x N] = A[128 x K] B[N x K]^T with runtime descriptor fields -------------------------------
__global__ void __launch_bounds__(128, 1)
selftest_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ Dout, int K, int N,
                uint32_t lbo16, uint32_t sbo16, uint32_t layout, uint32_t kstep_bytes) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int KB = K / KBLK;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)KB * A_BLK;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int idx = tid; idx < TM * K; idx += 128) {
        const int r = idx / K, c = idx % K;
        *(__half*)(sA + (size_t)(c / KBLK) * A_BLK + sw128_off(r, c % KBLK)) = A[(size_t)r * K + c];
    }
    for (int idx = tid; idx < N * K; idx += 128) {
        const int r = idx / K, c = idx % K;
        *(__half*)(sB + (size_t)(c / KBLK) * (N * 128) + sw128_off(r, c % KBLK)) = B[(size_t)r * K + c];
    }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 64);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(TM, N);
        uint32_t acc = 0;
        for (int kb = 0; kb < KB; ++kb)
            for (int ks = 0; ks < KBLK / 16; ++ks) {
                const uint64_t ad = make_desc(smem_u32(sA + (size_t)kb * A_BLK) + ks * kstep_bytes, lbo16, sbo16, layout);
                const uint64_t bd = make_desc(smem_u32(sB + (size_t)kb * (N * 128)) + ks * kstep_bytes, lbo16, sbo16, layout);
                umma_f16(tmem_base, ad, bd, idesc, acc);
                acc = 1;
            }
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) Dout[(size_t)tid * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}



// ---- self test, A operand in tensor memory: D[128 x N] = A[128 x K] B[N x K]^T -------------------------------------
__global__ void __launch_bounds__(128, 1)
selftest_ts_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ Dout, int K, int N) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int KB = K / KBLK;
    uint8_t* sB = smem;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int idx = tid; idx < N * K; idx += 128) {
        const int r = idx / K, c = idx % K;
        *(__half*)(sB + (size_t)(c / KBLK) * (N * 128) + sw128_off(r, c % KBLK)) = B[(size_t)r * K + c];
    }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t colA = 64;                                   // A image at columns [64, 64 + K/2)
    {   // thread = row: pack two consecutive K elements per 32-bit column
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        for (int c0 = 0; c0 < K / 2; c0 += 16) {
            uint32_t w[16];
            for (int i = 0; i < 16; ++i) {
                const __half lo = A[(size_t)tid * K + 2 * (c0 + i)], hi = A[(size_t)tid * K + 2 * (c0 + i) + 1];
                w[i] = (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
            }
            tmem_st16(tmem_base + lane_addr + colA + c0, w);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint32_t idesc = make_idesc(TM, N);
        uint32_t acc = 0;
        for (int kb = 0; kb < KB; ++kb)
            for (int ks = 0; ks < KBLK / 16; ++ks) {
                const uint64_t bd = make_desc(smem_u32(sB + (size_t)kb * (N * 128)) + ks * 32, 1, 64, 2);
                umma_f16_ts(tmem_base, tmem_base + colA + (uint32_t)(kb * 4 + ks) * 8u, bd, idesc, acc);
                acc = 1;
            }
        umma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) Dout[(size_t)tid * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}


// ---- micro-benchmark: cycles per tcgen05.mma (M = 128, K = 16, f16) for a given N, number of independent
//      accumulator chains and A source; one CTA per SM, no epilogue.  Used to size tiles (profiles/). -----------------
__global__ void __launch_bounds__(128, 1)
mma_bench_kernel(int N, int nchains, int ts_mode, int iters, long long* __restrict__ cycles_out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    uint8_t* sA = smem;                       // [128 x 64] f16 block
    uint8_t* sB = smem + A_BLK;               // [256 x 64] f16 block
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (A_BLK + 256 * 128) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;   // f16 ones
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);
    if (warp == 0) {
        // converged warp, one elected lane issues: operands are warp-uniform, so UTCHMMA takes them from uniform
        // registers without the per-lane serialisation loop a divergent `if (lane == 0)` block compiles to
        const uint32_t idesc = make_idesc(TM, N);
        const uint64_t ad = make_desc(smem_u32(sA), 1, 64, 2), bd = make_desc(smem_u32(sB), 1, 64, 2);
        const uint32_t chain_stride = (uint32_t)N;
        const uint32_t el = elect_one();
        const long long t0 = clock64();
        if (ts_mode >= 2) {
            // unrolled by 4, fixed operands per slot (ts_mode 4 / 8: with concurrent tcgen05.ld traffic from the other warps)
            const uint32_t acc0 = tmem_base, acc1 = tmem_base + (nchains > 1 ? chain_stride : 0u);
            for (int it = 0; it < iters; it += 4) {
                if (el) {
                    umma_f16_ts(acc0, tmem_base + 480u, bd, idesc, 1u);
                    umma_f16_ts(acc1, tmem_base + 480u, bd + 2ull, idesc, 1u);
                    umma_f16_ts(acc0, tmem_base + 488u, bd + 4ull, idesc, 1u);
                    umma_f16_ts(acc1, tmem_base + 488u, bd + 6ull, idesc, 1u);
                }
            }
        } else {
            for (int it = 0; it < iters; ++it) {
                const uint32_t acc = tmem_base + (uint32_t)(it % nchains) * chain_stride;
                const uint32_t ks = (uint32_t)(it & 3);
                if (el) {
                    if (ts_mode) umma_f16_ts(acc, tmem_base + 480u, bd + (uint64_t)(ks * 2), idesc, it >= nchains ? 1u : 0u);
                    else umma_f16(acc, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, it >= nchains ? 1u : 0u);
                }
            }
        }
        const long long t1 = clock64();
        if (el) umma_commit(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && el) { cycles_out[0] = t1 - t0; cycles_out[1] = t2 - t0; }
    } else if (ts_mode >= 4) {
        // interference probe: the other warps stream accumulator columns out of tensor memory while the MMAs run
        float acc = 0.f;
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        const int nld = (ts_mode >= 8) ? iters * 2 : iters / 2;
        for (int it = 0; it < nld; ++it) {
            float v[16];
            tmem_ld16(tmem_base + lane_addr + 256u + (uint32_t)((it & 7) * 16), v);
            tmem_ld_wait();
            acc += v[it & 15];
        }
        if (acc == 12345.678f) cycles_out[3] = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}



__global__ void __launch_bounds__(17 * 32, 1)
pipe_bench_kernel(int mode, int N, int iters, long long* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                       // [128 x 64] f16
    uint8_t* sB = smem + 128 * 128;           // [256 x 64] f16
    uint64_t& bar = *(uint64_t*)(smem + 128 * 128 + 256 * 128);
    uint32_t& tmem_base_s = *(uint32_t*)(smem + 128 * 128 + 256 * 128 + 8);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 * 128 + 256 * 128) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x2c002c00u;   // f16 1/16
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 16) tmem_alloc(smem_u32(&tmem_base_s), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_s, 0);
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    const int cg = warp >> 2;
    if (warp < 16) {
        {   // define the columns this warp touches (finite values)
            uint32_t z[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = __float_as_uint(-0.25f * (float)i);
            for (int c = 0; c < 4; ++c) tmem_st16(tmem_base + lane_addr + (uint32_t)(c * 64 + cg * 16), z);
            tmem_st_wait();
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const long long t0 = clock64();
        float sink = 0.f;
        if (mode == 0) {
            for (int it = 0; it < iters; ++it) {
                float v[4][16];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld16(tmem_base + lane_addr + (uint32_t)(c * 64 + cg * 16), v[c]);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 4; ++c) sink += v[c][it & 15];
            }
        } else if (mode == 1) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = 0x3c003c00u + (uint32_t)i;
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int c = 0; c < 8; ++c) tmem_st8(tmem_base + lane_addr + (uint32_t)((c & 3) * 64 + cg * 16 + (c >> 2) * 8), w);
                w[it & 7] ^= 1u;
            }
            tmem_st_wait();
        } else if (mode == 2) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = -0.01f * (float)(i + tid);
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = ex2f(v[i]) - 1.5f;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) sink += v[i];
        } else if (mode == 3 || mode == 4 || mode == 9) {
            for (int it = 0; it < iters; ++it) {
                const uint32_t col = tmem_base + lane_addr + (uint32_t)((it & 3) * 64 + cg * 16);
                float v[16];
                tmem_ld16(col, v);
                tmem_ld_wait();
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float p0 = ex2f(v[2 * i]), p1 = ex2f(v[2 * i + 1]);
                    const __half2 h = __floats2half2_rn(p0, p1);
                    const float2 hf = __half22float2(h);
                    const __half2 l = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
                    hi[i] = *(const uint32_t*)&h;
                    lo[i] = *(const uint32_t*)&l;
                }
                tmem_st8(col, hi);
                tmem_st8(col + 8, lo);
                tmem_st_wait();
                // restore an accumulator-like value so the next visit sees finite inputs
                uint32_t z[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) z[i] = __float_as_uint(-0.25f * (float)i);
                if ((it & 63) == 63) { tmem_st16(col, z); tmem_st_wait(); }
            }
        }
        if (mode == 12 || mode == 13) {
            double a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = 1.0 + i + tid;
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fma(a[i], 1.0000001, 0.5);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) sink += (float)a[i];
        } else if (mode == 14 || mode == 15) {
            float a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = 1.0f + i + tid;
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], 1.0000001f, 0.5f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) sink += a[i];
        }
        const long long t1 = clock64();
        if (sink == 12345.678f) out[7] = 1;
        if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
    } else {
        const uint32_t el = elect_one();
        if (mode >= 4 && mode != 13 && mode != 15) {
            const uint32_t idesc = make_idesc(128, N);
            const uint64_t ad = make_desc(smem_u32(sA), 1, 64, 2), bd = make_desc(smem_u32(sB), 1, 64, 2);
            const int nmma = (mode == 4 || mode == 9) ? iters * 2 : (mode >= 12 ? iters : (mode >= 10 ? 38 * 32 : iters));
            uint64_t& bar2 = *(uint64_t*)(smem + 128 * 128 + 256 * 128 + 16);
            uint64_t& bar3 = *(uint64_t*)(smem + 128 * 128 + 256 * 128 + 24);
            if (el) { mbar_init(smem_u32(&bar2), 1); mbar_init(smem_u32(&bar3), 1); mbar_arrive(smem_u32(&bar3)); fence_barrier_init(); }
            __syncwarp();
            const uint32_t acc = tmem_base + 256u;
            const long long t0 = clock64();
            if (mode == 10 || mode == 11) {
                const uint32_t id128 = make_idesc(128, 128), id16 = make_idesc(128, 16);
                for (int it = 0; it < nmma; it += 38) {
                    if (el) {
                        if (mode == 10) {
#pragma unroll
                            for (int q = 0; q < 12; ++q) {
                                umma_f16_ts(acc, tmem_base + 480u + (uint32_t)(q & 1) * 8u, bd + (uint64_t)((q & 3) * 2), id128, 1u);
                                umma_f16_ts(tmem_base + 448u, tmem_base + 496u, bd, id16, 1u);
                                umma_f16_ts(tmem_base + 448u, tmem_base + 504u, bd + 2ull, id16, 1u);
                            }
                            umma_f16_ts(acc, tmem_base + 480u, bd, id128, 1u);
                            umma_f16_ts(acc, tmem_base + 488u, bd + 2ull, id128, 1u);
                        } else {
#pragma unroll
                            for (int q = 0; q < 14; ++q) umma_f16_ts(acc, tmem_base + 480u + (uint32_t)(q & 1) * 8u, bd + (uint64_t)((q & 3) * 2), id128, 1u);
#pragma unroll
                            for (int q = 0; q < 24; ++q) umma_f16_ts(tmem_base + 448u, tmem_base + 496u + (uint32_t)(q & 1) * 8u, bd + (uint64_t)((q & 3) * 2), id16, 1u);
                        }
                    }
                    __syncwarp();
                }
            } else if (mode == 7 || mode == 8) {
                for (int it = 0; it < nmma; it += 7) {
                    if (mode == 8) { mbar_wait(smem_u32(&bar3), 0); tc_fence_after(); }
                    if (el) {
#pragma unroll
                        for (int q = 0; q < 7; ++q) umma_f16_ts(acc, tmem_base + 480u + (uint32_t)(q & 1) * 8u, bd + (uint64_t)((q & 3) * 2), idesc, 1u);
                        umma_commit(smem_u32(&bar2));
                    }
                    __syncwarp();
                }
            } else
            for (int it = 0; it < nmma; it += 4) {
                if (el) {
                    if (mode == 6 || mode == 9 || mode == 12 || mode == 14) {
                        umma_f16_ts(acc, tmem_base + 480u, bd, idesc, 1u);
                        umma_f16_ts(acc, tmem_base + 488u, bd + 2ull, idesc, 1u);
                        umma_f16_ts(acc, tmem_base + 480u, bd + 4ull, idesc, 1u);
                        umma_f16_ts(acc, tmem_base + 488u, bd + 6ull, idesc, 1u);
                    } else {
                        umma_f16(acc, ad, bd, idesc, 1u);
                        umma_f16(acc, ad + 2ull, bd + 2ull, idesc, 1u);
                        umma_f16(acc, ad + 4ull, bd + 4ull, idesc, 1u);
                        umma_f16(acc, ad + 6ull, bd + 6ull, idesc, 1u);
                    }
                }
            }
            const long long t1 = clock64();
            if (el) umma_commit(smem_u32(&bar));
            __syncwarp();
            mbar_wait(smem_u32(&bar), 0);
            const long long t2 = clock64();
            if (blockIdx.x == 0 && el) { out[1] = t1 - t0; out[2] = t2 - t0; out[3] = nmma; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc

int tc_pipe_bench(int mode, int N, int iters, long long* out_dev, cudaStream_t stream) {
    SC_REQUIRE(mode >= 0 && mode <= 15 && N >= 16 && N <= 256 && N % 16 == 0 && iters > 0 && iters % 448 == 0, "pipe_bench: arguments");
    const size_t smem = 128 * 128 + 256 * 128 + 64;
    SC_CUDA(cudaFuncSetAttribute(tc::pipe_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SC_CUDA(cudaMemsetAsync(out_dev, 0, 8 * sizeof(long long), stream));
    tc::pipe_bench_kernel<<<148, 17 * 32, smem, stream>>>(mode, N, iters, out_dev);
    SC_LAUNCH_CHECK();
    return OK;
}

int tc_mma_bench(int N, int nchains, int ts_mode, int iters, long long* cycles_dev, cudaStream_t stream) {
    SC_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && nchains >= 1 && nchains * N <= 448, "mma_bench: shape");
    const size_t smem = tc::A_BLK + 256 * 128 + 1024;
    SC_CUDA(cudaFuncSetAttribute(tc::mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc::mma_bench_kernel<<<148, 128, smem, stream>>>(N, nchains, ts_mode, iters, cycles_dev);
    SC_LAUNCH_CHECK();
    return OK;
}

int tc_selftest(const void* A_dev, const void* B_dev, float* D_dev, int K, int N, unsigned lbo16, unsigned sbo16,
                unsigned layout, unsigned kstep_bytes, cudaStream_t stream) {
    SC_REQUIRE(K % tc::KBLK == 0 && K >= 64 && K <= 256, "selftest: K must be a multiple of 64");
    SC_REQUIRE(N % 16 == 0 && N >= 16 && N <= 64, "selftest: N in [16, 64], multiple of 16");
    const size_t smem = 1024 + (size_t)(K / tc::KBLK) * (tc::A_BLK + (size_t)N * 128);
    SC_CUDA(cudaFuncSetAttribute(tc::selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (layout == 100) {      // A operand in tensor memory (TS mode)
        SC_CUDA(cudaFuncSetAttribute(tc::selftest_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc::selftest_ts_kernel<<<1, 128, smem, stream>>>((const __half*)A_dev, (const __half*)B_dev, D_dev, K, N);
        SC_LAUNCH_CHECK();
        return OK;
    }
    tc::selftest_kernel<<<1, 128, smem, stream>>>((const __half*)A_dev, (const __half*)B_dev, D_dev, K, N, lbo16, sbo16,
                                                  layout, kstep_bytes);
    SC_LAUNCH_CHECK();
    return OK;
}

}  // namespace scasml
