// tcgen05 route: per-fit operand images and coefficient records (gp_eval_tc.cu).
#pragma once
#include "gp.cuh"

namespace scasml {

// FP32 coefficient records of the tcgen05 epilogue.  Every GP weight is pre-multiplied by K_j = exp(-a |y_j|^2 / 2);
// the closed-form functionals are expanded against the per-point monomials
// 1, sx, xt, sx^2, sx xt, xt^2 (sx = sum_i x_i, xt = time) and the 5-index sums P1, P2, R1, R2.
enum TcRecA : int {          // k class (distance x - y) and ky class (x - roll y)
    RA_U0 = 0, RA_U1 = 1, RA_U2 = 2, RA_Y0 = 3,
    RA_Y1 = 4, RA_Y2 = 5, RA_Y3 = 6, RA_Y4 = 7,
    RA_G0 = 8, RA_GSX = 9, RA_GXT = 10, RA_GSXXT = 11,
    RA_GSX2 = 12, RA_SYR = 13, RA_Y0T = 14, RA_T2 = 15,
    RA_T0 = 16, RA_TXT = 17, RA_TSX = 18, RA_TXT2 = 19,
    RA_TSXXT = 20, RA_LW = 21,
};
enum TcRecB : int { RB_X0 = 0, RB_X1 = 1, RB_X2 = 2, RB_X3 = 3, RB_X4 = 4, RB_Q2 = 5 };   // kx class (roll x - y)
constexpr int TC_NFA = 24;
constexpr int TC_NFB = 8;
// PDE kernel records (three item kinds): kind 0 = k class [U0 U1 U2 Gsx2 | G0 Gsx Gxt Gsxxt | T0 Txt Tsx Txt2 | Tsxxt T2 Lw 0],
// kind 1 = ky class [Y0 Y1 Y2 Y3 | Y4 syr y0 0], kind 2 = kx class (= TcRecB)
constexpr int TC_NF0 = 16;
constexpr int TC_NF1 = 8;

struct TcState {
    uint8_t* images = nullptr;   // per centre tile: [C image][Croll image][records A][records B]
    size_t tile_bytes = 0;
    int KB = 0;                  // 64-wide K blocks of the permuted contraction axis
    int ntile_dom = 0, ntile_bdy = 0;
    double inv_ascale = 0.0;     // 1 / (a log2 e): un-scales the 5-index accumulators
    short perm[128];             // permuted K slot -> source coordinate (-1: zero padding)
    short iperm[128];            // source coordinate -> slot in k-steps {0, 2, 3, ...}
    short iperm1[128];           // source coordinate -> slot in k-step 1 (successors of the index set), else -1
    const short* tabs = nullptr; // device copy of [perm | iperm | iperm1] (constant-bank lookups with divergent indices serialise)
    long long* dbg = nullptr;    // optional timeline buffer (clock64 stamps of CTA dbg_block), see scasml_debug_tc_timeline
    int dbg_block = 0;
};

int tc_supported(const GpView& gp);
size_t tc_image_bytes(const GpView& gp, TcState* st);              // fills KB / tile counts / tile_bytes / perm
int tc_build_images(const GpView& gp, const TcState& st, cudaStream_t stream);
int tc_timeline(const GpView& gp, const TcState& st, const double* X, long R, int mode, int block, long long* stamps_dev,
                double* scratch_out, cudaStream_t stream);
int tc_mma_bench(int N, int nchains, int ts_mode, int iters, long long* cycles_dev, cudaStream_t stream);
int tc_pipe_bench(int mode, int N, int iters, long long* out_dev, cudaStream_t stream);   // tc_bench.cu
int tc_selftest(const void* A_dev, const void* B_dev, float* D_dev, int K, int N, unsigned lbo16, unsigned sbo16,
                unsigned layout, unsigned kstep_bytes, cudaStream_t stream);

}  // namespace scasml
