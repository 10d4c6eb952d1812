// tcgen05 route: per-fit operand images of the two chained GEMMs (gp_eval_tc.cu).
#pragma once
#include "gp.cuh"

namespace scasml {

// evaluation classes of the tcgen05 route
enum TcClass : int { TC_U = 0, TC_UG = 1, TC_PDE = 2 };
// kernel classes: k (x - y), ky (x - roll y), kx (roll x - y)
enum TcKernel : int { TK_K = 0, TK_KY = 1, TK_KX = 2 };
// outputs a coefficient column contributes to
enum TcOut : int { TO_U = 0, TO_G = 1, TO_L = 2, TO_T = 3, TO_PAD = 255 };
// per-point features: a column's monomial is F[f1] * F[f2]
enum TcFeat : int { TF_ONE = 0, TF_SX = 1, TF_XT = 2, TF_X0 = 3, TF_SXR = 4, TF_P2 = 5, TF_R2 = 6, TF_XI = 7, TF_XR = 12, TF_COUNT = 17 };
// per-centre coefficient formulas (coef_of() in gp_eval_tc.cu; NumPy statement: tests/tc_expansion_ref.py)
enum TcCoef : int {
    TCF_U0, TCF_U1, TCF_U2, TCF_G0, TCF_GSX, TCF_GXT, TCF_GSXXT, TCF_GSX2, TCF_T0, TCF_TXT, TCF_TSX, TCF_TXT2, TCF_TSXXT,
    TCF_L1, TCF_LR2, TCF_LR22, TCF_LX, TCF_LXR2, TCF_LXX, TCF_H, TCF_HSX, TCF_HG, TCF_HXT, TCF_HT, TCF_MX, TCF_ZERO
};

constexpr int TC_MAXCOL = 128;       // accumulator columns of the coefficient GEMM (all kernel classes of one evaluation class)
constexpr int TC_P_SHIFT = 6;        // P = 2^6 exp(a x.y): keeps the low halves of the f16 split in the normal range

struct TcColSpec {                   // one coefficient column (host-built table, mirrored on the device)
    unsigned char kern, out, f1, f2, coef, m, n, pad;
};
struct TcColDesc {                   // what the kernel's final contraction needs of a column
    unsigned char f1, f2, out, pad;
    float inv_scale_f;               // the same power of two in FP32 (exact; 0 when it leaves the FP32 range: the FP64 path is used then)
    double inv_scale;                // 1 / (column scale 2^s * 2^TC_P_SHIFT)
};

struct TcState {
    uint8_t* images = nullptr;       // one allocation: [B1 images][B3 images x 3 classes][column tables][scratch]
    size_t total_bytes = 0;
    int KB = 0;                      // 64-wide K blocks of the contraction axis
    int nstep = 0;                   // k-steps of 16 actually used: ceil((D + 1) / 16) (column D carries the exponent shift)
    int tn = 64, compact = 0;        // centres per tile; 1: tiles run over the compact centre list [domain | boundary] (resident-operand kernel)
    int ntile_dom = 0, ntile_bdy = 0, ntile_all = 0, npair_all = 0;
    size_t b1_off = 0, b1_tile_bytes = 0;            // per tile, per K block: rows [C | Crollinv | Croll] x 128 B
    int ncol[3][3] = {};                             // padded column count per (evaluation class, kernel class)
    size_t b3_off[3] = {}, b3_tile_bytes[3] = {};    // per tile: [k hi | k lo | ky hi | ky lo | kx hi | kx lo], ncol x 128 B each
    size_t desc_off[3] = {};                         // TcColDesc[TC_MAXCOL] per evaluation class
    size_t spec_off = 0;                             // TcColSpec[3][TC_MAXCOL]
    size_t ymax_off = 0;                             // one double: max_j |y_j|^2 (per-row exponent shift)
    size_t csum_off = 0;                             // double [3][TC_MAXCOL]: column sums of the split coefficients (baseline term)
    size_t scratch_off = 0;                          // double [3][ncentres][TC_MAXCOL] + column maxima
    TcColSpec spec[3][TC_MAXCOL];                    // host copy
};
struct TcDebug {                     // debug build only: clock64 stamps of CTA (block & 0xFFFFFF), experiment flags in block >> 24
    long long* stamps = nullptr;
    int block = 0;
};

// Per-point operand record of the resident-operand kernel (d + 2 <= 128): what its A operand needs of a point, written by whoever
// produces the point (the Picard samplers, while the coordinates are in registers) or by a pre-pass over a caller's X:
//   [ 16 nstep words (hi | lo << 16), f16 | |x|^2 f64 | sum_{i<d} x_i f64 ],   a log2(e) x_j = hi_j + lo_j   (columns >= D are zero)
// The record size is an odd multiple of 16 bytes: a 128-point tile is ONE contiguous bulk copy into shared memory, and the
// row-per-thread 16-byte reads that move it on into tensor memory are bank-conflict free.
__host__ __device__ inline int tc_rec_nstep(int D) {               // instantiated k-step counts; 0: K-streamed kernel (no records)
    const int n = (D + 1 + 15) / 16;
    return n <= 2 ? 2 : (n <= 4 ? 4 : (n <= 7 ? 7 : (n <= 8 ? 8 : 0)));
}
__host__ __device__ inline int tc_rec_bytes(int nstep) { return nstep * 64 + 16; }
// ... and, in a second array, the 12 coordinates the final contraction's monomials are made of, one contiguous 96-byte block per point:
// x_0, t, x_{I[m]} (m < 5), x_{I[m] + 1} (m < 5)   (I = the Hutchinson index set; models/GP.py:91-93, 139-180)
constexpr int TC_REC_NFEAT = 12;
struct TcRecIdx { int col[TC_REC_NFEAT]; };          // column of the FP64 row each feature is
inline TcRecIdx tc_rec_idx(const GpView& gp) {
    TcRecIdx r;
    r.col[0] = 0; r.col[1] = gp.d;
    for (int m = 0; m < MC_IDX; ++m) { r.col[2 + m] = gp.I[m]; r.col[2 + MC_IDX + m] = gp.I[m] + 1; }
    return r;
}
__host__ __device__ inline float tc_rec_ascale(double a) { return (float)(a * 1.4426950408889634); }   // S = log2 of exp(a x.y)

int tc_supported(const GpView& gp);
size_t tc_image_bytes(const GpView& gp, TcState* st);              // fills the layout fields and the column table
int tc_build_images(const GpView& gp, const TcState& st, cudaStream_t stream);
// rec / recf: the points' operand records [R][tc_rec_bytes] and feature blocks [R][TC_REC_NFEAT] (resident-operand kernel) from a caller that
// has them (the Picard samplers); null: a pre-pass over X writes them into a stream-ordered scratch buffer of this call
int launch_eval_tc(const GpView& gp, const void* tc_state, const double* X, long R, int mode,
                   double* out0, double* out1, double* out2, double* out3, cudaStream_t stream,
                   const TcDebug* dbg = nullptr, const uint8_t* rec = nullptr, const double* recf = nullptr);
int tc_timeline(const GpView& gp, const TcState& st, const double* X, long R, int mode, int block, long long* stamps_dev,
                double* scratch_out, cudaStream_t stream);
int tc_mma_bench(int N, int nchains, int ts_mode, int iters, long long* cycles_dev, cudaStream_t stream);
int tc_pipe_bench(int mode, int N, int iters, long long* out_dev, cudaStream_t stream);   // tc_bench.cu
int tc_selftest(const void* A_dev, const void* B_dev, float* D_dev, int K, int N, unsigned lbo16, unsigned sbo16,
                unsigned layout, unsigned kstep_bytes, cudaStream_t stream);

}  // namespace scasml
