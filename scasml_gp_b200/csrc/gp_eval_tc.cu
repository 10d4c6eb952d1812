// tcgen05 route of the fused surrogate evaluation (sm_100a): two chained GEMMs per (128 points x 2 x 48 centres) pair of sub-items.
//
// Same contract as gp_eval.cu (reference models/GP.py:630-687, 326-411, 746-769), different arithmetic.  Every functional the
// ScaSML correction needs of the surrogate is  sum_j kernel_c(x, y_j) polynomial(x, y_j)  with kernel_c one of
// k (x - y), ky (x - roll y), kx (roll x - y), and the Gaussian factorises, kernel_c = K_i K_j exp(a x . perm_c(y_j)).
// Expanding the polynomial in per-point monomials F[f1] F[f2] with per-centre coefficients C (tests/tc_expansion_ref.py
// states and checks the algebra against the closed forms of SURVEY App. B) turns the evaluation into
//     S = (a' x) perm_c(Y)^T          stage 1: distance GEMM, tcgen05.mma kind::f16, TS mode (A = point images in tensor memory),
//                                     FP32 accumulators in TMEM
//     P' = 2^s ex2(S) - 2^s           epilogue warps: tcgen05.ld -> ex2 -> f16 hi/lo split -> tcgen05.st, IN PLACE over S
//     T += P' C_c                     stage 2: coefficient GEMM, A operand = P' straight from tensor memory (TS mode)
//     out_o(x_i) = K_i sum_col F_i[f1] F_i[f2] (T[i, col] + 2^s sum_j C[j, col]) / scale_col     (FP32 products, FP64 sums)
// so the per-pair work outside the tensor pipe is one TMEM load, one ex2 and three conversion/subtract instructions: both
// measured per-pair limits of the SM (TMEM read ~16 FP32/clk, MUFU 16 ex2/clk; tools/tc_pipe_bench.py) instead of the
// 10-45 FP32 instructions per pair of a closed-form epilogue.
//   * centres must be float16-valued (DeepXDE float16 collocation points; checked in scasml_gp_set_centres, otherwise the
//     route is refused): the stage-1 B operand is EXACT in f16; sampled points are split a' x = hi + lo (two f16 terms,
//     2 MMA passes into the same accumulator);
//   * the exponent shift s (6 by default, lowered per point row when a |x| max|y| would overflow f16) is applied after ex2 and
//     divided out again with K_i; the constant 2^s is subtracted before the coefficient GEMM because the tensor core TRUNCATES its
//     FP32 accumulation (~0.66 ulp per MMA towards zero, tools/tc_accum_probe.py): the bias is proportional to the running sum;
//   * P and the column-scaled coefficients are split hi + lo in f16; products hh + hl + lh accumulate into the same FP32
//     TMEM columns over all centre tiles (relative error ~2e-7 rms, tests/test_tc_expansion.py emulates it);
//   * one N = 96 MMA per k-step serves a PAIR of sub-items (48-centre tile x kernel class): the centre image rows are
//     [C | rollinv(C) | roll(C)], because roll(x) . y = x . rollinv(y) needs no second A operand;
//   * warp roles (23 warps, NTHREADS_P): 4 contraction warps (final contraction of each finished class), 16 epilogue warps in two
//     groups of 8 (group = pair parity), 1 producer warp (cp.async.bulk operand rings), 2 + 2 MMA-issuing warps (distance / coefficient
//     GEMMs of the even / odd pairs; converged, one elected lane issues); three S/P slots;
//   * a point tile arrives as ONE bulk copy of its 128 operand records (gp_tc.cuh: f16 hi/lo words, |x|^2, sum x -- written by the Picard
//     samplers or by rec_image_kernel below), the 12 feature coordinates of a row as one 96-byte block: the kernel never reads the FP64 rows.
// Accuracy: ~2e-7 relative rms on every output; parity with the FP64 route is tested under the "nocast" policy.
// Timeline stamps and experiment flags exist only in the debug build (-DSCASML_DEBUG_HOOKS, libscasml_b200_dbg.so).
#include <cmath>
#include <cstring>
#include "picard.cuh"
#include "gp_tc.cuh"
#include "tc_ptx.cuh"

namespace scasml {

namespace tc {

constexpr int TM = 128;              // points per CTA (UMMA M)
constexpr int TN = 64;               // centres per tile
constexpr int A_BLK = TM * 128;      // bytes of one [128 x 64] f16 block
constexpr int NEPI = 16;             // epilogue warps
#ifdef SCASML_DEBUG_HOOKS
constexpr bool DBG = true;           // clock64 stamps + run-time experiment flags (tools/tc_timeline.py)
#else
constexpr bool DBG = false;          // product build: every dbg branch below folds away
#endif

// slim device view of TcState
struct TcDev {
    const uint8_t* b1;               // stage-1 centre images
    const uint8_t* b3;               // stage-2 coefficient images of this evaluation class
    const TcColDesc* desc;           // column table of this evaluation class
    size_t b1_tile_bytes, b3_tile_bytes;
    int nstep, ntile_dom, ntile_bdy;  // K-streamed kernel: 64-centre tiles of the padded domain / boundary sets
    int npair_all;                   // resident-operand kernel: pairs of 48-centre tiles in the images (compact format)
    int ntile_all;                   // resident-operand kernel: 48-centre tiles of the compact centre list (k, kx classes); ntile_dom = those of the ky class
    const double* csum;              // [TC_MAXCOL] per column: sum over the centres of the scaled, split coefficients (baseline term of the coefficient GEMM)
    const uint8_t* rec;              // resident-operand kernel: the points' operand records [R][tc_rec_bytes(NSTEP)] (gp_tc.cuh)
    const double* recf;              // ... and their feature blocks [R][TC_REC_NFEAT]: x_0, t, x_I[0..4], x_{I+1}[0..4]
    const double* ymax2;             // device scalar: max_j |y_j|^2 over the centres (per-row exponent shift, row_shift())
    long long* dbg;
    int dbg_block;
    int dbg_flags;                   // debug build: 1 skip stage-2 MMAs, 2 skip the epilogue arithmetic, 4 watchdog (a wait that runs out of probes names itself in the stamp buffer), 8 stage-2 hand-off by arrive instead of commit, 32 hh products only, 128 hl / lh products into the other T buffer
};

// ---- per-centre coefficient of one column (NumPy statement: tests/tc_expansion_ref.py::centre_coefficient) --------------
__device__ double coef_of(const TcColSpec& sp, const double* __restrict__ f, double a, int d) {
    const double Kj = exp(-0.5 * a * f[CF_NY]);
    const double A1 = f[CF_A1] * Kj, A3 = f[CF_A3] * Kj, A4 = f[CF_A4] * Kj, A5 = f[CF_A5] * Kj;
    const double sy = f[CF_SY], yt = f[CF_YT], y0 = f[CF_Y0], syr = f[CF_SYROLL], dd = (double)d;
    const double a2 = a * a, a3 = a2 * a, a4 = a2 * a2;
    double Q1 = 0, Q2 = 0, T1 = 0, T2 = 0;
    for (int m = 0; m < MC_IDX; ++m) {
        Q1 += f[CF_YI + m]; Q2 += f[CF_YI + m] * f[CF_YI + m];
        T1 += f[CF_YIR + m]; T2 += f[CF_YIR + m] * f[CF_YIR + m];
    }
    const int m = sp.m, n = sp.n;
    const double LW = A3 * dd * dd / (MC_IDX * MC_IDX);
    const double w3 = A3 * dd;
    auto hc = [&](int i) {             // h = w3 (a^2/5 (P2 + T2 - 2 sum xI yr) - a) against [1, P2, xI_0..4]
        return i == 0 ? w3 * (a2 / MC_IDX * T2 - a) : (i == 1 ? w3 * a2 / MC_IDX : -2.0 * w3 * a2 / MC_IDX * f[CF_YIR + i - 2]);
    };
    switch (sp.coef) {
        case TCF_U0: return A1 - a * A4 * yt - a * A5 * sy;
        case TCF_U1: return a * A4;
        case TCF_U2: return a * A5;
        case TCF_G0: return a * A1 * sy - a2 * A4 * yt * sy + a * dd * A5 - a2 * A5 * sy * sy;
        case TCF_GSX: return -a * A1 + a2 * A4 * yt + 2.0 * a2 * A5 * sy;
        case TCF_GXT: return a2 * A4 * sy;
        case TCF_GSXXT: return -a2 * A4;
        case TCF_GSX2: return -a2 * A5;
        case TCF_T0: return a * A1 * yt + a * A4 - a2 * A4 * yt * yt - a2 * A5 * yt * sy;
        case TCF_TXT: return -a * A1 + 2.0 * a2 * A4 * yt + a2 * A5 * sy;
        case TCF_TSX: return a2 * A5 * yt;
        case TCF_TXT2: return -a2 * A4;
        case TCF_TSXXT: return -a2 * A5;
        // lap_x lap_y = LW (a^4 q2^2 - 14 a^3 q2 + 35 a^2), q2 = R2 + T2 - 2 sum xr yr
        case TCF_L1: return LW * (a4 * T2 * T2 - 14.0 * a3 * T2 + 35.0 * a2);
        case TCF_LR2: return LW * (2.0 * a4 * T2 - 14.0 * a3);
        case TCF_LR22: return LW * a4;
        case TCF_LX: return LW * (-4.0 * a4 * T2 + 28.0 * a3) * f[CF_YIR + m];
        case TCF_LXR2: return -4.0 * LW * a4 * f[CF_YIR + m];
        case TCF_LXX: return LW * 4.0 * a4 * f[CF_YIR + m] * f[CF_YIR + n] * (m == n ? 1.0 : 2.0);
        case TCF_H: return hc(m);
        case TCF_HSX: return -a * hc(m);
        case TCF_HG: return a * syr * hc(m) + (m == 0 ? -w3 * 2.0 * a2 / MC_IDX * T1 : (m >= 2 ? w3 * 2.0 * a2 / MC_IDX : 0.0));
        case TCF_HXT: return -a * hc(m);
        case TCF_HT: return a * y0 * hc(m);
        case TCF_MX: {                 // kx class: (a^2/5 (R2 + Q2 - 2 sum xr yI) - a) (X0 + X1 x0 + X2 sxr) + X3 + X4 R1
            const double mc = m == 0 ? a2 / MC_IDX * Q2 - a : (m == 1 ? a2 / MC_IDX : -2.0 * a2 / MC_IDX * f[CF_YI + m - 2]);
            const double pc = n == 0 ? dd * (A1 - a * A4 * yt - a * A5 * sy) : (n == 1 ? dd * a * A4 : dd * a * A5);
            double v = mc * pc;
            if (n == 0 && m == 0) v += dd * 2.0 * a2 / MC_IDX * A5 * Q1;
            if (n == 0 && m >= 2) v += -dd * 2.0 * a2 / MC_IDX * A5;
            return v;
        }
        default: return 0.0;
    }
}

// pass 1: FP64 coefficient matrix [class][centre][TC_MAXCOL] + per-column maxima (column scaling of the f16 split)
__global__ void coef_kernel(GpView gp, const TcColSpec* __restrict__ spec, int ncentres, double* __restrict__ coef,
                            unsigned long long* __restrict__ colmax) {
    const int cls = blockIdx.y;
    const int j = blockIdx.x;                                  // padded centre index
    const int col = threadIdx.x;                               // TC_MAXCOL threads
    const TcColSpec sp = spec[cls * TC_MAXCOL + col];
    double v = 0.0;
    const bool dom = j < gp.NdPad;
    if (sp.out != TO_PAD && (dom || sp.kern != TK_KY)) v = coef_of(sp, gp.feat + (size_t)j * CF_STRIDE, gp.a, gp.d);
    coef[((size_t)cls * ncentres + j) * TC_MAXCOL + col] = v;
    const double av = fabs(v);
    if (av > 0.0 && av < 1e300) atomicMax(colmax + cls * TC_MAXCOL + col, (unsigned long long)__double_as_longlong(av));
}

// byte offset of element (row r, col c) inside a [rows x 64] f16 block, Swizzle<3,4,3> (r may exceed 63: 8-row groups are 1024 B apart)
__device__ __forceinline__ uint32_t sw_off(int r, int c) {
    return (uint32_t)(r * 128 + ((((c >> 3) ^ (r & 7)) & 7) << 4) + (c & 7) * 2);
}

// Centre tiles come in two formats (TcState::tn / compact):
//   resident-operand kernel: 48 centres per tile over the COMPACT centre list [domain 0..Nd) | boundary 0..Nb)] -- 1 200 centres
//   are exactly 25 tiles, the ky class (domain centres only) runs over the first ceil(Nd / 48) of them (boundary centres carry
//   zero ky coefficients);
//   K-streamed kernel: 64 centres per tile over the padded list of GpView ([domain | pad to 64 | boundary | pad to 64]).
// Row r of tile `tile` -> index into GpView::C / feat, or -1 for a padding row (zero centre, zero coefficients).
__device__ __forceinline__ int tile_centre(const GpView& gp, int tile, int r, int tn, int compact) {
    const int jc = tile * tn + r;
    if (!compact) return jc;
    if (jc < gp.Nd) return jc;
    if (jc < gp.Nd + gp.Nb) return gp.NdPad + (jc - gp.Nd);
    return -1;
}

// pass 2a: stage-1 centre images, f16, 128 B swizzle.  Columns >= D are zero (the exponent shift of a point row is applied after
// ex2, see row_shift()).
//   padded format (K-streamed kernel): per tile, per K block: rows [C (64) | rollinv(C) (64) | roll(C) (64)];
//   compact format (resident-operand kernel): per row set v in {C, rollinv(C), roll(C)}, per PAIR of tiles, per K block: [tile 2p rows (48) |
//   tile 2p + 1 rows (48)] -- exactly the shared-memory image of a pair's ring slot, so the producer moves a pair with ONE bulk copy (the lone
//   producer thread needed ~1 300 cycles per pair for six copies, two waits and two expect_tx: tools/tc_timeline.py).
__global__ void b1_image_kernel(GpView gp, uint8_t* __restrict__ b1, size_t tile_bytes, int KB, int tn, int compact, int npair_all) {
    const int tile = blockIdx.x;
    uint8_t* base = b1 + (size_t)tile * tile_bytes;
    const int D = gp.D;
    const int TN = tn;
    const size_t pairb = (size_t)KB * 2 * TN * 128;                   // compact: bytes of a pair block
    for (int idx = threadIdx.x; idx < TN * KB * KBLK; idx += blockDim.x) {
        const int r = idx / (KB * KBLK), c = idx % (KB * KBLK);
        const int cj = tile_centre(gp, tile, r, tn, compact);
        const double* y = gp.C + (size_t)(cj < 0 ? 0 : cj) * D;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0;
        if (c < D && cj >= 0) {
            v0 = y[c];
            v1 = y[(c == 0) ? D - 1 : c - 1];            // roll(x) . y = x . rollinv(y),  rollinv(y)_c = y_{c-1}
            v2 = y[(c + 1 == D) ? 0 : c + 1];            // roll(y)_c = y_{c+1}            (models/GP.py:91-93)
        }
        if (compact) {
            uint8_t* blk = b1 + (size_t)(tile >> 1) * pairb + (size_t)(c / KBLK) * (2 * TN * 128) + (size_t)(tile & 1) * (TN * 128) + sw_off(r, c % KBLK);
            *(__half*)(blk) = __double2half(v0);
            *(__half*)(blk + (size_t)npair_all * pairb) = __double2half(v1);
            *(__half*)(blk + (size_t)2 * npair_all * pairb) = __double2half(v2);
        } else {
            uint8_t* blk = base + (size_t)(c / KBLK) * (3 * TN * 128);
            *(__half*)(blk + sw_off(r, c % KBLK)) = __double2half(v0);
            *(__half*)(blk + sw_off(TN + r, c % KBLK)) = __double2half(v1);
            *(__half*)(blk + sw_off(2 * TN + r, c % KBLK)) = __double2half(v2);
        }
    }
}

// pass 2b: stage-2 coefficient images of one evaluation class, each [ncol rows x 64 K slots] f16 of which the first tn are centres
// (K-major B operand, 128 B swizzle); column table with the inverse scales.
//   padded format: per tile [k hi | k lo | kx hi | kx lo | ky hi | ky lo];
//   compact format: per kernel class, per PAIR of tiles: [tile 2p hi | tile 2p lo | tile 2p + 1 hi | tile 2p + 1 lo] (one bulk copy per pair).
__global__ void b3_image_kernel(GpView gp, const TcColSpec* __restrict__ spec, int cls, int ncentres, const double* __restrict__ coef,
                                const unsigned long long* __restrict__ colmax, uint8_t* __restrict__ b3, size_t tile_bytes,
                                int nk, int nkx, int nky, TcColDesc* __restrict__ desc, int tn, int compact, int npair_all) {
    const int tile = blockIdx.x;
    uint8_t* base = b3 + (size_t)tile * tile_bytes;
    const int ntot = nk + nkx + nky;
    const int TN = tn;
    for (int idx = threadIdx.x; idx < ntot * TN; idx += blockDim.x) {
        const int col = idx / TN, j = idx % TN;
        const int cj = tile_centre(gp, tile, j, tn, compact);
        const double cm = __longlong_as_double((long long)colmax[cls * TC_MAXCOL + col]);
        int e = 0;
        if (cm > 0.0) frexp(cm, &e);
        const double scale = (cm > 0.0) ? ldexp(1.0, 13 - e) : 1.0;       // column maximum -> [2^12, 2^13)
        const double v = (cj < 0) ? 0.0 : coef[((size_t)cls * ncentres + (size_t)cj) * TC_MAXCOL + col] * scale;
        const __half hi = __double2half(v);
        const __half lo = __double2half(v - (double)__half2float(hi));
        int row = col;
        size_t off = 0;
        int n = nk;
        if (col >= nk + nkx) { row = col - nk - nkx; off = (size_t)2 * (nk + nkx) * 128; n = nky; }
        else if (col >= nk) { row = col - nk; off = (size_t)2 * nk * 128; n = nkx; }
        if (compact) {
            // kernel-class regions in the column-table order [k | kx | ky]; `off` = 2 * 128 * (columns before this class)
            uint8_t* blk = b3 + (size_t)npair_all * 2 * off + (size_t)(tile >> 1) * (4 * n * 128) + (size_t)(tile & 1) * (2 * n * 128);
            *(__half*)(blk + sw_off(row, j)) = hi;
            *(__half*)(blk + (size_t)n * 128 + sw_off(row, j)) = lo;
        } else {
            *(__half*)(base + off + sw_off(row, j)) = hi;
            *(__half*)(base + off + (size_t)n * 128 + sw_off(row, j)) = lo;
        }
        if (tile == 0 && j == 0) {
            const TcColSpec sp = spec[cls * TC_MAXCOL + col];
            TcColDesc dsc;
            dsc.f1 = sp.f1; dsc.f2 = sp.f2; dsc.out = sp.out;
            dsc.inv_scale = 1.0 / (scale * (double)(1 << TC_P_SHIFT));
            const bool f32ok = dsc.inv_scale > 1e-30 && dsc.inv_scale < 1e30;      // a power of two: exact in FP32 when in range
            dsc.pad = f32ok ? 0 : 1;                                               // 1: the contraction scales this column in FP64
            dsc.inv_scale_f = f32ok ? (float)dsc.inv_scale : 1.0f;
            desc[col] = dsc;
        }
    }
}

// Exponent shift of a point row: P = 2^s exp(a x.y) must stay below the f16 maximum (2^16) for every centre, and
// |a x.y| log2(e) <= a log2(e) |x| max|y| (Cauchy-Schwarz).  s = TC_P_SHIFT inside the collocation box (a x.y <= 4 + 4/d);
// points that left it far enough -- small d, long Brownian excursions, user test points outside the box -- get a smaller,
// still integer (exact in f16) shift instead of saturated P values.  The final contraction multiplies 2^(TC_P_SHIFT - s) back.
__device__ __forceinline__ int row_shift(double a, double nx, double ymax2) {
    const double smax = a * 1.4426950408889634 * sqrt(nx * ymax2);
    if (!(smax > 15.9 - (double)TC_P_SHIFT)) return TC_P_SHIFT;      // also NaN rows (they poison only themselves)
    const double s = floor(15.9 - smax);
    return s < -14.0 ? -14 : (int)s;
}

// max_j |y_j|^2 over all (padded) centres -> one device double (bit pattern of a non-negative double orders like an integer)
__global__ void ymax_kernel(GpView gp, int ncentres, unsigned long long* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncentres) return;
    const double ny = gp.feat[(size_t)j * CF_STRIDE + CF_NY];
    if (ny > 0.0 && ny < 1e300) atomicMax(out, (unsigned long long)__double_as_longlong(ny));
}

// Column sums of the scaled, split coefficients (exactly the values the stage-2 MMAs multiply with), in a fixed summation order:
// the coefficient GEMM runs on P - 2^s (the tensor core truncates its FP32 accumulation, ~0.66 ulp per MMA towards zero,
// tools/tc_accum_probe.py; the bias is proportional to the running sum, so the constant part 2^s sum_j C_j is taken out of it)
// and the final contraction adds 2^s csum[col] back.
__global__ void __launch_bounds__(256) csum_kernel(int ncentres, const double* __restrict__ coef, const unsigned long long* __restrict__ colmax,
                                                   double* __restrict__ csum) {
    __shared__ double part[256];
    const int cls = blockIdx.y, col = blockIdx.x;
    const double cm = __longlong_as_double((long long)colmax[cls * TC_MAXCOL + col]);
    int e = 0;
    if (cm > 0.0) frexp(cm, &e);
    const double scale = (cm > 0.0) ? ldexp(1.0, 13 - e) : 1.0;           // same scale as b3_image_kernel
    double acc = 0.0;
    for (int j = threadIdx.x; j < ncentres; j += 256) {
        const double v = coef[((size_t)cls * ncentres + j) * TC_MAXCOL + col] * scale;
        const __half hi = __double2half(v);
        const __half lo = __double2half(v - (double)__half2float(hi));
        acc += (double)__half2float(hi) + (double)__half2float(lo);
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o >= 1; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) csum[cls * TC_MAXCOL + col] = part[0];
}

// ---- the fused evaluation kernel ---------------------------------------------------------------------------
// Persistent CTAs (one per SM) loop over 128-point tiles.  Sub-item = (48-centre tile, kernel class).  The sub-item stream of a
// point tile is CLASS-MAJOR,
//   u / u + div:  k over all centre tiles, then ky over the domain tiles
//   PDE:          k, then kx over all centre tiles, then ky over the domain tiles
// so only ONE class's coefficient accumulator T (<= 48 columns) is live at a time; it is double-buffered by class epoch, and the
// tensor-memory columns that frees hold a THIRD S/P slot.  Sub-items are consumed in PAIRS of the same class (an odd class ends in
// a single): 14 stage-1 MMAs (N = 96: both sub-items' centre rows, A = a' x hi | lo from tensor memory) into one of three
// 96-column S slots, one group of epilogue warps turns S into P - 2^s in place, 2 x 9 stage-2 MMAs (A = P from tensor memory)
// add P C into T.  With two slots the period of a pair was its tensor work plus three serial hand-offs (profiles/r1_tc_timeline.md);
// the third slot keeps stage-1 / stage-2 work of other pairs available while a pair is in its epilogue.
// Warp roles (23 warps): 16 epilogue warps in two groups of 8 (group = pair parity; a warp converts one sub-item of its pair for one
// lane quadrant: three 16-column chunks, the next chunk's tcgen05.ld in flight), 1 producer warp (cp.async.bulk rings: centre rows
// three to four pairs deep, coefficient images two to four pairs deep), 2 + 2 MMA-issuing warps (distance / coefficient GEMMs of the
// even / odd pairs; converged, one elected lane issues), 4 contraction warps (one per lane quadrant): the row's features and K_i at the
// start of a tile, then the final contraction of each finished class (FP32 column products, FP64 sums).  The NEXT point tile's operand
// records are bulk-copied into the staging buffer by stage-1 issuer 0 during the current tile's main loop.
// Tensor memory (512 columns): S/P slots at 0 / 96 / 192; T buffers at 288 / 336; A images (hi | lo, 64 columns apart) at 384.
template <int CLASS> struct Cfg;
template <> struct Cfg<TC_U>   { static constexpr int NK = 16, NKX = 0,  NKY = 16; };
template <> struct Cfg<TC_UG>  { static constexpr int NK = 16, NKX = 0,  NKY = 32; };
template <> struct Cfg<TC_PDE> { static constexpr int NK = 48, NKX = 32, NKY = 48; };

// K-streamed kernel (64-centre tiles, two 128-column slots)
constexpr uint32_t COL_T = 256, COL_A = 384, A_IMG_COLS = 64;
constexpr int NSLOT = 2;                                             // S/P pair slots
constexpr uint32_t B1_BLK = TN * 128;                                // one K block of one class: [64 rows x 128 B]
// resident-operand kernel (48-centre tiles, three 96-column slots)
constexpr int TN2 = 48;                                              // centres per sub-item
constexpr int NSLOT2 = 3;                                            // S/P pair slots
constexpr uint32_t SLOTC = 2 * TN2;                                  // columns of a slot
constexpr uint32_t COL_T2 = NSLOT2 * SLOTC;                          // 288: T buffers (class epoch parity), NTMAX columns each
constexpr uint32_t NTMAX = 48;
constexpr uint32_t B1_BLK2 = TN2 * 128;                              // one K block of one class: [48 rows x 128 B]
constexpr int NB1_MAX = 6;                                           // centre-row ring depth (pairs): Ring<>::NB1 <= NB1_MAX
// coefficient-image ring depth (pairs).  A slot is refilled only after stage 2 of its previous pair COMPLETED, and a bulk copy out of
// L2 takes ~1 000 cycles to land: with two slots the coefficient GEMM of pair p waited for its images whenever the pair period fell
// below ~1 600 cycles (tools/tc_timeline.py: the stage-2 issuer never waited for P, yet needed 1 300 - 1 600 cycles per pair)
template <int CLASS, int NSTEP> struct Ring {
    static constexpr int NB3 = (CLASS == TC_PDE) ? (NSTEP == 8 ? 2 : 3) : (CLASS == TC_UG ? 2 : 4);
    static constexpr int NB1 = (CLASS == TC_PDE) ? 3 : 4;
};
// staging buffer = the 128 operand records of a point tile as they lie in global memory (gp_tc.cuh): hi image | lo image | |x|^2 | sum x; the
// record size, an odd multiple of 16, makes the 16-byte reads of 8 consecutive rows hit 8 bank groups
template <int NSTEP> struct Stage { static constexpr int REC = NSTEP * 64 + 16; };
constexpr int NLOAD = 4;                                             // loader warps
constexpr int NTHREADS_P = (NEPI + 2 + NLOAD + 3) * 32;             // loaders | epilogue | producer | stage-1 issuer 0 | stage-2 issuer 0 | stage-1 issuer 1 | stage-2 issuer 1
// The loader warps take the LOWEST warp ids: the warp scheduler favours them, and their background work (staging the next point tile,
// contracting finished classes) otherwise crawled behind the epilogue warps of its scheduler and set the tile period.  Epilogue warp e
// (0..15) = warp W_EPI0 + e; W_EPI0 is a multiple of 4, so warp % 4 is still the tensor-memory lane quadrant of every loader / epilogue warp.
constexpr int W_LOAD0 = 0, W_EPI0 = NLOAD, W_PROD = W_EPI0 + NEPI, W_S1A = W_PROD + 1, W_S2A = W_PROD + 2, W_S1B = W_PROD + 3, W_S2B = W_PROD + 4;
static_assert(W_EPI0 % 4 == 0, "lane quadrant = warp % 4");
constexpr int MAX_PAIRS = 512;                                       // pair table entries (16 bits each): kernel class | two | first | last | centre tile
// barriers: b1_full b1_empty (NB1 each) | b3_full b3_empty (NB3 <= 4 each) | s_full (2 NSLOT2) p_ready slot_free (NSLOT2 each) | t_full[2] t_free[2] | a_ready stage_full (unused) | ord[2]
constexpr int NB3_MAX = 4;
// s_full: one barrier per (slot, group) = pair index mod 6.  With one barrier per slot (pair mod 3) the waiters of consecutive phases alternated
// between the two epilogue groups: a group waiting for pair g had never observed the phase of pair g - 3 (the other group's, issued by the other
// stage-1 issuer), and a parity wait cannot tell "phase k pending" from "phase k - 1 still pending" -- whenever that issuer ran ~1.5 pairs late the
// wait returned at once and the group converted a stale slot (rare, timing-dependent NaN rows; a delay in an issuer deadlocked: tools/stress_eval.py).
// Barrier g mod 6 is only ever waited on by the group of parity g mod 2, phase after phase.
constexpr int NSF = 2 * NSLOT2;
constexpr int NBAR = 2 * NB1_MAX + 2 * NB3_MAX + NSF + 2 * NSLOT2 + 4 + 3 + 2 + 4;   // + ord[2]: stage-2 issue order hand-off between the two stage-2 issuers; + stat_ready[2] stat_free[2]
static_assert(COL_T2 + 2 * NTMAX == COL_A && COL_A + 2 * A_IMG_COLS == 512, "tensor-memory map");

// debug build only (tools/stress_eval.py): bounded wait that names the waiter before it traps
__device__ __noinline__ void mbar_wait_tag(uint32_t bar, uint32_t parity, long long* dbg, int line) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        ++spins;
        if (spins == (1u << 21)) {                                   // every stuck warp names its wait (stamps 900 + warp: line | parity << 16 | CTA << 32) ...
            if ((threadIdx.x & 31) == 0) dbg[900 + (threadIdx.x >> 5)] = (long long)line | ((long long)parity << 16) | ((long long)blockIdx.x << 32);
            if (atomicCAS((unsigned long long*)(dbg + 960), 0ull, (unsigned long long)line) == 0ull) {
                dbg[961] = blockIdx.x; dbg[962] = threadIdx.x >> 5; dbg[963] = parity; dbg[964] = bar;
            }
            __threadfence_system();
        }
        if (spins > (1u << 24)) { asm volatile("trap;"); }           // ... and the kernel traps once the others have had time to do the same
    }
}

__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo_elem, float hi_elem) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}

// compile-time accumulate flag: the issuing thread runs alone, every extra (uniform-datapath) instruction per MMA costs its
// full latency, so operands are immediates off a per-item base and the k-step count is a template parameter
template <bool ACC>
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "n"(ACC ? 1 : 0) : "memory");
}

// class list of an evaluation class (class-major sub-item stream): kernel class, coefficient columns, first column in the
// column table [k | kx | ky], byte offset of its images inside a tile's coefficient-image record
template <int CLASS> struct ClsList {
    using C = Cfg<CLASS>;
    static constexpr int NCLS = (CLASS == TC_PDE) ? 3 : 2;
    __device__ __forceinline__ static int kern(int ci) { return (CLASS == TC_PDE) ? (ci == 0 ? TK_K : (ci == 1 ? TK_KX : TK_KY)) : (ci == 0 ? TK_K : TK_KY); }
    __device__ __forceinline__ static int ncol(int kc) { return kc == TK_K ? C::NK : (kc == TK_KX ? C::NKX : C::NKY); }
    __device__ __forceinline__ static int coloff(int kc) { return kc == TK_K ? 0 : (kc == TK_KX ? C::NK : C::NK + C::NKX); }
    __device__ __forceinline__ static uint32_t b3off(int kc) { return 2u * 128u * (uint32_t)coloff(kc); }
    __device__ __forceinline__ static uint32_t b1row(int kc) { return kc == TK_K ? 0u : (kc == TK_KX ? 1u : 2u); }   // rows [C | Crollinv | Croll]
};


template <int CLASS, int NSTEP>
__global__ void __launch_bounds__(NTHREADS_P, 1)
eval_tc_kernel(GpView gp, TcDev st, const double* __restrict__ X, long R, int mode,
               double* __restrict__ out0, double* __restrict__ out1, double* __restrict__ out2, double* __restrict__ out3) {
    using C = Cfg<CLASS>;
    using CL = ClsList<CLASS>;
    constexpr int NCLS = CL::NCLS;
    constexpr int KB = (NSTEP + 3) / 4;                              // 64-wide K blocks
    constexpr int NT = C::NK + C::NKX + C::NKY;                      // columns of the column table: [k | kx | ky]
    constexpr int NMAX = C::NK > C::NKY ? C::NK : C::NKY;
    constexpr uint32_t B3_SUB = 2 * NMAX * 128;                      // one sub-item's coefficient images (hi | lo)
    constexpr uint32_t B3_SLOT = 2 * B3_SUB;                         // a pair
    constexpr uint32_t B1_SLOT = KB * 2 * B1_BLK2;                   // a pair's centre rows: per K block [sub-item a rows | sub-item b rows]
    constexpr int NB3 = Ring<CLASS, NSTEP>::NB3, NB1 = Ring<CLASS, NSTEP>::NB1;
    constexpr int RECB = Stage<NSTEP>::REC;                          // bytes of a point's operand record
    constexpr int REC_STAT = NSTEP * 64;                             // offset of (|x|^2, sum x) inside a record, behind its 16 NSTEP words (hi | lo << 16)
    static_assert(RECB == NSTEP * 64 + 16, "record layout (gp_tc.cuh::tc_rec_bytes)");

    extern __shared__ __align__(1024) uint8_t smem_raw[];            // no static smem in this kernel: window offset 0
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0u) { asm volatile("trap;"); }
    uint8_t* sB1 = smem;                                             // NB1 pair slots
    uint8_t* sB3 = sB1 + NB1 * (size_t)B1_SLOT;                      // NB3 pair slots
    uint8_t* sStage = sB3 + NB3 * (size_t)B3_SLOT;                   // the operand records of the next / current point tile (bulk copy -> tensor memory)
    float* feat = (float*)(sStage + (size_t)TM * RECB);              // [128][TF_COUNT] features (floats) of the tile being contracted
    // column table of the contraction, in the form its inner loop wants: per column the byte offsets of its two features in a
    // row's float feature vector (f1 | f2 << 16), its scale 1 / (2^s 2^TC_P_SHIFT) spread over the four outputs as a one-hot float4
    // (the contraction is issue-bound) and the column sum of its coefficients as (hi, lo) floats (baseline term)
    float4* smask = (float4*)(feat + TM * TF_COUNT);                 // [TC_MAXCOL]
    uint32_t* soff = (uint32_t*)(smask + TC_MAXCOL);                 // [TC_MAXCOL]
    float2* scs = (float2*)(soff + TC_MAXCOL);                       // [TC_MAXCOL]
    double* acc4 = (double*)(scs + TC_MAXCOL);                       // [128][4] output sums (u, div, lap, dt) of the tile being contracted, across its classes
    double2* stat2 = (double2*)(acc4 + TM * 4);                      // [2][128] (|x|^2, sum x) of the rows of the current / next point tile, by tile parity
    uint64_t* bars = (uint64_t*)(stat2 + 2 * TM);
    uint32_t* tmem_slot = (uint32_t*)(bars + NBAR);
    uint16_t* ptab = (uint16_t*)(tmem_slot + 4);                     // [npair] pairs of the class-major stream of one point tile

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = gp.D, d = gp.d;
    const long ntiles = (R + TM - 1) / TM;
    const int nit = (int)((ntiles - (long)blockIdx.x + (long)gridDim.x - 1) / (long)gridDim.x);   // point tiles of this CTA
    long long* const dbg = (DBG && st.dbg != nullptr && (int)blockIdx.x == st.dbg_block) ? st.dbg : nullptr;   // stamps: second tile of that CTA
    const int dflags = DBG ? st.dbg_flags : 0;                        // experiment switches (debug build only): 1 no stage-2 MMAs, 2 no epilogue arithmetic
#define TC_STAMP(slot) do { if (DBG && dbg) dbg[(slot)] = clock64(); } while (0)
    if (tid == 0) TC_STAMP(0);
    // debug build: a wait that runs out of probes reports (source line, CTA, warp, parity) through the stamp buffer (pinned host memory survives the trap)
    // watchdog runs (flag 4): every role leaves its progress per CTA behind the stamps (dbg[1024 + 16 CTA + role] = tile << 16 | pair)
#define TC_PROG(role, a, b) do { if (DBG && (dflags & 4)) st.dbg[1024 + 16 * blockIdx.x + (role)] = ((long long)(a) << 16) | (long long)(b); } while (0)
#define MBW(bar, par) do { if (DBG && (dflags & 4)) mbar_wait_tag((bar), (par), st.dbg, __LINE__); else mbar_wait((bar), (par)); } while (0)
    const uint32_t bar0 = smem_u32(bars);
    auto b1_full = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    auto b1_empty = [&](int i) { return bar0 + 8u * (uint32_t)(NB1_MAX + i); };
    auto b3_full = [&](int i) { return bar0 + 8u * (uint32_t)(2 * NB1_MAX + i); };
    auto b3_empty = [&](int i) { return bar0 + 8u * (uint32_t)(2 * NB1_MAX + NB3_MAX + i); };
    auto s_full = [&](int i) { return bar0 + 8u * (uint32_t)(2 * NB1_MAX + 2 * NB3_MAX + i); };
    auto p_ready = [&](int i) { return bar0 + 8u * (uint32_t)(2 * NB1_MAX + 2 * NB3_MAX + NSF + i); };
    auto slot_free = [&](int i) { return bar0 + 8u * (uint32_t)(2 * NB1_MAX + 2 * NB3_MAX + NSF + NSLOT2 + i); };
    const uint32_t t_full0 = bar0 + 8u * (uint32_t)(2 * NB1_MAX + 2 * NB3_MAX + NSF + 2 * NSLOT2), t_free0 = t_full0 + 16u, a_ready = t_full0 + 32u,
                   stage_full = t_full0 + 40u, ord0 = t_full0 + 56u, stat_ready0 = t_full0 + 72u, stat_free0 = t_full0 + 88u;
    // T accumulators: one per class epoch e = tile * NCLS + class, double-buffered by epoch parity.  One barrier pair per buffer: a
    // waiter is never more than one phase behind.
    auto t_full = [&](int e) { return t_full0 + 8u * (uint32_t)(e & 1); };
    auto t_free = [&](int e) { return t_free0 + 8u * (uint32_t)(e & 1); };

    if (tid == 0) {
        for (int i = 0; i < NB1; ++i) { mbar_init(b1_full(i), 1); mbar_init(b1_empty(i), 1); }
        for (int i = 0; i < NB3; ++i) { mbar_init(b3_full(i), 1); mbar_init(b3_empty(i), 1); }
        for (int i = 0; i < NSF; ++i) mbar_init(s_full(i), 1);
        for (int i = 0; i < NSLOT2; ++i) { mbar_init(p_ready(i), NEPI / 2); mbar_init(slot_free(i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(t_full0 + 8u * i, 2); mbar_init(t_free0 + 8u * i, NLOAD); }
        mbar_init(a_ready, NEPI); mbar_init(stage_full, 1);
        mbar_init(ord0, 1); mbar_init(ord0 + 8u, 1);
        mbar_init(stat_ready0, 4); mbar_init(stat_ready0 + 8u, 4);
        mbar_init(stat_free0, NLOAD); mbar_init(stat_free0 + 8u, NLOAD);
        fence_barrier_init();
    }
    if (warp == W_S1A) tmem_alloc(smem_u32(tmem_slot), 512);
    for (int c = tid; c < TC_MAXCOL; c += NTHREADS_P) {              // column table -> shared memory (read by every contraction)
        const TcColDesc dsc = st.desc[c < NT ? c : 0];
        const bool dead = (c >= NT || dsc.out == TO_PAD);
        const float sc = dead ? 0.0f : (float)dsc.inv_scale;        // a power of two: exact in FP32 (range checked when the images are built)
        soff[c] = dead ? 0u : ((uint32_t)dsc.f1 * 4u) | (((uint32_t)dsc.f2 * 4u) << 16);
        smask[c] = make_float4(dsc.out == TO_U ? sc : 0.0f, dsc.out == TO_G ? sc : 0.0f, dsc.out == TO_L ? sc : 0.0f, dsc.out == TO_T ? sc : 0.0f);
        const double cs = dead ? 0.0 : st.csum[c];
        const float csh = (float)cs;
        scs[c] = make_float2(csh, (float)(cs - (double)csh));
    }
    tc_fence_before();
    __syncthreads();                                                 // TMEM base address + barriers + column table visible
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const int nt_all = st.ntile_all, nt_dom = st.ntile_dom;          // 48-centre tiles of the k / kx classes and of the ky class
    auto ntile_of = [&](int ci) { return CL::kern(ci) == TK_KY ? nt_dom : nt_all; };
    int npair = 0;
    for (int ci = 0; ci < NCLS; ++ci) npair += (ntile_of(ci) + 1) >> 1;   // pairs per point tile (an odd class ends in a single)
    // pair table (every role walks the same stream; the issuing threads run alone, so their bookkeeping is one shared-memory load):
    // bits 0-1 kernel class, 2 two sub-items, 3 first pair of its class, 4 last pair of its class, 5.. centre tile of the first sub-item
    if (tid == 0) {
        int n = 0;
        for (int ci = 0; ci < NCLS; ++ci) {
            const int nt = ntile_of(ci);
            for (int t = 0; t < nt; t += 2)
                ptab[n++] = (uint16_t)(CL::kern(ci) | ((t + 1 < nt) ? 4 : 0) | (t == 0 ? 8 : 0) | ((t + 2 >= nt) ? 16 : 0) | (t << 5));
        }
    }
    __syncthreads();
#define PT_KERN(e) ((int)((e) & 3u))
#define PT_TWO(e) (((e) & 4u) != 0u)
#define PT_FIRST(e) (((e) & 8u) != 0u)
#define PT_LAST(e) (((e) & 16u) != 0u)
#define PT_TILE(e) ((int)((e) >> 5))

    if (warp == W_PROD) {
        // ===== producer: per pair, its sub-items' centre rows (all K blocks) and coefficient images.  The pair stream simply
        // repeats for every point tile, so the rings run across tile boundaries.  A centre-row slot is reusable once stage 1 of its
        // previous pair completed (b1_empty), a coefficient slot once stage 2 of its previous pair completed (b3_empty). =====
        if (lane == 0) {
            const int ptotal = nit * npair;
            int p1 = 0, j1 = 0, j3 = 0;
            auto load_b1 = [&]() {
                const int s = p1 % NB1;
                if (p1 >= NB1) MBW(b1_empty(s), (uint32_t)((p1 / NB1) - 1) & 1u);
                const uint32_t e = ptab[j1];
                const int kc = PT_KERN(e), t0 = PT_TILE(e);
                const bool two = PT_TWO(e);
                if (++j1 == npair) j1 = 0;
                // one bulk copy per pair: the image holds a pair's slot layout [K block][sub-item][48 rows x 128 B] contiguously (a single's second
                // half is zero padding in the image)
                (void)two;
                mbar_expect_tx(b1_full(s), B1_SLOT);
                bulk_g2s(smem_u32(sB1 + (size_t)s * B1_SLOT), st.b1 + ((size_t)CL::b1row(kc) * st.npair_all + (size_t)(t0 >> 1)) * B1_SLOT, B1_SLOT, b1_full(s));
                ++p1;
            };
            for (int p = 0; p < ptotal; ++p) {
                TC_PROG(0, p1, p);
                while (p1 < ptotal && p1 < p + NB1) load_b1();               // centre rows run NB1 - 1 pairs ahead of the coefficient images
                const int q = p % NB3;
                if (p >= NB3) MBW(b3_empty(q), (uint32_t)((p / NB3) - 1) & 1u);
                const uint32_t e = ptab[j3];
                const int kc = PT_KERN(e), t0 = PT_TILE(e);
                const bool two = PT_TWO(e);
                if (++j3 == npair) j3 = 0;
                // one bulk copy per pair: [sub-item][hi | lo][ncol rows x 128 B]
                (void)two;
                const uint32_t bytes = 4u * 128u * (uint32_t)CL::ncol(kc);
                mbar_expect_tx(b3_full(q), bytes);
                bulk_g2s(smem_u32(sB3 + (size_t)q * B3_SLOT), st.b3 + (size_t)st.npair_all * 2 * CL::b3off(kc) + (size_t)(t0 >> 1) * bytes, bytes, b3_full(q));
            }
        }
        __syncwarp();
    } else if (warp == W_S1A || warp == W_S1B) {
        // ===== stage-1 issuers (distance GEMMs), one for the even and one for the odd pairs.  tcgen05.mma is issued by one thread, the
        // MMA queue is only a few entries deep and every barrier operation of a lone thread costs 100 - 400 cycles: with one issuer per
        // stage each needed ~1 500 cycles per pair (waits 560, issue 530, commits 360: tools/tc_timeline.py) and the tensor pipe
        // (~1 200 cycles of work per pair) ran dry while both were between batches.  With two issuers per stage one of them is issuing
        // while the other one waits / commits.  Pairs of different parity write different S slots: no order between the two is needed.
        // The warp stays converged and one elected lane issues. =====
        const int w = (warp == W_S1A) ? 0 : 1;
        const uint32_t el = elect_one();
        const uint64_t b1desc0 = make_desc(smem_u32(sB1), 1, 64, 2);
        const uint32_t idS2 = make_idesc(TM, 2 * TN2), idS1 = make_idesc(TM, TN2);
        const uint32_t aBase = tmem_base + COL_A;
        int p = 0;                                                   // global pair counter
        // The operand records of a point tile are one contiguous piece of global memory: issuer 0 moves them into the staging buffer with ONE bulk
        // copy per tile -- tile 0 at once, tile it + 1 as soon as tile `it` has left the buffer (a_ready: its images are in tensor memory and every
        // warp has read its rows' scalars), i.e. a whole main loop before it is needed.  (Staging used to be computed here from the FP64 rows, at the
        // tile boundary: ~11 k cycles of load latency and FP64 conversions per tile with the tensor pipe idle, tools/tc_timeline.py.)
        auto copy_records = [&](int ts) {
            const long row0 = ((long)blockIdx.x + (long)ts * gridDim.x) * TM;
            const uint32_t nrow = (uint32_t)((R - row0 < TM) ? (R - row0) : TM);    // a partial tile leaves stale rows behind: they only reach their own (unwritten) outputs
            mbar_expect_tx(stage_full, nrow * (uint32_t)RECB);
            bulk_g2s(smem_u32(sStage), st.rec + (size_t)row0 * RECB, nrow * (uint32_t)RECB, stage_full);
        };
        if (w == 0 && el) copy_records(0);
        for (int it = 0; it < nit; ++it) {
            const bool stamp = (it == 1);
            if (el && it == 1 && w == 0) TC_STAMP(244);
            MBW(a_ready, (uint32_t)it & 1u);                   // A images of this point tile are in tensor memory
            tc_fence_after();
            if (w == 0 && el && it + 1 < nit) copy_records(it + 1);
            for (int j = 0; j < npair; ++j, ++p) {
                if ((p & 1) != w) continue;
                if (el) TC_PROG(2 + w, it, j);
                const int s = p % NSLOT2, s1 = p % NB1;
                const uint32_t accS = tmem_base + (uint32_t)s * SLOTC;
                const uint32_t idesc = PT_TWO(ptab[j]) ? idS2 : idS1;
                if (el && stamp && j < 60) TC_STAMP(4 + 4 * j);
                // the centre rows first: they landed long ago (the ring runs pairs ahead), and every barrier operation of this lone thread costs
                // 100-200 cycles -- taken before the wait that IS on the pair's critical path instead of after it
                MBW(b1_full(s1), (uint32_t)(p / NB1) & 1u);    // bulk-copy bytes landed (async proxy)
                if (el && stamp && j < 60) TC_STAMP(257 + 8 * j);
                // the slot's previous P has been consumed: stage 2 of pair p - NSLOT2 has COMPLETED
                if (p >= NSLOT2) { MBW(slot_free(s), (uint32_t)((p / NSLOT2) - 1) & 1u); tc_fence_after(); }
                if (el && stamp && j < 60) TC_STAMP(256 + 8 * j);
                const uint64_t bb = b1desc0 + (uint64_t)(((uint32_t)s1 * B1_SLOT) >> 4);
                if (el) {
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                        for (int half = 1; half >= 0; --half) {      // low halves first (tiny terms), then the high halves
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                if (kb * 4 + ks < NSTEP) {
                                    const uint32_t aa = aBase + (uint32_t)half * A_IMG_COLS + (uint32_t)(kb * 4 + ks) * 8u;
                                    const uint64_t bd = bb + (uint64_t)((kb * 2 * B1_BLK2 + ks * 32) >> 4);
                                    if (kb == 0 && half == 1 && ks == 0) umma_ts<false>(accS, aa, bd, idesc);
                                    else umma_ts<true>(accS, aa, bd, idesc);
                                }
                            }
                        }
                    }
                }
                if (el && stamp && j < 60) TC_STAMP(258 + 8 * j);
                if (el) { umma_commit(b1_empty(s1)); umma_commit(s_full(p % NSF)); }   // centre-row slot reusable; accumulators ready
                if (el && stamp && j < 60) TC_STAMP(259 + 8 * j);
                __syncwarp();
            }
        }
    } else if (warp == W_S2A || warp == W_S2B) {
        // ===== stage-2 issuers: T += (P - 2^s) C for the sub-items of a pair; per 16 centres the three split products hh + hl + lh.
        // Even / odd pairs alternate between the two issuers; the accumulation into T keeps the stream order (deterministic results,
        // and the class's first MMA, which overwrites T, stays first): an issuer starts issuing pair p only after the other one has
        // issued pair p - 1 (ord barriers; tcgen05.mma executes in issue order). =====
        const int w = (warp == W_S2A) ? 0 : 1;
        const uint32_t el = elect_one();
        const uint64_t b3desc0 = make_desc(smem_u32(sB3), 1, 64, 2);
        const uint32_t ord_mine = ord0 + 8u * (uint32_t)w, ord_other = ord0 + 8u * (uint32_t)(1 - w);
        int p = 0, epoch = 0;                                        // global pair counter; class epochs started
        for (int it = 0; it < nit; ++it) {
            const bool stamp = (it == 1);
            for (int j = 0; j < npair; ++j, ++p) {
                const uint32_t e = ptab[j];
                if ((p & 1) == w) {
                    if (el) TC_PROG(4 + w, it, j);
                    const int s = p % NSLOT2, q = p % NB3;
                    const int kc = PT_KERN(e);
                    const bool two = PT_TWO(e), first = PT_FIRST(e);
                    // coefficient images landed (long ago) and, for the first pair of a class, its T buffer contracted (class epoch - 2): both
                    // before the wait on the critical path -- P written over S by the pair's epilogue group
                    MBW(b3_full(q), (uint32_t)(p / NB3) & 1u);
                    if (el && stamp && j < 60) TC_STAMP(261 + 8 * j);
                    if (first && epoch >= 2) MBW(t_free(epoch), (uint32_t)((epoch >> 1) - 1) & 1u);
                    MBW(p_ready(s), (uint32_t)(p / NSLOT2) & 1u);
                    if (el && stamp && j < 60) TC_STAMP(260 + 8 * j);
                    // the other issuer has issued pair p - 1
                    if (p >= 1) MBW(ord_mine, (uint32_t)(w == 1 ? (p >> 1) : ((p >> 1) - 1)) & 1u);
                    tc_fence_after();
                    const uint32_t ncol = (uint32_t)CL::ncol(kc);
                    const uint32_t idesc = make_idesc(TM, (int)ncol);
                    const uint32_t tacc = tmem_base + COL_T2 + (uint32_t)(epoch & 1) * NTMAX;
                    if (el && !(dflags & 1)) {
                        // timeline experiments (debug build, wrong results): 32 = hh products only; 128 = hl / lh products into the other T buffer
                        // (two accumulation chains instead of one)
                        const bool only_hh = (dflags & 32) != 0;
                        const uint32_t tacc2 = (dflags & 128) ? (tmem_base + COL_T2 + (uint32_t)((epoch + 1) & 1) * NTMAX) : tacc;
#pragma unroll 1
                        for (int sub = 0; sub < (two ? 2 : 1); ++sub) {
                            const uint32_t pbase = tmem_base + (uint32_t)s * SLOTC + (uint32_t)sub * TN2;
                            const uint64_t b3 = b3desc0 + (uint64_t)(((uint32_t)q * B3_SLOT + (uint32_t)sub * (2u * ncol * 128u)) >> 4);
                            const uint64_t clo0 = b3 + (uint64_t)((ncol * 128u) >> 4);
                            if (first && sub == 0) umma_ts<false>(tacc, pbase, b3, idesc); else umma_ts<true>(tacc, pbase, b3, idesc);
                            if (!only_hh) {
                                umma_ts<true>(tacc2, pbase, clo0, idesc);
                                umma_ts<true>(tacc2, pbase + 8u, b3, idesc);
                            }
#pragma unroll
                            for (int ks = 1; ks < TN2 / 16; ++ks) {
                                umma_ts<true>(tacc, pbase + (uint32_t)ks * 16u, b3 + (uint64_t)((ks * 32) >> 4), idesc);
                                if (!only_hh) {
                                    umma_ts<true>(tacc2, pbase + (uint32_t)ks * 16u, clo0 + (uint64_t)((ks * 32) >> 4), idesc);
                                    umma_ts<true>(tacc2, pbase + (uint32_t)ks * 16u + 8u, b3 + (uint64_t)((ks * 32) >> 4), idesc);
                                }
                            }
                        }
                    }
                    if (el && stamp && j < 60) TC_STAMP(262 + 8 * j);
                    if (el) {
                        // the other issuer may issue pair p + 1.  The fence orders this thread's MMAs before the hand-off: without it the other
                        // issuer's first MMAs could overtake this pair's last ones on their way to the tensor pipe (results then differed from run
                        // to run in the last bit of a few rows -- a different accumulation order: tools/stress_eval.py)
                        if (dflags & 8) { tc_fence_before(); mbar_arrive(ord_other); } else umma_commit(ord_other);
                        umma_commit(b3_empty(q));                     // coefficient slot reusable (producer)
                        umma_commit(slot_free(s));                    // P consumed (stage-1 issuers)
                        // The class's T is complete (contraction warps) when the MMAs of BOTH issuers have completed: a commit only covers the
                        // MMAs of the committing thread, so t_full takes two -- from the issuer of the class's last pair and from the other one
                        // after ITS last pair of the class (the one before).  With the last pair's commit alone the contraction now and then read a
                        // T that still lacked the other issuer's final products in some rows (last-bit differences from run to run).
                        if (PT_LAST(e)) {
                            umma_commit(t_full(epoch));
                            if (PT_FIRST(e)) umma_commit(t_full(epoch));          // a class of one pair
                        } else if (j + 1 < npair && PT_LAST(ptab[j + 1]) && !PT_FIRST(ptab[j + 1])) {
                            umma_commit(t_full(epoch));
                        }
                        if (stamp && j < 60) TC_STAMP(5 + 4 * j);
                    }
                    __syncwarp();
                }
                if (PT_LAST(e)) ++epoch;
            }
            if (el && it == 0 && w == 0) TC_STAMP(248);
        }
    } else if (warp >= W_LOAD0 && warp < W_LOAD0 + NLOAD) {
        // ===== contraction warps (one per tensor-memory lane quadrant, thread <-> point row).  Per point tile: the row's scalars (K_i with the
        // exponent shift divided out, row sum) from the |x|^2 / sum x the epilogue warps left in shared memory when they staged the tile, then one
        // contraction per finished class:  out_o += K_i sum_col F[f1] F[f2] (T[col] + 2^s csum[col]) inv_scale[col].
        // The staging itself used to run on these four warps in the background; a warp issues ~0.25 instructions per cycle whatever it does, and their
        // ~10 k instructions per tile took longer than the main loop (65 k vs 55 k cycles in the PDE class, 50 k vs 36 k in the u class:
        // tools/tc_timeline.py), so the tile period was theirs.  Sixteen warps do it at the tile boundary in ~3 k cycles. =====
        const int qd = warp & 3;                                     // tensor-memory lane quadrant = row group of this warp
        const double ymax2 = __ldg(st.ymax2);
        const int rq = qd * 32 + lane;                               // this thread's row
        for (int tp = 0; tp < nit; ++tp) {
            // |x|^2 and the row sum of this thread's row: the epilogue warps left them in stat2 before they released the staging buffer.  (These
            // warps do not take part in a_ready themselves: the stage-1 issuers would wait for the previous tile's last contraction, ~2 k cycles.)
            // stat2 is double-buffered by tile parity, with a full / free barrier pair per buffer: the epilogue warps run up to three pairs ahead of
            // stage 2 -- with one pair per class (a GP of < 49 collocation points) that is two point tiles ahead of these warps -- so they wait for
            // stat_free before they overwrite a buffer, and nobody is ever two phases behind on a parity wait.  (Waiting on a_ready here, and a
            // first version without stat_free, deadlocked exactly those small GPs: tests/test_gpu_tc.py::test_pipeline_protocol_under_many_tiles.)
            MBW(stat_ready0 + 8u * (uint32_t)(tp & 1), (uint32_t)(tp >> 1) & 1u);
            double ki_prev, sx_prev;
            float sc_prev;
            {
                const double2 stat = stat2[(tp & 1) * TM + rq];
                __syncwarp();
                if (lane == 0) mbar_arrive(stat_free0 + 8u * (uint32_t)(tp & 1));
                const int sh = row_shift(gp.a, stat.x, ymax2);
                sc_prev = exp2f((float)sh);
                ki_prev = ldexp(exp(-0.5 * gp.a * stat.x), TC_P_SHIFT - sh); sx_prev = stat.y;
            }
            if (tp == 1 && qd == 0 && lane == 0) TC_STAMP(1);
            {   // features of the row, once per tile, BEFORE the wait for the first class's T: the loads' latency used to sit between t_full and
                // t_free of that class, i.e. on the path of the coefficient GEMMs two class epochs later (with the 12 scattered loads from the FP64
                // row that was worth 2 ms per C3 step).  Floats: FP64 conversions run at 1/8 rate here.  The previous tile's contraction is done.
                const long row = ((long)blockIdx.x + (long)tp * gridDim.x) * TM + rq;
                const bool ok = row < R;
                float* F = feat + rq * TF_COUNT;                     // private to this thread
                // the row's 12 feature coordinates: one contiguous 96-byte block, written with the operand record
                double fb[TC_REC_NFEAT];
                {
                    const double2* fp = (const double2*)(st.recf + (ok ? row : 0) * TC_REC_NFEAT);
#pragma unroll
                    for (int k = 0; k < TC_REC_NFEAT / 2; ++k) { const double2 v = __ldg(fp + k); fb[2 * k] = ok ? v.x : 0.0; fb[2 * k + 1] = ok ? v.y : 0.0; }
                }
                const double x0 = fb[0], xt = fb[1];
                double P2 = 0.0, R2 = 0.0;
#pragma unroll
                for (int m = 0; m < MC_IDX; ++m) {
                    const double xi = fb[2 + m], xir = fb[2 + MC_IDX + m];
                    F[TF_XI + m] = (float)xi; F[TF_XR + m] = (float)xir;
                    P2 = fma(xi, xi, P2); R2 = fma(xir, xir, R2);
                }
                F[TF_ONE] = 1.0f; F[TF_SX] = (float)sx_prev; F[TF_XT] = (float)xt; F[TF_X0] = (float)x0; F[TF_SXR] = (float)(sx_prev - x0 + xt);
                F[TF_P2] = (float)P2; F[TF_R2] = (float)R2;
                acc4[rq * 4 + 0] = 0.0; acc4[rq * 4 + 1] = 0.0; acc4[rq * 4 + 2] = 0.0; acc4[rq * 4 + 3] = 0.0;
            }
            for (int ce = 0; ce < NCLS; ++ce) {
                const int e = tp * NCLS + ce;
                if (qd == 0 && lane == 0) TC_PROG(6, tp, ce);
                MBW(t_full(e), (uint32_t)(e >> 1) & 1u);
                {
                    tc_fence_after();
                    const long row = ((long)blockIdx.x + (long)tp * gridDim.x) * TM + rq;
                    const bool ok = row < R;
                    float* F = feat + rq * TF_COUNT;                 // private to this thread
                    if (tp == 0 && qd == 0 && lane == 0) TC_STAMP(246);
                    const int kc = CL::kern(ce);
                    const int ncol = CL::ncol(kc), coff = CL::coloff(kc);
                    const uint32_t tbase = tmem_base + ((uint32_t)(qd * 32) << 16) + COL_T2 + (uint32_t)(e & 1) * NTMAX;
                    double aU = 0.0, aG = 0.0, aL = 0.0, aT = 0.0;
                    constexpr int CH = 16;                           // T columns per tensor-memory load round
                    const char* Fb = (const char*)F;
#pragma unroll 1
                    for (int cb = 0; cb < ncol; cb += CH) {
                        float tv[CH];
                        tmem_ld16(tbase + (uint32_t)cb, tv);
                        tmem_ld_wait();
                        if (cb + CH >= ncol) {                       // last round: this T buffer may be overwritten (two class epochs later)
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(t_free(e));
                        }
                        // FP32 products and FP32 partial sums per output over the round (one-hot scale masks: no output bookkeeping), FP64 across rounds
                        float pU = 0.0f, pG = 0.0f, pL = 0.0f, pT = 0.0f;
#pragma unroll
                        for (int i = 0; i < CH; ++i) {
                            const uint32_t o = soff[coff + cb + i];
                            const float4 m = smask[coff + cb + i];
                            const float2 cs = scs[coff + cb + i];
                            const float w = fmaf(sc_prev, cs.x, fmaf(sc_prev, cs.y, tv[i]));      // sum_j P C = T + 2^s sum_j C (baseline term)
                            const float t = (*(const float*)(Fb + (o & 0xffffu)) * *(const float*)(Fb + (o >> 16))) * w;
                            pU = fmaf(t, m.x, pU); pG = fmaf(t, m.y, pG); pL = fmaf(t, m.z, pL); pT = fmaf(t, m.w, pT);
                        }
                        aU += (double)pU; aG += (double)pG; aL += (double)pL; aT += (double)pT;
                    }
                    if (ce + 1 < NCLS) {
                        acc4[rq * 4 + 0] += aU; acc4[rq * 4 + 1] += aG; acc4[rq * 4 + 2] += aL; acc4[rq * 4 + 3] += aT;
                    } else {
                        aU += acc4[rq * 4 + 0]; aG += acc4[rq * 4 + 1]; aL += acc4[rq * 4 + 2]; aT += acc4[rq * 4 + 3];
                        if (ok) {
                            const double ki = ki_prev;
                            const double u = ki * aU;
                            if (CLASS == TC_U) {
                                const double gt = 1.0 - 1.0 / (1.0 + exp(sx_prev + __ldg(X + row * (long)D + d)));   // equations.py:259
                                out0[row] = (mode == EVAL_TERMINAL) ? gt - u : u;
                            } else if (CLASS == TC_UG) {
                                out0[row] = u;
                                out1[row] = ki * aG;
                            } else {
                                const double gg = ki * aG, l = ki * aL, tt = ki * aT;
                                const double s2 = gp.sig2;
                                out0[row] = tt + (s2 * u - 1.0 / (double)d - 0.5 * s2) * gg + 0.5 * s2 * l;   // GP.py:767-768
                                if (out1) out1[row] = gg;
                                if (out2) out2[row] = l;
                                if (out3) out3[row] = tt;
                            }
                        }
                        if (tp == 0 && qd == 0 && lane == 0) TC_STAMP(247);
                    }
                }
            }
        }
    } else {
        // ===== epilogue warps: group = pair parity; warp <-> (lane quadrant, sub-item of the pair); thread <-> point row x 48 centres =====
        const int ew = warp - W_EPI0;                                // epilogue warp index 0..15
        const int r = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const bool skip_math = (dflags & 2) != 0;
        const double ymax2 = __ldg(st.ymax2);
        const int grp = ew >> 3;
        const int sub = (ew >> 2) & 1;
        long g = 0;                                                  // global pair counter
        for (int it = 0; it < nit; ++it) {
            // --- the operand records of tile `it` were bulk-copied into the staging buffer during the previous main loop
            MBW(stage_full, (uint32_t)it & 1u);
            if (it == 1 && ew == 0 && lane == 0) TC_STAMP(1);
            const uint8_t* rowp = sStage + (size_t)r * RECB;
            const double2 stat_r = *(const double2*)(rowp + REC_STAT);   // (|x|^2, sum x) of this thread's point row
            const double nx_r = stat_r.x;
            if (ew < 4) {                                            // for the contraction warps (one epilogue warp per lane quadrant)
                if (it >= 2) MBW(stat_free0 + 8u * (uint32_t)(it & 1), (uint32_t)((it >> 1) - 1) & 1u);   // they have read tile it - 2
                stat2[(it & 1) * TM + r] = stat_r;
                __syncwarp();
                if (lane == 0) mbar_arrive(stat_ready0 + 8u * (uint32_t)(it & 1));
            }
            // --- A images of tile `it`: staging buffer -> tensor memory (lane = row).  All stage-1 MMAs of the previous tile have completed (the
            // copying warps saw s_full of its last pairs at the end of their loop), so the images can be overwritten.
            if (ew < 8) {                                            // image = ew >> 2 (hi | lo); 32-bit column c = K elements 2c, 2c + 1
                const int img = ew >> 2;
                const uint32_t sel = img ? 0x7632u : 0x5410u;        // a record word is (hi | lo << 16) of one K element: pick this image's halves
                const uint32_t taddr = tmem_base + COL_A + (uint32_t)img * A_IMG_COLS + lane_addr;
#pragma unroll
                for (int ks = 0; ks < NSTEP; ++ks) {
                    uint32_t wv[8];
#pragma unroll
                    for (int h4 = 0; h4 < 4; ++h4) {
                        const uint4 q = *(const uint4*)(rowp + ks * 64 + h4 * 16);
                        wv[2 * h4] = __byte_perm(q.x, q.y, sel);
                        wv[2 * h4 + 1] = __byte_perm(q.z, q.w, sel);
                    }
                    tmem_st8(taddr + (uint32_t)ks * 8u, wv);
                }
                tmem_st_wait();
                tc_fence_before();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);                     // images in tensor memory / this warp is done with the staging buffer
            if (it == 1 && ew == 0 && lane == 0) TC_STAMP(2);
            // 2^s of this thread's point row: P - 2^s = 2^s ex2(S) - 2^s
            const float sc = exp2f((float)row_shift(gp.a, nx_r, ymax2)), nsc = -sc;
            const bool stamp = (it == 1);
            // --- main loop: S -> P - 2^s in place.  The fixed cost of a round trip (barrier wake-up, first tcgen05.ld, tcgen05.wait::st,
            // fence, arrive) is paid once per 48 columns per thread, with the next chunk's tcgen05.ld in flight during the current
            // chunk's ex2 work.  The two groups run half a period apart.
            for (int j = 0; j < npair; ++j, ++g) {
                if ((int)(g & 1) != grp) continue;                   // pair ownership alternates between the two groups
                if ((ew & 7) == 0 && lane == 0) TC_PROG(7 + grp, it, j);
                const bool two = PT_TWO(ptab[j]);
                const int s = (int)(g % NSLOT2);
                MBW(s_full((int)(g % NSF)), (uint32_t)(g / NSF) & 1u);
                tc_fence_after();
                if ((ew & 7) == 0 && lane == 0 && stamp && j < 60) TC_STAMP(6 + 4 * j);
                if (!skip_math && (sub == 0 || two)) {
                    const uint32_t base = tmem_base + lane_addr + (uint32_t)s * SLOTC + (uint32_t)sub * TN2;
                    constexpr int nch = TN2 / 16;                    // chunks of 16 columns
                    float v[2][16];
                    tmem_ld16(base, v[0]);
#pragma unroll
                    for (int ch = 0; ch < nch; ++ch) {
                        tmem_ld_wait();                              // chunk ch has arrived
                        if (ch + 1 < nch) tmem_ld16(base + (uint32_t)(ch + 1) * 16u, v[(ch + 1) & 1]);
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float p0 = fmaf(ex2f(v[ch & 1][2 * i]), sc, nsc), p1 = fmaf(ex2f(v[ch & 1][2 * i + 1]), sc, nsc);   // P - 2^s
                            hi[i] = pack_f16x2_sat(p0, p1);
                            const float2 hf = __half22float2(*(const __half2*)&hi[i]);
                            lo[i] = pack_f16x2_sat(p0 - hf.x, p1 - hf.y);
                        }
                        // P over S in place: 16 FP32 columns -> [8 columns of hi pairs | 8 columns of lo pairs] (the A operand of stage 2)
                        const uint32_t cb = base + (uint32_t)ch * 16u;
                        tmem_st8(cb, hi);
                        tmem_st8(cb + 8u, lo);
                    }
                    tmem_st_wait();
                }
                tc_fence_before();
                __syncwarp();
                if ((ew & 7) == 0 && lane == 0 && stamp && j < 60) TC_STAMP(7 + 4 * j);
                if (lane == 0) mbar_arrive(p_ready(s));
            }
            // the A images of the next tile overwrite this tile's: every stage-1 MMA of the tile must have completed, also the last pair's,
            // which may belong to the other group -- the copying warps wait for the final pair's accumulators explicitly
            // (two stage-1 issuers: the last pair of each of them)
            if (ew < 8) {
                for (long gl = (g >= 2 ? g - 2 : 0); gl < g; ++gl)   // the tile's last two pairs (g was advanced past them)
                    MBW(s_full((int)(gl % NSF)), (uint32_t)(gl / NSF) & 1u);
                tc_fence_after();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) TC_STAMP(3);
    if (warp == W_S1A) tmem_dealloc(tmem_base, 512);
#undef TC_STAMP
#undef MBW
#undef TC_PROG
#undef PT_KERN
#undef PT_TWO
#undef PT_FIRST
#undef PT_LAST
#undef PT_TILE
}

// Operand records of a caller's points (public evaluation API, tools): one warp per row, lane <-> column (coalesced 256-byte row segments).
// The very same arithmetic and summation tree as the Picard samplers' emission (picard.cu::rec_emit / rec_reduce2), so a point's record does
// not depend on who wrote it: per-lane FMA / add chains over c = lane, lane + 32, ..., then the xor-16-8-4-2-1 butterfly.
__global__ void __launch_bounds__(256) rec_image_kernel(const double* __restrict__ X, long R, int D, int d, int nstep, float ascale_f,
                                                        TcRecIdx fi, uint8_t* __restrict__ rec, double* __restrict__ recf) {
    const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= R) return;
    const int NC = nstep * 16, recb = nstep * 64 + 16;
    uint32_t* rw = (uint32_t*)(rec + (size_t)row * recb);
    const double* xr = X + row * (long)D;
    double nx = 0.0, sx = 0.0;
    for (int c = lane; c < ((NC + 31) & ~31); c += 32) {
        const double val = (c < D) ? __ldg(xr + c) : 0.0;
        nx = fma(val, val, nx);
        sx += val;                                                   // all D columns; the time column is taken out below
        const float sv = (float)val * ascale_f;
        // hi + lo split in FP32 (sv rounded to 24 bits; sv - hi is exact in FP32): |error| <= 2^-22 |sv|
        const float hf = __half2float(__float2half_rn(sv));
        uint32_t w;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(sv - hf), "f"(hf));      // upper half: lo, lower half: hi
        if (c < NC) rw[c] = w;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) { nx += __shfl_xor_sync(0xffffffffu, nx, o); sx += __shfl_xor_sync(0xffffffffu, sx, o); }
    if (lane == 0) { double* stp = (double*)(rw + NC); stp[0] = nx; stp[1] = sx - __ldg(xr + d); }
    if (lane < TC_REC_NFEAT) recf[row * TC_REC_NFEAT + lane] = __ldg(xr + fi.col[lane]);
}

template <int CLASS, int NSTEP>
static size_t smem_bytes() {
    using C = Cfg<CLASS>;
    constexpr size_t NMAX = C::NK > C::NKY ? C::NK : C::NKY;
    return Ring<CLASS, NSTEP>::NB1 * (size_t)(((NSTEP + 3) / 4) * 2 * B1_BLK2) + Ring<CLASS, NSTEP>::NB3 * (2 * 2 * NMAX * 128) + (size_t)TM * Stage<NSTEP>::REC
           + (size_t)TM * TF_COUNT * 4 + TC_MAXCOL * (sizeof(float4) + sizeof(uint32_t) + sizeof(float2)) + (size_t)TM * 4 * 8
           + 2 * TM * sizeof(double2) + NBAR * 8 + 16 + MAX_PAIRS * sizeof(uint16_t);
}

template <int CLASS, int NSTEP>
static int launch(const GpView& gp, const TcDev& st, const double* X, long R, int mode,
                  double* o0, double* o1, double* o2, double* o3, cudaStream_t stream) {
    // per-device attributes: set / queried on every launch (microseconds) so that a process may use several devices
    const size_t smem = smem_bytes<CLASS, NSTEP>();
    int dev = 0, nsm = 0;
    SC_CUDA(cudaGetDevice(&dev));
    SC_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    SC_CUDA(cudaFuncSetAttribute(eval_tc_kernel<CLASS, NSTEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SC_REQUIRE(3 * ((st.ntile_all + 1) / 2) <= MAX_PAIRS, "tcgen05 route: too many collocation points for the pair table (use the FP64 route)");
    const long ntiles = cdiv(R, TM);
    const unsigned grid = (unsigned)(ntiles < nsm ? ntiles : nsm);   // persistent: one CTA per SM
    eval_tc_kernel<CLASS, NSTEP><<<grid, NTHREADS_P, smem, stream>>>(gp, st, X, R, mode, o0, o1, o2, o3);
    SC_LAUNCH_CHECK();
    return OK;
}

// ---- K-streamed variant for d + 2 > 128 (up to 1024 contraction columns) --------------------------------------------------
// The point images no longer fit tensor memory (or shared memory), so a pre-pass (`ks_image_kernel`) writes them to a global
// scratch buffer in the MMA's shared-memory layout (per 128-point tile, per 64-wide K block: [hi 16 KB | lo 16 KB], f16,
// 128 B swizzle) together with K_i and the row sums, and the evaluation kernel streams A AND B K-blocks through a 3-stage
// ring: 8 SS-mode N = 128 MMAs per stage (4 k-steps x hi/lo), one barrier wait + one commit per stage.  Everything after
// stage 1 (in-place ex2 / split, coefficient GEMM from tensor memory, FP64 final contraction) is the same as above; with the
// A images out of tensor memory the layout is two 128-column S/P slots + T.  At d = 1000 the distance GEMM dominates
// (128 MMAs per pair against 24 small ones), so the hand-off latencies that bind the small-d kernel are amortised.
constexpr int KS_STAGES = 3;
constexpr uint32_t KS_STAGE_BYTES = 2 * A_BLK + 2 * B1_BLK;          // A hi | A lo | B rows of sub-item a | sub-item b
constexpr int KS_NBAR = 2 * KS_STAGES + 4 * NSLOT + 2;               // full empty | b3_full b3_empty s_full p_ready | t_full t_free
constexpr int KS_THREADS = (NEPI + 2) * 32;

__global__ void __launch_bounds__(256) ks_image_kernel(GpView gp, const double* __restrict__ X, long R, int KB, const double* __restrict__ ymax2p,
                                                       uint8_t* __restrict__ img, double* __restrict__ Ki, double* __restrict__ sxs, float* __restrict__ rscg) {
    const int tile = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = gp.D, d = gp.d;
    const double ascale = gp.a * 1.4426950408889634;
    uint8_t* base = img + (size_t)tile * KB * 2 * A_BLK;
    const int hl = lane & 15, sub = lane >> 4;
    for (int rp = warp; rp < TM / 2; rp += 8) {                    // two rows per warp pass
        const int r = rp * 2 + sub;
        const long row = (long)tile * TM + r;
        double nx = 0.0, sx = 0.0;
        for (int cb = 0; cb < KB * KBLK; cb += 128) {              // 16 lanes x 8 columns per pass
            const int c0 = cb + hl * 8;
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                float sv[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int c = c0 + e + q;
                    const double val = (row < R && c < D) ? __ldg(X + row * (long)D + c) : 0.0;
                    nx = fma(val, val, nx);
                    if (c < d) sx += val;
                    sv[q] = (float)(ascale * val);
                }
                const __half2 h = __floats2half2_rn(sv[0], sv[1]);
                const float2 hf = __half22float2(h);
                const __half2 l = __floats2half2_rn(sv[0] - hf.x, sv[1] - hf.y);
                hi[e >> 1] = *(const uint32_t*)&h;
                lo[e >> 1] = *(const uint32_t*)&l;
            }
            if (c0 < KB * KBLK) {
                const int kb = c0 / KBLK;
                const uint32_t off = (uint32_t)r * 128u + (uint32_t)(((((c0 % KBLK) >> 3) ^ (r & 7)) & 7) << 4);
                *(uint4*)(base + (size_t)kb * 2 * A_BLK + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *(uint4*)(base + (size_t)kb * 2 * A_BLK + A_BLK + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
        }
        for (int o = 8; o >= 1; o >>= 1) { nx += __shfl_xor_sync(0xffffffffu, nx, o); sx += __shfl_xor_sync(0xffffffffu, sx, o); }
        const int sh = row_shift(gp.a, nx, __ldg(ymax2p));
        if (hl == 0) {
            Ki[(size_t)tile * TM + r] = ldexp(exp(-0.5 * gp.a * nx), TC_P_SHIFT - sh); sxs[(size_t)tile * TM + r] = sx;
            rscg[(size_t)tile * TM + r] = exp2f((float)sh);
        }
    }
}

template <int CLASS>
__global__ void __launch_bounds__(KS_THREADS, 1)
eval_tc_ks_kernel(GpView gp, TcDev st, const double* __restrict__ X, long R, int mode, const uint8_t* __restrict__ img,
                  const double* __restrict__ Kig, const double* __restrict__ sxg, const float* __restrict__ rscg,
                  double* __restrict__ out0, double* __restrict__ out1, double* __restrict__ out2, double* __restrict__ out3) {
    using C = Cfg<CLASS>;
    constexpr bool PDE = (CLASS == TC_PDE);
    constexpr int NT = C::NK + C::NKX + C::NKY;
    constexpr int NMAX = C::NK > C::NKY ? C::NK : C::NKY;
    constexpr uint32_t B3_SUB = 2 * NMAX * 128, B3_SLOT = 2 * B3_SUB;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0u) { asm volatile("trap;"); }
    uint8_t* sRing = smem;                                           // KS_STAGES x [A hi | A lo | B a | B b]
    uint8_t* sB3 = sRing + KS_STAGES * (size_t)KS_STAGE_BYTES;       // NSLOT pair slots
    double* feat = (double*)(sB3 + NSLOT * (size_t)B3_SLOT);
    double* xchg = feat + TM * TF_COUNT;
    TcColDesc* sdesc = (TcColDesc*)(xchg + 3 * TM * 4);
    double* scsum = (double*)(sdesc + TC_MAXCOL);                    // [TC_MAXCOL] column sums of the coefficients (baseline term)
    uint64_t* bars = (uint64_t*)(scsum + TC_MAXCOL);
    uint32_t* tmem_slot = (uint32_t*)(bars + KS_NBAR);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int D = gp.D, d = gp.d;
    const long ntiles = (R + TM - 1) / TM;
    const int nit = (int)((ntiles - (long)blockIdx.x + (long)gridDim.x - 1) / (long)gridDim.x);
    const int KB = (st.nstep + 3) / 4;
    const uint32_t bar0 = smem_u32(bars);
    auto k_full = [&](int i) { return bar0 + 8u * (uint32_t)i; };
    auto k_empty = [&](int i) { return bar0 + 8u * (uint32_t)(KS_STAGES + i); };
    auto b3_full = [&](int i) { return bar0 + 8u * (uint32_t)(2 * KS_STAGES + i); };
    auto b3_empty = [&](int i) { return bar0 + 8u * (uint32_t)(2 * KS_STAGES + NSLOT + i); };
    auto s_full = [&](int i) { return bar0 + 8u * (uint32_t)(2 * KS_STAGES + 2 * NSLOT + i); };
    auto p_ready = [&](int i) { return bar0 + 8u * (uint32_t)(2 * KS_STAGES + 3 * NSLOT + i); };
    const uint32_t t_full = bar0 + 8u * (uint32_t)(2 * KS_STAGES + 4 * NSLOT), t_free = t_full + 8u;

    if (tid == 0) {
        for (int i = 0; i < KS_STAGES; ++i) { mbar_init(k_full(i), 1); mbar_init(k_empty(i), 1); }
        for (int i = 0; i < NSLOT; ++i) { mbar_init(b3_full(i), 1); mbar_init(b3_empty(i), 1); mbar_init(s_full(i), 1); mbar_init(p_ready(i), NEPI); }
        mbar_init(t_full, 1); mbar_init(t_free, NEPI);
        fence_barrier_init();
    }
    if (warp == NEPI + 1) tmem_alloc(smem_u32(tmem_slot), 512);
    for (int c = tid; c < TC_MAXCOL; c += KS_THREADS) {
        TcColDesc dsc = st.desc[c < NT ? c : 0];
        const bool dead = (c >= NT || dsc.out == TO_PAD || (st.ntile_dom == 0 && c >= C::NK + C::NKX));
        if (dead) { dsc.out = TO_PAD; dsc.f1 = 0; dsc.f2 = 0; dsc.inv_scale = 0.0; }
        sdesc[c] = dsc;
        scsum[c] = dead ? 0.0 : st.csum[c];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    const int ndom = st.ntile_dom, nbdy = st.ntile_bdy;
    constexpr int CPD = PDE ? 3 : 2;
    const int nitem = CPD * ndom + (CPD - 1) * nbdy;
    const int npair = (nitem + 1) >> 1;
    auto next_item = [&](int& t, int& kc) {
        if (kc == TK_K) kc = PDE ? TK_KX : (t < ndom ? TK_KY : TK_K);
        else if (kc == TK_KX) kc = (t < ndom) ? TK_KY : TK_K;
        else kc = TK_K;
        if (kc == TK_K) ++t;
    };

    if (warp == NEPI) {
        // ===== producer: per pair and K block one ring stage [A hi | A lo | B rows a | B rows b]; per pair the coefficient images
        if (lane == 0) {
            long gs = 0;                                             // global stage counter
            int p3 = 0;
            for (int it = 0; it < nit; ++it) {
                const long tile = (long)blockIdx.x + (long)it * gridDim.x;
                const uint8_t* aimg = img + (size_t)tile * KB * 2 * A_BLK;
                int t1 = 0, k1 = TK_K, t3 = 0, k3 = TK_K;
                for (int j = 0; j < npair; ++j) {
                    const bool two = 2 * j + 1 < nitem;
                    int ta = t1, ka = k1; next_item(t1, k1);
                    int tb = t1, kbb = k1; if (two) next_item(t1, k1);
                    const uint8_t* srca = st.b1 + (size_t)ta * st.b1_tile_bytes + (size_t)(ka == TK_K ? 0 : (ka == TK_KX ? 1 : 2)) * B1_BLK;
                    const uint8_t* srcb = st.b1 + (size_t)tb * st.b1_tile_bytes + (size_t)(kbb == TK_K ? 0 : (kbb == TK_KX ? 1 : 2)) * B1_BLK;
                    // coefficient images of the pair first (small; needed after the K loop)
                    {
                        const int s = p3 % NSLOT;
                        if (p3 >= NSLOT) mbar_wait(b3_empty(s), (uint32_t)((p3 / NSLOT) - 1) & 1u);
                        uint32_t off[2], bytes[2]; int tt[2];
                        for (int sub = 0; sub < (two ? 2 : 1); ++sub) {
                            off[sub] = k3 == TK_K ? 0u : (k3 == TK_KX ? 2u * C::NK * 128u : 2u * (C::NK + C::NKX) * 128u);
                            bytes[sub] = 2u * 128u * (uint32_t)(k3 == TK_K ? C::NK : (k3 == TK_KX ? C::NKX : C::NKY));
                            tt[sub] = t3;
                            next_item(t3, k3);
                        }
                        mbar_expect_tx(b3_full(s), bytes[0] + (two ? bytes[1] : 0u));
                        for (int sub = 0; sub < (two ? 2 : 1); ++sub)
                            bulk_g2s(smem_u32(sB3 + (size_t)s * B3_SLOT + (size_t)sub * B3_SUB), st.b3 + (size_t)tt[sub] * st.b3_tile_bytes + off[sub],
                                     bytes[sub], b3_full(s));
                        ++p3;
                    }
                    for (int kb = 0; kb < KB; ++kb, ++gs) {
                        const int sg = (int)(gs % KS_STAGES);
                        if (gs >= KS_STAGES) mbar_wait(k_empty(sg), (uint32_t)((gs / KS_STAGES) - 1) & 1u);
                        uint8_t* dst = sRing + (size_t)sg * KS_STAGE_BYTES;
                        mbar_expect_tx(k_full(sg), 2 * A_BLK + (two ? 2u : 1u) * B1_BLK);
                        bulk_g2s(smem_u32(dst), aimg + (size_t)kb * 2 * A_BLK, 2 * A_BLK, k_full(sg));
                        bulk_g2s(smem_u32(dst + 2 * A_BLK), srca + (size_t)kb * (3 * B1_BLK), B1_BLK, k_full(sg));
                        if (two) bulk_g2s(smem_u32(dst + 2 * A_BLK + B1_BLK), srcb + (size_t)kb * (3 * B1_BLK), B1_BLK, k_full(sg));
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == NEPI + 1) {
        // ===== MMA issuer =====
        const uint32_t el = elect_one();
        const uint64_t ring0 = make_desc(smem_u32(sRing), 1, 64, 2);
        const uint64_t b3desc0 = make_desc(smem_u32(sB3), 1, 64, 2);
        const uint32_t idS2 = make_idesc(TM, 2 * TN), idS1 = make_idesc(TM, TN);
        const uint32_t idK = make_idesc(TM, C::NK), idKY = make_idesc(TM, C::NKY), idKX = make_idesc(TM, C::NKX > 0 ? C::NKX : 16);
        long gs = 0;
        int p1 = 0, p2 = 0;
        for (int it = 0; it < nit; ++it) {
            bool first_k = true, first_kx = true, first_ky = true;
            int kc2 = TK_K, t2 = 0;
            auto stage1 = [&](int j) {
                const int s = p1 % NSLOT;
                const uint32_t accS = tmem_base + (uint32_t)s * 128u;
                const uint32_t idesc = (2 * j + 1 < nitem) ? idS2 : idS1;
                ++p1;
                for (int kb = 0; kb < KB; ++kb, ++gs) {
                    const int sg = (int)(gs % KS_STAGES);
                    mbar_wait(k_full(sg), (uint32_t)(gs / KS_STAGES) & 1u);
                    const uint64_t ad = ring0 + (uint64_t)(((uint32_t)sg * KS_STAGE_BYTES) >> 4);
                    const uint64_t bd = ad + (uint64_t)((2 * A_BLK) >> 4);
                    const int nks = (st.nstep - kb * 4 < 4) ? (st.nstep - kb * 4) : 4;
                    if (el) {
#pragma unroll
                        for (int half = 1; half >= 0; --half) {      // low halves first (tiny terms), then the high halves
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                if (ks < nks)
                                    umma_f16(accS, ad + (uint64_t)((half * A_BLK + ks * 32) >> 4), bd + (uint64_t)((ks * 32) >> 4), idesc,
                                             (kb == 0 && half == 1 && ks == 0) ? 0u : 1u);
                            }
                        }
                        umma_commit(k_empty(sg));
                    }
                    __syncwarp();
                }
                if (el) umma_commit(s_full(s));
                __syncwarp();
            };
            auto stage2 = [&](int j) {
                const int s = p2 % NSLOT;
                {
                    const uint32_t par = (uint32_t)(p2 / NSLOT) & 1u;
                    uint32_t spins = 0;
                    bool a = false, b = false;
                    while (true) {
                        if (!a) a = mbar_test_wait(p_ready(s), par);
                        if (!b) b = mbar_test_wait(b3_full(s), par);
                        if (a && b) break;
                        if (++spins > SPIN_LIMIT) { asm volatile("trap;"); }
                    }
                }
                if (j == 0 && it > 0) mbar_wait(t_free, (uint32_t)(it - 1) & 1u);
                tc_fence_after();
                ++p2;
                const int nsub = (2 * j + 1 < nitem) ? 2 : 1;
                for (int sub = 0; sub < nsub; ++sub) {
                    const uint32_t pbase = tmem_base + (uint32_t)s * 128u + (uint32_t)sub * 64u;
                    const uint64_t b3 = b3desc0 + (uint64_t)(((uint32_t)s * B3_SLOT + (uint32_t)sub * B3_SUB) >> 4);
                    const int kc = kc2;
                    const uint32_t tacc = tmem_base + COL_T + (kc == TK_K ? 0u : (kc == TK_KX ? (uint32_t)C::NK : (uint32_t)(C::NK + C::NKX)));
                    const uint32_t nrows = kc == TK_K ? C::NK : (kc == TK_KX ? C::NKX : C::NKY);
                    const uint32_t idesc = kc == TK_K ? idK : (kc == TK_KX ? idKX : idKY);
                    const bool first = kc == TK_K ? first_k : (kc == TK_KX ? first_kx : first_ky);
                    const uint64_t clo0 = b3 + (uint64_t)((nrows * 128) >> 4);
                    if (el) {
                        if (first) umma_ts<false>(tacc, pbase, b3, idesc); else umma_ts<true>(tacc, pbase, b3, idesc);
                        umma_ts<true>(tacc, pbase, clo0, idesc);
                        umma_ts<true>(tacc, pbase + 8u, b3, idesc);
#pragma unroll
                        for (int ks = 1; ks < 4; ++ks) {
                            umma_ts<true>(tacc, pbase + (uint32_t)ks * 16u, b3 + (uint64_t)((ks * 32) >> 4), idesc);
                            umma_ts<true>(tacc, pbase + (uint32_t)ks * 16u, clo0 + (uint64_t)((ks * 32) >> 4), idesc);
                            umma_ts<true>(tacc, pbase + (uint32_t)ks * 16u + 8u, b3 + (uint64_t)((ks * 32) >> 4), idesc);
                        }
                    }
                    if (kc == TK_K) first_k = false; else if (kc == TK_KX) first_kx = false; else first_ky = false;
                    next_item(t2, kc2);
                }
                if (el) umma_commit(b3_empty(s));
                __syncwarp();
            };
            for (int j = 0; j < NSLOT && j < npair; ++j) stage1(j);
            for (int j = 0; j < npair; ++j) {
                stage2(j);
                if (j + NSLOT < npair) stage1(j + NSLOT);
            }
            if (el) umma_commit(t_full);
            __syncwarp();
        }
    } else {
        // ===== epilogue warps =====
        const int r = (warp & 3) * 32 + lane;
        const int cg = warp >> 2;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        long g = 0;
        for (int it = 0; it < nit; ++it) {
            const long row0 = ((long)blockIdx.x + (long)it * gridDim.x) * TM;
            const float sc = __ldg(rscg + row0 + r), nsc = -sc;     // 2^s of this thread's point row (scratch rows are padded to whole tiles)
            for (int j = 0; j < npair; ++j, ++g) {
                const int s = (int)(g % NSLOT);
                const bool two = 2 * j + 1 < nitem;
                mbar_wait(s_full(s), (uint32_t)(g / NSLOT) & 1u);
                tc_fence_after();
                const uint32_t base = tmem_base + lane_addr + (uint32_t)s * 128u + (uint32_t)cg * 16u;
                float v[2][16];
                tmem_ld16(base, v[0]);
                if (two) tmem_ld16(base + 64u, v[1]);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    if (c == 0 || two) {
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float p0 = fmaf(ex2f(v[c][2 * i]), sc, nsc), p1 = fmaf(ex2f(v[c][2 * i + 1]), sc, nsc);   // P - 2^s
                            hi[i] = pack_f16x2_sat(p0, p1);
                            const float2 hf = __half22float2(*(const __half2*)&hi[i]);
                            lo[i] = pack_f16x2_sat(p0 - hf.x, p1 - hf.y);
                        }
                        tmem_st8(base + (uint32_t)c * 64u, hi);
                        tmem_st8(base + (uint32_t)c * 64u + 8u, lo);
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_ready(s));
            }
            // --- final contraction (same as the small-d kernel)
            const long row = row0 + r;
            const bool ok = row < R;
            if (cg == 0) {
                const double* xr = X + row * (long)D;
                double* F = feat + r * TF_COUNT;
                const double sx = ok ? sxg[row0 + r] : 0.0;
                const double xt = ok ? __ldg(xr + d) : 0.0, x0 = ok ? __ldg(xr) : 0.0;
                double P2 = 0.0, R2 = 0.0;
#pragma unroll
                for (int m = 0; m < MC_IDX; ++m) {
                    const double xi = ok ? __ldg(xr + gp.I[m]) : 0.0, xir = ok ? __ldg(xr + gp.I[m] + 1) : 0.0;
                    F[TF_XI + m] = xi; F[TF_XR + m] = xir;
                    P2 = fma(xi, xi, P2); R2 = fma(xir, xir, R2);
                }
                F[TF_ONE] = 1.0; F[TF_SX] = sx; F[TF_XT] = xt; F[TF_X0] = x0; F[TF_SXR] = sx - x0 + xt; F[TF_P2] = P2; F[TF_R2] = R2;
            }
            mbar_wait(t_full, (uint32_t)it & 1u);
            tc_fence_after();
            constexpr int NPER = NT / 4;
            float tv[NPER];
#pragma unroll
            for (int c4 = 0; c4 < NPER; c4 += 4) tmem_ld4(tmem_base + lane_addr + COL_T + (uint32_t)(cg * NPER + c4), tv + c4);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_free);
            asm volatile("bar.sync 1, %0;" ::"n"(NEPI * 32) : "memory");
            const double* F = feat + r * TF_COUNT;
            double aU = 0.0, aG = 0.0, aL = 0.0, aT = 0.0, run = 0.0;
            int cur = TO_PAD;
            auto flush = [&](int o) {
                aU += (o == TO_U) ? run : 0.0; aG += (o == TO_G) ? run : 0.0;
                aL += (o == TO_L) ? run : 0.0; aT += (o == TO_T) ? run : 0.0;
                run = 0.0;
            };
#pragma unroll
            for (int i = 0; i < NPER; ++i) {
                const TcColDesc dsc = sdesc[cg * NPER + i];
                if ((int)dsc.out != cur) { flush(cur); cur = dsc.out; }
                run = fma(F[dsc.f1] * F[dsc.f2], ((double)tv[i] + (double)sc * scsum[cg * NPER + i]) * dsc.inv_scale, run);   // + 2^s sum_j C: baseline term
            }
            flush(cur);
            if (cg > 0) {
                double* p = xchg + ((size_t)(cg - 1) * TM + r) * 4;
                p[0] = aU; p[1] = aG; p[2] = aL; p[3] = aT;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NEPI * 32) : "memory");
            if (cg == 0 && ok) {
#pragma unroll
                for (int g2 = 1; g2 < 4; ++g2) {
                    const double* p = xchg + ((size_t)(g2 - 1) * TM + r) * 4;
                    aU += p[0]; aG += p[1]; aL += p[2]; aT += p[3];
                }
                const double ki = Kig[row0 + r];
                const double u = ki * aU;
                if (CLASS == TC_U) {
                    const double gt = 1.0 - 1.0 / (1.0 + exp(F[TF_SX] + F[TF_XT]));                  // equations.py:259
                    out0[row] = (mode == EVAL_TERMINAL) ? gt - u : u;
                } else if (CLASS == TC_UG) {
                    out0[row] = u;
                    out1[row] = ki * aG;
                } else {
                    const double gg = ki * aG, l = ki * aL, tt = ki * aT;
                    const double s2 = gp.sig2;
                    out0[row] = tt + (s2 * u - 1.0 / (double)d - 0.5 * s2) * gg + 0.5 * s2 * l;       // GP.py:767-768
                    if (out1) out1[row] = gg;
                    if (out2) out2[row] = l;
                    if (out3) out3[row] = tt;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NEPI * 32) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NEPI + 1) tmem_dealloc(tmem_base, 512);
}

template <int CLASS>
static int launch_ks(const GpView& gp, const TcDev& st, const double* X, long R, int mode, const uint8_t* img, const double* Ki,
                     const double* sxs, const float* rsc, double* o0, double* o1, double* o2, double* o3, cudaStream_t stream) {
    using C = Cfg<CLASS>;
    constexpr size_t NMAX = C::NK > C::NKY ? C::NK : C::NKY;
    const size_t smem = KS_STAGES * (size_t)KS_STAGE_BYTES + NSLOT * (2 * 2 * NMAX * 128) + (size_t)TM * TF_COUNT * 8 + 3 * (size_t)TM * 4 * 8
                        + TC_MAXCOL * (sizeof(TcColDesc) + sizeof(double)) + KS_NBAR * 8 + 16;
    int dev = 0, nsm = 0;
    SC_CUDA(cudaGetDevice(&dev));
    SC_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    SC_CUDA(cudaFuncSetAttribute(eval_tc_ks_kernel<CLASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long ntiles = cdiv(R, TM);
    const unsigned grid = (unsigned)(ntiles < nsm ? ntiles : nsm);
    eval_tc_ks_kernel<CLASS><<<grid, KS_THREADS, smem, stream>>>(gp, st, X, R, mode, img, Ki, sxs, rsc, o0, o1, o2, o3);
    SC_LAUNCH_CHECK();
    return OK;
}

}  // namespace tc

// ---- host API --------------------------------------------------------------------------------------------------

// the contraction axis (d + 1 coordinates + the exponent-shift column): up to two 64-wide K blocks stay resident in tensor
// memory (small-d kernel), up to sixteen are streamed (K-streamed kernel)
int tc_supported(const GpView& gp) { return gp.D + 1 <= 16 * tc::KBLK; }

static void add_col(TcState* st, int cls, int& n, int kern, int out, int f1, int f2, int coef, int m = 0, int nn = 0) {
    TcColSpec& s = st->spec[cls][n++];
    s.kern = (unsigned char)kern; s.out = (unsigned char)out; s.f1 = (unsigned char)f1; s.f2 = (unsigned char)f2;
    s.coef = (unsigned char)coef; s.m = (unsigned char)m; s.n = (unsigned char)nn; s.pad = 0;
}

// column table: kernel classes in the order [k | kx | ky] (= accumulator columns), columns of a class in the order of
// tests/tc_expansion_ref.py::columns, each class padded to its UMMA N
static void build_columns(TcState* st) {
    static const int NCOL[3][3] = {{16, 16, 0}, {16, 32, 0}, {48, 48, 32}};
    const int H[7] = {TF_ONE, TF_P2, TF_XI, TF_XI + 1, TF_XI + 2, TF_XI + 3, TF_XI + 4};
    const int MX[7] = {TF_ONE, TF_R2, TF_XR, TF_XR + 1, TF_XR + 2, TF_XR + 3, TF_XR + 4};
    const int G3[3] = {TF_ONE, TF_X0, TF_SXR};
    for (int cls = 0; cls < 3; ++cls) {
        for (int c = 0; c < TC_MAXCOL; ++c) { TcColSpec z{}; z.out = TO_PAD; z.coef = TCF_ZERO; st->spec[cls][c] = z; }
        int n = 0;
        // k class
        add_col(st, cls, n, TK_K, TO_U, TF_ONE, TF_ONE, TCF_U0); add_col(st, cls, n, TK_K, TO_U, TF_XT, TF_ONE, TCF_U1);
        add_col(st, cls, n, TK_K, TO_U, TF_SX, TF_ONE, TCF_U2);
        if (cls >= TC_UG) {
            add_col(st, cls, n, TK_K, TO_G, TF_ONE, TF_ONE, TCF_G0); add_col(st, cls, n, TK_K, TO_G, TF_SX, TF_ONE, TCF_GSX);
            add_col(st, cls, n, TK_K, TO_G, TF_XT, TF_ONE, TCF_GXT); add_col(st, cls, n, TK_K, TO_G, TF_SX, TF_XT, TCF_GSXXT);
            add_col(st, cls, n, TK_K, TO_G, TF_SX, TF_SX, TCF_GSX2);
        }
        if (cls == TC_PDE) {
            add_col(st, cls, n, TK_K, TO_T, TF_ONE, TF_ONE, TCF_T0); add_col(st, cls, n, TK_K, TO_T, TF_XT, TF_ONE, TCF_TXT);
            add_col(st, cls, n, TK_K, TO_T, TF_SX, TF_ONE, TCF_TSX); add_col(st, cls, n, TK_K, TO_T, TF_XT, TF_XT, TCF_TXT2);
            add_col(st, cls, n, TK_K, TO_T, TF_SX, TF_XT, TCF_TSXXT);
            add_col(st, cls, n, TK_K, TO_L, TF_ONE, TF_ONE, TCF_L1); add_col(st, cls, n, TK_K, TO_L, TF_R2, TF_ONE, TCF_LR2);
            add_col(st, cls, n, TK_K, TO_L, TF_R2, TF_R2, TCF_LR22);
            for (int m = 0; m < MC_IDX; ++m) add_col(st, cls, n, TK_K, TO_L, TF_XR + m, TF_ONE, TCF_LX, m);
            for (int m = 0; m < MC_IDX; ++m) add_col(st, cls, n, TK_K, TO_L, TF_XR + m, TF_R2, TCF_LXR2, m);
            for (int m = 0; m < MC_IDX; ++m)
                for (int q = m; q < MC_IDX; ++q) add_col(st, cls, n, TK_K, TO_L, TF_XR + m, TF_XR + q, TCF_LXX, m, q);
        }
        st->ncol[cls][TK_K] = NCOL[cls][TK_K];
        n = NCOL[cls][TK_K];
        // kx class
        if (cls == TC_PDE)
            for (int i = 0; i < 7; ++i)
                for (int j = 0; j < 3; ++j) add_col(st, cls, n, TK_KX, TO_L, MX[i], G3[j], TCF_MX, i, j);
        st->ncol[cls][TK_KX] = NCOL[cls][TK_KX];
        n = NCOL[cls][TK_K] + NCOL[cls][TK_KX];
        // ky class
        for (int i = 0; i < 7; ++i) add_col(st, cls, n, TK_KY, TO_U, H[i], TF_ONE, TCF_H, i);
        if (cls >= TC_UG) {
            for (int i = 0; i < 7; ++i) add_col(st, cls, n, TK_KY, TO_G, H[i], TF_SX, TCF_HSX, i);
            for (int i = 0; i < 7; ++i) add_col(st, cls, n, TK_KY, TO_G, H[i], TF_ONE, TCF_HG, i);
        }
        if (cls == TC_PDE) {
            for (int i = 0; i < 7; ++i) add_col(st, cls, n, TK_KY, TO_T, H[i], TF_XT, TCF_HXT, i);
            for (int i = 0; i < 7; ++i) add_col(st, cls, n, TK_KY, TO_T, H[i], TF_ONE, TCF_HT, i);
        }
        st->ncol[cls][TK_KY] = NCOL[cls][TK_KY];
    }
}

size_t tc_image_bytes(const GpView& gp, TcState* st) {
    st->nstep = (gp.D + 1 + 15) / 16;
    if (st->nstep <= 8) st->nstep = st->nstep <= 2 ? 2 : (st->nstep <= 4 ? 4 : (st->nstep <= 7 ? 7 : 8));   // instantiated k-step counts (small-d kernel)
    st->KB = (st->nstep + 3) / 4;
    st->compact = st->nstep <= 8 ? 1 : 0;            // resident-operand kernel: 48-centre tiles over the compact centre list
    st->tn = st->compact ? tc::TN2 : tc::TN;
    if (st->compact) {
        st->ntile_all = (int)cdiv(gp.Nd + gp.Nb, tc::TN2);
        st->ntile_dom = (int)cdiv(gp.Nd, tc::TN2);
        st->ntile_bdy = 0;
    } else {
        st->ntile_dom = gp.NdPad / tc::TN;
        st->ntile_bdy = gp.NbPad / tc::TN;
        st->ntile_all = st->ntile_dom + st->ntile_bdy;
    }
    build_columns(st);
    const int ntile = st->ntile_all;
    const size_t ncentres = (size_t)gp.NdPad + gp.NbPad;      // coefficient matrix rows: GpView's padded centre list
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 1023) & ~(size_t)1023; return o; };
    st->npair_all = (ntile + 1) / 2;
    const int ntile_img = st->compact ? 2 * st->npair_all : ntile;   // compact format: whole pairs (an odd last tile is padded with a zero tile)
    st->b1_tile_bytes = (size_t)st->KB * 3 * st->tn * 128;
    st->b1_off = take((size_t)ntile_img * st->b1_tile_bytes);
    for (int cls = 0; cls < 3; ++cls) {
        const int nt = st->ncol[cls][0] + st->ncol[cls][1] + st->ncol[cls][2];
        st->b3_tile_bytes[cls] = (size_t)2 * nt * 128;
        st->b3_off[cls] = take((size_t)ntile_img * st->b3_tile_bytes[cls]);
    }
    for (int cls = 0; cls < 3; ++cls) st->desc_off[cls] = take(TC_MAXCOL * sizeof(TcColDesc));
    st->spec_off = take(3 * TC_MAXCOL * sizeof(TcColSpec));
    st->ymax_off = take(sizeof(double));
    st->csum_off = take(3 * TC_MAXCOL * sizeof(double));
    st->scratch_off = take(3 * ncentres * TC_MAXCOL * sizeof(double) + 3 * TC_MAXCOL * sizeof(unsigned long long));
    st->total_bytes = off;
    return off;
}

int tc_build_images(const GpView& gp, const TcState& st, cudaStream_t stream) {
    SC_REQUIRE(tc_supported(gp), "tcgen05 route supports d <= 1022 (larger d: FP64 route)");
    SC_REQUIRE(st.images != nullptr, "tc: image buffer is null");
    const int ntile = st.ntile_all;
    const int ncentres = gp.NdPad + gp.NbPad;
    TcColSpec* spec_dev = (TcColSpec*)(st.images + st.spec_off);
    double* coef = (double*)(st.images + st.scratch_off);
    unsigned long long* colmax = (unsigned long long*)(coef + 3 * (size_t)ncentres * TC_MAXCOL);
    SC_CUDA(cudaMemcpyAsync(spec_dev, st.spec, sizeof(st.spec), cudaMemcpyHostToDevice, stream));   // st lives in the handle
    SC_CUDA(cudaMemsetAsync(colmax, 0, 3 * TC_MAXCOL * sizeof(unsigned long long), stream));
    tc::coef_kernel<<<dim3(ncentres, 3), TC_MAXCOL, 0, stream>>>(gp, spec_dev, ncentres, coef, colmax);
    SC_LAUNCH_CHECK();
    SC_CUDA(cudaMemsetAsync(st.images + st.ymax_off, 0, sizeof(double), stream));
    tc::ymax_kernel<<<(unsigned)cdiv(ncentres, 256), 256, 0, stream>>>(gp, ncentres, (unsigned long long*)(st.images + st.ymax_off));
    SC_LAUNCH_CHECK();
    tc::b1_image_kernel<<<ntile, 256, 0, stream>>>(gp, st.images + st.b1_off, st.b1_tile_bytes, st.KB, st.tn, st.compact, st.npair_all);
    SC_LAUNCH_CHECK();
    for (int cls = 0; cls < 3; ++cls) {
        tc::b3_image_kernel<<<ntile, 256, 0, stream>>>(gp, spec_dev, cls, ncentres, coef, colmax, st.images + st.b3_off[cls],
                                                        st.b3_tile_bytes[cls], st.ncol[cls][TK_K], st.ncol[cls][TK_KX], st.ncol[cls][TK_KY],
                                                        (TcColDesc*)(st.images + st.desc_off[cls]), st.tn, st.compact, st.npair_all);
        SC_LAUNCH_CHECK();
    }
    tc::csum_kernel<<<dim3(TC_MAXCOL, 3), 256, 0, stream>>>(ncentres, coef, colmax, (double*)(st.images + st.csum_off));
    SC_LAUNCH_CHECK();
    return OK;
}

int launch_eval_tc(const GpView& gp, const void* tc_state, const double* X, long R, int mode,
                   double* out0, double* out1, double* out2, double* out3, cudaStream_t stream,
                   const TcDebug* dbg, const uint8_t* rec, const double* recf) {
    if (R <= 0) return OK;
    const TcState* st = (const TcState*)tc_state;
    if (st == nullptr) st = (const TcState*)gp.tc;
    SC_REQUIRE(st != nullptr && st->images != nullptr, "tcgen05 route unavailable for this GP (d > 1022 or not fitted): use the FP64 route");
    SC_REQUIRE(tc_supported(gp), "tcgen05 route supports d <= 1022 (larger d: FP64 route)");
    SC_REQUIRE(X && out0, "eval: null pointer");
    const int cls = (mode == EVAL_PDE) ? TC_PDE : (mode == EVAL_UG ? TC_UG : TC_U);
    if (cls == TC_UG) SC_REQUIRE(out1 != nullptr, "eval UG: out1 is null");
    tc::TcDev dv;
    dv.b1 = st->images + st->b1_off;
    dv.b3 = st->images + st->b3_off[cls];
    dv.desc = (const TcColDesc*)(st->images + st->desc_off[cls]);
    dv.b1_tile_bytes = st->b1_tile_bytes;
    dv.b3_tile_bytes = st->b3_tile_bytes[cls];
    dv.nstep = st->nstep; dv.ntile_dom = st->ntile_dom; dv.ntile_bdy = st->ntile_bdy; dv.ntile_all = st->ntile_all; dv.npair_all = st->npair_all;
    dv.rec = rec; dv.recf = recf;
    dv.ymax2 = (const double*)(st->images + st->ymax_off);
    dv.csum = (const double*)(st->images + st->csum_off) + cls * TC_MAXCOL;
    dv.dbg = dbg ? dbg->stamps : nullptr; dv.dbg_block = dbg ? (dbg->block & 0xFFFFFF) : 0; dv.dbg_flags = dbg ? (dbg->block >> 24) : 0;
    if (st->nstep > 8) {
        // K-streamed kernel: point images in a stream-ordered scratch buffer of this call (cudaMallocAsync: concurrent evaluations of
        // one handle on different streams never share it), processed in chunks
        const int KB = (st->nstep + 3) / 4;
        const long chunk_pts = 148L * tc::TM * 4;                     // four point tiles per SM per launch
        const size_t per_tile = (size_t)KB * 2 * tc::A_BLK + 2 * tc::TM * sizeof(double) + tc::TM * sizeof(float);
        const size_t need = (size_t)cdiv(R < chunk_pts ? R : chunk_pts, tc::TM) * per_tile;
        uint8_t* scratch = nullptr;
        { const int prc = ensure_scratch_pool(); if (prc != OK) return prc; }
        SC_CUDA(cudaMallocAsync((void**)&scratch, need, stream));
        int rcode = OK;
        for (long r0 = 0; r0 < R; r0 += chunk_pts) {
            const long rc = (R - r0 < chunk_pts) ? (R - r0) : chunk_pts;
            const long nt = cdiv(rc, tc::TM);
            uint8_t* img = scratch;
            double* Ki = (double*)(img + (size_t)nt * KB * 2 * tc::A_BLK);
            double* sx = Ki + nt * tc::TM;
            float* rsc = (float*)(sx + nt * tc::TM);
            const double* Xc = X + r0 * (long)gp.D;
            tc::ks_image_kernel<<<(unsigned)nt, 256, 0, stream>>>(gp, Xc, rc, KB, dv.ymax2, img, Ki, sx, rsc);
            if (cudaGetLastError() != cudaSuccess) { rcode = ERR_CUDA; set_error("ks_image_kernel launch failed"); break; }
            double* p0 = out0 + r0; double* p1 = out1 ? out1 + r0 : nullptr; double* p2 = out2 ? out2 + r0 : nullptr; double* p3 = out3 ? out3 + r0 : nullptr;
            if (cls == TC_U) rcode = tc::launch_ks<TC_U>(gp, dv, Xc, rc, mode, img, Ki, sx, rsc, p0, p1, p2, p3, stream);
            else if (cls == TC_UG) rcode = tc::launch_ks<TC_UG>(gp, dv, Xc, rc, mode, img, Ki, sx, rsc, p0, p1, p2, p3, stream);
            else rcode = tc::launch_ks<TC_PDE>(gp, dv, Xc, rc, mode, img, Ki, sx, rsc, p0, p1, p2, p3, stream);
            if (rcode != OK) break;
        }
        cudaFreeAsync(scratch, stream);
        return rcode;
    }
    // resident-operand kernel.  The k-step count is a template parameter (lean single-thread MMA issue): images are zero-padded up to it
    auto run = [&](const double* Xc, long rc, const uint8_t* recs, const double* recfs, double* p0, double* p1, double* p2, double* p3) -> int {
        tc::TcDev d2 = dv;
        d2.rec = recs; d2.recf = recfs;
#define SC_TC_DISPATCH(NS)                                                                                                  \
    do {                                                                                                                    \
        if (cls == TC_U) return tc::launch<TC_U, NS>(gp, d2, Xc, rc, mode, p0, p1, p2, p3, stream);                          \
        if (cls == TC_UG) return tc::launch<TC_UG, NS>(gp, d2, Xc, rc, mode, p0, p1, p2, p3, stream);                        \
        return tc::launch<TC_PDE, NS>(gp, d2, Xc, rc, mode, p0, p1, p2, p3, stream);                                         \
    } while (0)
        if (st->nstep <= 2) SC_TC_DISPATCH(2);
        if (st->nstep <= 4) SC_TC_DISPATCH(4);
        if (st->nstep <= 7) SC_TC_DISPATCH(7);
        SC_TC_DISPATCH(8);
#undef SC_TC_DISPATCH
    };
    SC_REQUIRE(st->nstep == tc_rec_nstep(gp.D), "tc: k-step count of the images and of the operand records differ");
    SC_REQUIRE((rec == nullptr) == (recf == nullptr), "tc: operand records and feature blocks come together");
    if (rec != nullptr) return run(X, R, rec, recf, out0, out1, out2, out3);
    // a caller without records (public evaluation API, tools): a pre-pass writes them into a stream-ordered scratch buffer of this call, in
    // chunks of eight point tiles per SM (the scratch stays L2-sized: 70 MB at d = 100)
    const size_t recb = (size_t)tc_rec_bytes(st->nstep);
    const long chunk_pts = 148L * tc::TM * 8;
    uint8_t* scratch = nullptr;
    { const int prc = ensure_scratch_pool(); if (prc != OK) return prc; }
    const size_t npts = (size_t)(R < chunk_pts ? R : chunk_pts);
    const size_t recs_bytes = (npts * recb + 255) & ~(size_t)255;
    SC_CUDA(cudaMallocAsync((void**)&scratch, recs_bytes + npts * TC_REC_NFEAT * sizeof(double), stream));
    double* scratchf = (double*)(scratch + recs_bytes);
    int rcode = OK;
    for (long r0 = 0; r0 < R && rcode == OK; r0 += chunk_pts) {
        const long rc = (R - r0 < chunk_pts) ? (R - r0) : chunk_pts;
        const double* Xc = X + r0 * (long)gp.D;
        tc::rec_image_kernel<<<(unsigned)cdiv(rc * 32, 256), 256, 0, stream>>>(Xc, rc, gp.D, gp.d, st->nstep, tc_rec_ascale(gp.a), tc_rec_idx(gp), scratch, scratchf);
        if (cudaGetLastError() != cudaSuccess) { rcode = ERR_CUDA; set_error("rec_image_kernel launch failed"); break; }
        rcode = run(Xc, rc, scratch, scratchf, out0 + r0, out1 ? out1 + r0 : nullptr, out2 ? out2 + r0 : nullptr, out3 ? out3 + r0 : nullptr);
    }
    cudaFreeAsync(scratch, stream);
    return rcode;
}

#ifdef SCASML_DEBUG_HOOKS
// timeline of one CTA: stamps[0] entry, [1] A images built, [2] prologue done, [3] exit,
// tile t: [4+4t] stage-1 issue start, [5+4t] stage-2 issue end, [6+4t] epilogue start, [7+4t] epilogue end   (SM clock cycles)
int tc_timeline(const GpView& gp, const TcState& st, const double* X, long R, int mode, int block, long long* stamps_dev,
                double* scratch_out, cudaStream_t stream) {
    TcDebug dbg;
    dbg.stamps = stamps_dev;
    dbg.block = block;
    return launch_eval_tc(gp, &st, X, R, mode, scratch_out, scratch_out + R, scratch_out + 2 * R, scratch_out + 3 * R, stream, &dbg);
}
#endif

}  // namespace scasml
