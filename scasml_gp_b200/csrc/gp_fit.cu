// GP fit on the device, FP64 (sm_100a).  Reference: models/GP.py:182-268 (Gram, "Cholesky"),
// :430-444, 487-604, 705-719 (Newton on J(sol) = b^T (K + nu I)^{-1} b, alpha = (K + nu I)^{-1} z).
//
//   gram_kernel          25 Gram blocks from three distance dot products, closed forms of SURVEY App. B,
//                        entries rounded to float16 (models/GP.py:258 semantics), nugget on the diagonal
//   cholesky (blocked)   diagonal-block factor + its inverse (fused triangular solve of the panel),
//                        trailing update through dgemm_kernel
//   inverse              P = L^-T L^-1 (block forward substitution + one GEMM)
//   newton               closed-form gradient / Hessian, blocked LU with partial pivoting
#include <vector>
#include <cmath>
#include "gp.cuh"
#include "gp_fit.cuh"

namespace scasml {

namespace {

// ------------------------------------------------------------------ Gram ---------------------------------
constexpr int GT = 16;   // 16x16 pairs per CTA

__global__ void __launch_bounds__(256) gram_kernel(GpView gp, double* __restrict__ K, long phi, double nugget, int f16_entries) {
    __shared__ double Xs[GT][GT + 1], Xrs[GT][GT + 1], Ys[GT][GT + 1], Yrs[GT][GT + 1];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int Nd = gp.Nd, Nb = gp.Nb, Nc = Nd + Nb, D = gp.D, d = gp.d;
    const int i = blockIdx.y * GT + ty;      // row centre (unpadded index: domain then boundary)
    const int j = blockIdx.x * GT + tx;      // column centre
    // padded storage index of a centre
    auto slot = [&](int c) { return c < Nd ? c : gp.NdPad + (c - Nd); };
    const int li = blockIdx.y * GT + tx, lj = blockIdx.x * GT + tx;   // loaders: thread (ty = k, tx = centre)
    double d1 = 0.0, d2 = 0.0, d3 = 0.0;
    for (int k0 = 0; k0 < D; k0 += GT) {
        const int k = k0 + ty;
        double xv = 0.0, xr = 0.0, yv = 0.0, yr = 0.0;
        if (k < D) {
            const int kr = (k + 1 == D) ? 0 : k + 1;
            if (li < Nc) { const double* p = gp.C + (long)slot(li) * D; xv = p[k]; xr = p[kr]; }
            if (lj < Nc) { const double* p = gp.C + (long)slot(lj) * D; yv = p[k]; yr = p[kr]; }
        }
        __syncthreads();
        Xs[ty][tx] = xv; Xrs[ty][tx] = xr; Ys[ty][tx] = yv; Yrs[ty][tx] = yr;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GT; ++kk) {
            const double x = Xs[kk][ty], y = Ys[kk][tx];
            d1 = fma(x, y, d1);
            d2 = fma(x, Yrs[kk][tx], d2);
            d3 = fma(Xrs[kk][ty], y, d3);
        }
    }
    if (i >= Nc || j >= Nc) return;
    const double* fx = gp.feat + (long)slot(i) * CF_STRIDE;
    const double* fy = gp.feat + (long)slot(j) * CF_STRIDE;
    const double a = gp.a, a2 = a * a, a3 = a2 * a, dd = (double)d, inv5 = 1.0 / MC_IDX;
    const double nn = fx[CF_NY] + fy[CF_NY];
    const double k = exp(-0.5 * a * (nn - 2.0 * d1));
    const double ky = exp(-0.5 * a * (nn - 2.0 * d2));
    const double kx = exp(-0.5 * a * (nn - 2.0 * d3));
    const double S = fx[CF_SY] - fy[CF_SY], rt = fx[CF_YT] - fy[CF_YT];
    const double Sy = fx[CF_SY] - fy[CF_SYROLL], Sx = fx[CF_SYROLL] - fy[CF_SY];
    const double ryd = fx[CF_YT] - fy[CF_Y0], rxd = fx[CF_Y0] - fy[CF_YT];
    double m1 = 0, m2 = 0, n1 = 0, n2 = 0, q2 = 0;
#pragma unroll
    for (int m = 0; m < MC_IDX; ++m) {
        const double ry = fx[CF_YI + m] - fy[CF_YIR + m];  m1 += ry; m2 = fma(ry, ry, m2);
        const double rx = fx[CF_YIR + m] - fy[CF_YI + m];  n1 += rx; n2 = fma(rx, rx, n2);
        const double q = fx[CF_YIR + m] - fy[CF_YIR + m];  q2 = fma(q, q, q2);
    }
    const double MH = a2 * m2 * inv5 - a, MHx = a2 * n2 * inv5 - a;
    // functional values, indexed [rowop][colop] with op order {id, lap, dt, div}
    double v[4][4];
    v[0][0] = k;
    v[0][1] = dd * MH * ky;
    v[0][2] = a * rt * k;
    v[0][3] = a * S * k;
    v[1][0] = dd * MHx * kx;
    {
        const double A = a2 * q2 - MC_IDX * a;
        v[1][1] = (dd * dd / (MC_IDX * MC_IDX)) * k * (A * A + 2.0 * MC_IDX * a2 - 4.0 * a3 * q2);
    }
    v[1][2] = a * rxd * dd * MHx * kx;
    v[1][3] = dd * (-2.0 * a2 * n1 * inv5 - a2 * Sx + a3 * Sx * n2 * inv5) * kx;
    v[2][0] = -a * rt * k;
    v[2][1] = -a * ryd * dd * MH * ky;
    v[2][2] = (a - a2 * rt * rt) * k;
    v[2][3] = -a2 * rt * S * k;
    v[3][0] = -a * S * k;
    v[3][1] = dd * (2.0 * a2 * m1 * inv5 + a2 * Sy - a3 * Sy * m2 * inv5) * ky;
    v[3][2] = -a2 * rt * S * k;
    v[3][3] = (a * dd - a2 * S * S) * k;
    const bool xdom = i < Nd, ydom = j < Nd;
    const long rbase[4] = {xdom ? 0 : Nd, (long)Nd + Nb, 2L * Nd + Nb, 3L * Nd + Nb};
    const long cbase[4] = {ydom ? 0 : Nd, (long)Nd + Nb, 2L * Nd + Nb, 3L * Nd + Nb};
    const int il = xdom ? i : i - Nd, jl = ydom ? j : j - Nd;
    const int nrop = xdom ? 4 : 1, ncop = ydom ? 4 : 1;
    for (int ro = 0; ro < nrop; ++ro)
        for (int co = 0; co < ncop; ++co) {
            double val = v[ro][co];
            if (f16_entries) val = round_f16(val);
            const long r = rbase[ro] + il, c = cbase[co] + jl;
            if (r == c) val += nugget;
            K[r * phi + c] = val;
        }
}

// ------------------------------------------------------------------ DGEMM --------------------------------
// C[M x N] = alpha * A * B + beta * C, A(i,k) = A[i*sai + k*sak], B(k,j) = B[k*sbk + j*sbj], C row-major.
// lower_only: skip 64x64 tiles strictly above the diagonal (symmetric updates).
constexpr int GM = 64, GN = 64, GK = 8;

__global__ void __launch_bounds__(256) dgemm_kernel(int M, int N, int Kd, double alpha,
                                                    const double* A, long sai, long sak,
                                                    const double* B, long sbk, long sbj,
                                                    double beta, double* C, long ldc, int lower_only) {
    if (lower_only && blockIdx.x > blockIdx.y) return;
    __shared__ double As[GK][GM + 2], Bs[GK][GN + 2];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.y * GM, j0 = blockIdx.x * GN;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int k0 = 0; k0 < Kd; k0 += GK) {
        __syncthreads();
        for (int idx = tid; idx < GM * GK; idx += 256) {
            int r, kk;
            if (sak == 1) { r = idx / GK; kk = idx % GK; } else { kk = idx / GM; r = idx % GM; }
            const int gi = i0 + r, gk = k0 + kk;
            As[kk][r] = (gi < M && gk < Kd) ? A[gi * sai + gk * sak] : 0.0;
        }
        for (int idx = tid; idx < GN * GK; idx += 256) {
            int c, kk;
            if (sbk == 1) { c = idx / GK; kk = idx % GK; } else { kk = idx / GN; c = idx % GN; }
            const int gj = j0 + c, gk = k0 + kk;
            Bs[kk][c] = (gj < N && gk < Kd) ? B[gk * sbk + gj * sbj] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            double av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = i0 + ty * 4 + i;
        if (gi >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gj = j0 + tx * 4 + j;
            if (gj >= N) continue;
            double* c = C + gi * ldc + gj;
            *c = alpha * acc[i][j] + (beta == 0.0 ? 0.0 : beta * (*c));
        }
    }
}


// Large-tile variant for the O(n^3) products of the fit (trailing updates of the blocked Cholesky / LU, L^-T L^-1) on the FP64 TENSOR
// cores: mma.sync.m8n8k4.f64 (DMMA).  Vector FP64 runs at ~1/8 of the FP32 rate on this part (an 8 x 8-per-thread FMA tile topped out at
// 5.3 TFLOP/s), the FP64 tensor path is what cuBLAS' 35 TFLOP/s comes from.  128 x 128 tile, K tile 16, 8 warps as 4 (M) x 2 (N): a warp
// owns 32 x 64 = 4 x 8 DMMA tiles (64 accumulator doubles per thread); next K tile's global loads in flight during the current tile's MMAs.
// Fragment layout (PTX ISA, m8n8k4 .f64): a = A[row lane/4][k lane%4], b = B[k lane%4][col lane/4], c0,c1 = C[row lane/4][col 2 (lane%4) + {0,1}].
// Shared-memory pitch 132 doubles: the 16 lanes of a half-warp (4 rows x 4 k) hit 16 distinct banks.  Same operand conventions as dgemm_kernel.
constexpr int HM = 128, HN = 128, HK = 16;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) dgemm_big_kernel(int M, int N, int Kd, double alpha,
                                                        const double* __restrict__ A, long sai, long sak,
                                                        const double* __restrict__ B, long sbk, long sbj,
                                                        double beta, double* C, long ldc, int lower_only) {
    if (lower_only && blockIdx.x > blockIdx.y) return;
    __shared__ double As[HK][HM + 4], Bs[HK][HN + 4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;               // warp tile: rows 32 wm .., columns 64 wn ..
    const int fr = lane >> 2, fk = lane & 3;               // fragment row / k (A), column / k (B)
    const int i0 = blockIdx.y * HM, j0 = blockIdx.x * HN;
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    double pa[8], pb[8];                                   // global -> register staging of the next K tile
    auto fetch = [&](int k0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + 256 * e;
            int r, kk;
            if (sak == 1) { r = idx / HK; kk = idx % HK; } else { kk = idx / HM; r = idx % HM; }
            const int gi = i0 + r, gk = k0 + kk;
            pa[e] = (gi < M && gk < Kd) ? A[gi * sai + gk * sak] : 0.0;
            int c, kb;
            if (sbk == 1) { c = idx / HK; kb = idx % HK; } else { kb = idx / HN; c = idx % HN; }
            const int gj = j0 + c, gk2 = k0 + kb;
            pb[e] = (gj < N && gk2 < Kd) ? B[gk2 * sbk + gj * sbj] : 0.0;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int idx = tid + 256 * e;
            int r, kk;
            if (sak == 1) { r = idx / HK; kk = idx % HK; } else { kk = idx / HM; r = idx % HM; }
            As[kk][r] = pa[e];
            int c, kb;
            if (sbk == 1) { c = idx / HK; kb = idx % HK; } else { kb = idx / HN; c = idx % HN; }
            Bs[kb][c] = pb[e];
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < Kd; k0 += HK) {
        __syncthreads();
        stash();
        __syncthreads();
        if (k0 + HK < Kd) fetch(k0 + HK);
#pragma unroll
        for (int k4 = 0; k4 < HK; k4 += 4) {
            double a[4], b[8];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) a[mt] = As[k4 + fk][wm * 32 + mt * 8 + fr];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) b[nt] = Bs[k4 + fk][wn * 64 + nt * 8 + fr];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
    }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int gi = i0 + wm * 32 + mt * 8 + fr;
        if (gi >= M) continue;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int gj = j0 + wn * 64 + nt * 8 + fk * 2 + q;
                if (gj >= N) continue;
                double* c = C + gi * ldc + gj;
                *c = alpha * acc[mt][nt][q] + (beta == 0.0 ? 0.0 : beta * (*c));
            }
        }
    }
}

// ------------------------------------------------------------------ Cholesky ------------------------------
constexpr int CB = 64;

// factor the nb x nb (nb <= 64) diagonal block at A[j0][j0] in place (lower).  256 threads as a 16 x 16 grid over the block: per
// column one thread takes the square root, the column is scaled, and every thread updates its 16 elements of the trailing part
// (no integer divisions, three barriers per column).
__global__ void __launch_bounds__(256) chol_diag_kernel(double* __restrict__ A, long n, int j0, int nb, int* __restrict__ fail) {
    __shared__ double Ls[CB][CB + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    for (int idx = tid; idx < CB * CB; idx += 256) {
        const int r = idx / CB, c = idx % CB;
        Ls[r][c] = (r < nb && c < nb && c <= r) ? A[(long)(j0 + r) * n + j0 + c] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        // column j: every thread below the diagonal reads the raw pivot and scales its element by the reciprocal square root itself
        // (no separate pivot step, no division); thread j stores the square root after the barrier (the update never reads L[j][j])
        const double p = Ls[j][j];
        const bool bad = !(p > 0.0);
        if (tid > j && tid < nb) Ls[tid][j] *= bad ? 1.0 : rsqrt(p);
        __syncthreads();
        if (tid == j) { if (bad) *fail = 1; Ls[j][j] = bad ? 1.0 : sqrt(p); }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int r = ty + 16 * a;
            const double lrj = Ls[r][j];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int c = tx + 16 * b;
                if (c > j && c <= r && r < nb) Ls[r][c] = fma(-lrj, Ls[c][j], Ls[r][c]);
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < CB * CB; idx += 256) {
        const int r = idx / CB, c = idx % CB;
        if (r < nb && c < nb) A[(long)(j0 + r) * n + j0 + c] = (c <= r) ? Ls[r][c] : 0.0;
    }
}

// fused panel solve of the blocked Cholesky: L21 = A21 L11^-T, one thread per row of A21 (forward substitution against the freshly
// factored diagonal block held in shared memory); the row tile goes through shared memory so that global traffic stays coalesced.
constexpr int TRSM_ROWS = 128;
__global__ void __launch_bounds__(TRSM_ROWS) trsm_panel_kernel(double* __restrict__ A, long n, int j0, int nb, long rest) {
    extern __shared__ double trsm_smem[];
    double (*L11)[CB + 1] = (double (*)[CB + 1])trsm_smem;
    double (*Xs)[CB + 1] = (double (*)[CB + 1])(trsm_smem + CB * (CB + 1));   // [TRSM_ROWS][CB + 1]
    const int tid = threadIdx.x;
    const long row0 = (long)blockIdx.x * TRSM_ROWS;
    for (int idx = tid; idx < CB * CB; idx += TRSM_ROWS) {
        const int r = idx / CB, c = idx % CB;
        L11[r][c] = (r < nb && c < nb) ? A[(long)(j0 + r) * n + j0 + c] : (r == c ? 1.0 : 0.0);
    }
    double* A21 = A + (long)(j0 + nb) * n + j0;
    for (int idx = tid; idx < TRSM_ROWS * CB; idx += TRSM_ROWS) {
        const int r = idx / CB, c = idx % CB;
        Xs[r][c] = (row0 + r < rest && c < nb) ? A21[(row0 + r) * n + c] : 0.0;
    }
    __syncthreads();
    // right-looking substitution with the row in registers (fully unrolled: every FMA of a step is independent, the L11 entries are
    // shared-memory broadcasts, the diagonal enters through its reciprocal); padded columns (nb < 64) see an identity block
    __shared__ double rdiag[CB];
    if (tid < CB) rdiag[tid] = 1.0 / L11[tid][tid];
    __syncthreads();
    double x[CB];
#pragma unroll
    for (int c = 0; c < CB; ++c) x[c] = Xs[tid][c];
#pragma unroll
    for (int c = 0; c < CB; ++c) {
        x[c] *= rdiag[c];
#pragma unroll
        for (int c2 = c + 1; c2 < CB; ++c2) x[c2] = fma(-x[c], L11[c2][c], x[c2]);
    }
#pragma unroll
    for (int c = 0; c < CB; ++c) Xs[tid][c] = x[c];
    __syncthreads();
    for (int idx = tid; idx < TRSM_ROWS * CB; idx += TRSM_ROWS) {
        const int r = idx / CB, c = idx % CB;
        if (row0 + r < rest && c < nb) A21[(row0 + r) * n + c] = Xs[r][c];
    }
}
constexpr size_t TRSM_SMEM = (size_t)(CB + TRSM_ROWS) * (CB + 1) * sizeof(double);

// inverses of ALL 64 x 64 diagonal blocks of a factored matrix at once (one CTA per block, off the factorisation's critical path);
// they seed the explicit triangular inverse and serve the block substitutions of chol_solve.
__global__ void __launch_bounds__(CB) diag_inverse_kernel(const double* __restrict__ A, long n, double* __restrict__ invdiag) {
    extern __shared__ double dinv_smem[];
    double (*Ls)[CB + 1] = (double (*)[CB + 1])dinv_smem;
    double (*Iv)[CB + 1] = (double (*)[CB + 1])(dinv_smem + CB * (CB + 1));  // TRANSPOSED: Iv[c][r] = inv(L)[r][c]
    const long j0 = (long)blockIdx.x * CB;
    const int nb = (int)((n - j0 < CB) ? (n - j0) : CB);
    const int t = threadIdx.x;
    for (int rr = 0; rr < CB; ++rr) Ls[rr][t] = (rr < nb && t < nb && t <= rr) ? A[(j0 + rr) * n + j0 + t] : (rr == t ? 1.0 : 0.0);
    __syncthreads();
    const int c = t;                                                        // column c of inv(L) = row c of Iv: private to this thread
    for (int rr = 0; rr < CB; ++rr) {
        double s0 = (rr == c) ? 1.0 : 0.0, s1 = 0.0;
        int k = c;
        for (; k + 1 < rr; k += 2) { s0 = fma(-Ls[rr][k], Iv[c][k], s0); s1 = fma(-Ls[rr][k + 1], Iv[c][k + 1], s1); }
        if (k < rr) s0 = fma(-Ls[rr][k], Iv[c][k], s0);
        Iv[c][rr] = (rr < c) ? 0.0 : (s0 + s1) / Ls[rr][rr];
    }
    __syncthreads();
    double* out = invdiag + (size_t)blockIdx.x * CB * CB;
    for (int rr = 0; rr < CB; ++rr) out[rr * CB + t] = Iv[t][rr];
}
constexpr size_t DINV_SMEM = 2 * CB * (CB + 1) * sizeof(double);

__global__ void zero_upper_kernel(double* A, long n) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const long r = idx / n, c = idx % n;
    if (c > r) A[idx] = 0.0;
}

__global__ void mirror_lower_kernel(double* A, long n) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * n) return;
    const long r = idx / n, c = idx % n;
    if (c > r) A[idx] = A[c * n + r];
}

// ------------------------------------------------------------------ Newton pieces -------------------------
struct FitDims { int N, Nb; long phi; double s2, c1; };

// b = [z1; g; z3; F(z1,z3,z5); z5]   (models/GP.py:436, 705-719)
__global__ void build_b_kernel(FitDims f, const double* __restrict__ sol, const double* __restrict__ g, double* __restrict__ b) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f.phi) return;
    const int N = f.N, Nb = f.Nb;
    double v;
    if (i < N) v = sol[i];
    else if (i < N + Nb) v = g[i - N];
    else if (i < 2L * N + Nb) v = sol[N + (i - N - Nb)];
    else if (i < 3L * N + Nb) {
        const long k = i - 2L * N - Nb;
        const double z1 = sol[k], z3 = sol[N + k], z5 = sol[2 * N + k];
        v = -f.s2 * z1 * z5 + f.c1 * z5 - 0.5 * f.s2 * z3;
    } else v = sol[2 * N + (i - 3L * N - Nb)];
    b[i] = v;
}

// y = A x for row-major A[n x n], one warp per row
__global__ void __launch_bounds__(256) gemv_kernel(long n, const double* __restrict__ A, const double* __restrict__ x, double* __restrict__ y) {
    const long r = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const double* a = A + r * n;
    double acc = 0.0;
    for (long c = lane; c < n; c += 32) acc = fma(a[c], x[c], acc);
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[r] = acc;
}

// grad J = 2 Jm^T w;  rhs = -grad
__global__ void grad_kernel(FitDims f, const double* __restrict__ sol, const double* __restrict__ w,
                            double* __restrict__ grad, double* __restrict__ rhs) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int N = f.N;
    if (r >= 3 * N) return;
    const int blk = r / N, k = r % N;
    const long off[3] = {0, (long)N + f.Nb, 3L * N + f.Nb};
    const long o4 = 2L * N + f.Nb;
    const double z1 = sol[k], z5 = sol[2 * N + k];
    const double Dv = (blk == 0) ? -f.s2 * z5 : (blk == 1 ? -0.5 * f.s2 : -f.s2 * z1 + f.c1);
    const double g = 2.0 * (w[off[blk] + k] + Dv * w[o4 + k]);
    grad[r] = g; rhs[r] = -g;
}

// H = 2 Jm^T P Jm + 2 C + damping I   
__global__ void hessian_kernel(FitDims f, const double* __restrict__ sol, const double* __restrict__ w,
                               const double* __restrict__ P, double damping, double* __restrict__ H) {
    const int N = f.N;
    const long n3 = 3L * N;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n3 * n3) return;
    const long r = idx / n3, c = idx % n3;
    const int bi = (int)(r / N), bj = (int)(c / N), ri = (int)(r % N), cj = (int)(c % N);
    const long off[3] = {0, (long)N + f.Nb, 3L * N + f.Nb};
    const long o4 = 2L * N + f.Nb, phi = f.phi;
    auto Dval = [&](int blk, int k) {
        return (blk == 0) ? -f.s2 * sol[2 * N + k] : (blk == 1 ? -0.5 * f.s2 : -f.s2 * sol[k] + f.c1);
    };
    const double Di = Dval(bi, ri), Dj = Dval(bj, cj);
    double v = P[(off[bi] + ri) * phi + off[bj] + cj] + Di * P[(o4 + ri) * phi + off[bj] + cj]
             + P[(off[bi] + ri) * phi + o4 + cj] * Dj + Di * P[(o4 + ri) * phi + o4 + cj] * Dj;
    v *= 2.0;
    if (ri == cj && ((bi == 0 && bj == 2) || (bi == 2 && bj == 0))) v += 2.0 * (-f.s2) * w[o4 + ri];
    if (r == c) v += damping;
    H[idx] = v;
}

__global__ void __launch_bounds__(256) dot_kernel(long n, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out) {
    __shared__ double red[8];
    double acc = 0.0;
    for (long i = threadIdx.x; i < n; i += 256) acc = fma(x[i], y[i], acc);
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0.0; for (int i = 0; i < 8; ++i) s += red[i]; *out = s; }
}

__global__ void axpy_kernel(long n, double alpha, const double* __restrict__ x, double* __restrict__ y) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] += alpha * x[i];
}

// ------------------------------------------------------------------ LU (partial pivoting) -----------------
constexpr int LB = 64;


// ---- pivoted LU, panel on a column-major copy --------------------------------------------------------------------------------
// The single-CTA panel factorisation used to scan matrix columns in place (stride n doubles: one DRAM sector per element, ~50 us per
// column at n = 12 000).  Now the (rows x nb) panel is gathered into a column-major buffer, factored there with coalesced column
// scans, scattered back, and its row interchanges are applied to the rest of the matrix (and to rhs) by a row-coalesced kernel.
__global__ void __launch_bounds__(256) lu_panel_gather_kernel(const double* __restrict__ A, long n, long j0, int nb, long rows, double* __restrict__ W) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * nb) return;
    const long i = idx / nb; const int c = (int)(idx % nb);
    W[(long)c * rows + i] = A[(j0 + i) * n + j0 + c];
}
__global__ void __launch_bounds__(256) lu_panel_scatter_kernel(double* __restrict__ A, long n, long j0, int nb, long rows, const double* __restrict__ W) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * nb) return;
    const long i = idx / nb; const int c = (int)(idx % nb);
    A[(j0 + i) * n + j0 + c] = W[(long)c * rows + i];
}
// W: [nb][rows] column-major panel; piv[j] = row (panel-relative) exchanged with row j.  Pivot = largest |.|, lowest index on ties.
__global__ void __launch_bounds__(1024) lu_panel_factor_kernel(double* __restrict__ W, long rows, int nb, int* __restrict__ piv, int* __restrict__ fail) {
    __shared__ double s_val[32];
    __shared__ long s_idx[32];
    __shared__ long s_piv;
    __shared__ double s_rowj[LB];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int j = 0; j < nb; ++j) {
        double* colj = W + (long)j * rows;
        double best = -1.0; long bi = j;
        for (long i = j + tid; i < rows; i += 1024) {
            const double v = fabs(colj[i]);
            if (v > best) { best = v; bi = i; }
        }
        for (int o = 16; o >= 1; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
        __syncthreads();
        if (tid == 0) {
            double b = s_val[0]; long ix = s_idx[0];
            for (int w = 1; w < 32; ++w)
                if (s_val[w] > b || (s_val[w] == b && s_idx[w] < ix)) { b = s_val[w]; ix = s_idx[w]; }
            if (!(b > 0.0)) { *fail = 1; ix = j; }
            s_piv = ix;
            piv[j] = (int)ix;
        }
        __syncthreads();
        const long p = s_piv;
        if (tid < nb) {                                        // interchange rows j and p of the panel; keep row j for the update
            double* col = W + (long)tid * rows;
            const double a = col[j], b = col[p];
            col[j] = b; col[p] = a;
            s_rowj[tid] = b;
        }
        __syncthreads();
        const double pivv = s_rowj[j];
        for (long i = j + 1 + tid; i < rows; i += 1024) {
            const double lij = colj[i] / pivv;
            colj[i] = lij;
            for (int c = j + 1; c < nb; ++c) W[(long)c * rows + i] = fma(-lij, s_rowj[c], W[(long)c * rows + i]);
        }
        __syncthreads();
    }
}
// ---- multi-CTA panel factorisation (cooperative launch) ----
// The single-CTA panel above was 62 % of a 12 000^2 solve (1.1 ms per 32-column panel: one SM doing the read-modify-write of a 3 MB panel out of
// L2).  Here up to 128 CTAs each keep a slab of the panel's rows (a row = 32 contiguous doubles of the row-major matrix: no gather / scatter
// passes) in shared memory; per column: local pivot candidates -> grid barrier -> every CTA picks the global pivot (largest |.|, lowest index
// on ties: the rule of the single-CTA kernel, so the factors are bit-identical), the owners publish the pivot row and row j -> grid barrier ->
// interchange and rank-1 update in shared memory.  Two grid barriers per column.
struct LuPanelScratch { double cand_val[128]; long long cand_idx[128]; double row_p[LB]; double row_j[LB]; unsigned barrier; unsigned pad; };

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned nblocks, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nblocks;
        __threadfence();
        atomicAdd(ctr, 1u);
        while (*(volatile unsigned*)ctr < target) { }
        __threadfence();
    }
    __syncthreads();
}

constexpr int LUP_THREADS = 256;
__global__ void __launch_bounds__(LUP_THREADS) lu_panel_factor_mc_kernel(double* __restrict__ A, long n, long j0, int nb, long rows, int rpc,
                                                                          LuPanelScratch* __restrict__ sc, int* __restrict__ piv, int* __restrict__ fail) {
    extern __shared__ double lup_smem[];
    double* slab = lup_smem;                                   // [rpc][LB + 1]
    __shared__ double s_val[LUP_THREADS / 32];
    __shared__ long long s_idx[LUP_THREADS / 32];
    __shared__ double s_rowj[LB];
    __shared__ long long s_piv;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const long r0 = (long)cta * rpc;                           // first panel-relative row of this CTA
    const int nloc = (int)((rows - r0 < rpc) ? ((rows - r0 > 0) ? rows - r0 : 0) : rpc);
    for (int idx = tid; idx < nloc * LB; idx += LUP_THREADS) {
        const int r = idx / LB, c = idx % LB;
        slab[r * (LB + 1) + c] = (c < nb) ? A[(j0 + r0 + r) * n + j0 + c] : 0.0;
    }
    unsigned target = 0;
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        // local pivot candidate among the rows i >= j of this slab
        double best = -1.0; long long bi = -1;
        for (int r = tid; r < nloc; r += LUP_THREADS) {
            const long i = r0 + r;
            if (i < j) continue;
            const double v = fabs(slab[r * (LB + 1) + j]);
            if (v > best) { best = v; bi = i; }
        }
        for (int o = 16; o >= 1; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (ov > best || (ov == best && (bi < 0 || oi < bi)))) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
        __syncthreads();
        if (tid == 0) {
            double b = s_val[0]; long long ix = s_idx[0];
            for (int w = 1; w < LUP_THREADS / 32; ++w)
                if (s_idx[w] >= 0 && (s_val[w] > b || (s_val[w] == b && (ix < 0 || s_idx[w] < ix)))) { b = s_val[w]; ix = s_idx[w]; }
            sc->cand_val[cta] = b; sc->cand_idx[cta] = ix;
        }
        grid_barrier(&sc->barrier, (unsigned)G, target);
        if (wid == 0) {                                        // every CTA picks the global pivot from the candidates
            double b = -1.0; long long ix = -1;
            for (int g = lane; g < G; g += 32) {
                const double v = ((volatile double*)sc->cand_val)[g]; const long long gi = ((volatile long long*)sc->cand_idx)[g];
                if (gi >= 0 && (v > b || (v == b && (ix < 0 || gi < ix)))) { b = v; ix = gi; }
            }
            for (int o = 16; o >= 1; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, b, o);
                const long long oi = __shfl_xor_sync(0xffffffffu, ix, o);
                if (oi >= 0 && (ov > b || (ov == b && (ix < 0 || oi < ix)))) { b = ov; ix = oi; }
            }
            if (!(b > 0.0)) { if (cta == 0 && lane == 0) *fail = 1; ix = j; }
            if (lane == 0) { s_piv = ix; if (cta == 0) piv[j] = (int)ix; }
        }
        __syncthreads();
        const long long p = s_piv;
        // the owners publish the pivot row and row j
        if (p >= r0 && p < r0 + nloc && tid < LB) sc->row_p[tid] = slab[(int)(p - r0) * (LB + 1) + tid];
        if (j >= r0 && j < r0 + nloc && tid < LB) sc->row_j[tid] = slab[(int)(j - r0) * (LB + 1) + tid];
        grid_barrier(&sc->barrier, (unsigned)G, target);
        if (tid < LB) {
            const double vp = ((volatile double*)sc->row_p)[tid];
            s_rowj[tid] = vp;                                  // row j after the interchange
            if (p != j) {
                const double vj = ((volatile double*)sc->row_j)[tid];
                if (p >= r0 && p < r0 + nloc) slab[(int)(p - r0) * (LB + 1) + tid] = vj;
            }
            if (j >= r0 && j < r0 + nloc) slab[(int)(j - r0) * (LB + 1) + tid] = vp;
        }
        __syncthreads();
        const double pivv = s_rowj[j];
        for (int r = tid; r < nloc; r += LUP_THREADS) {
            if (r0 + r <= j) continue;
            double* row = slab + r * (LB + 1);
            const double lij = row[j] / pivv;
            row[j] = lij;
            for (int c = j + 1; c < nb; ++c) row[c] = fma(-lij, s_rowj[c], row[c]);
        }
        __syncthreads();
    }
    for (int idx = tid; idx < nloc * LB; idx += LUP_THREADS) {
        const int r = idx / LB, c = idx % LB;
        if (c < nb) A[(j0 + r0 + r) * n + j0 + c] = slab[r * (LB + 1) + c];
    }
}

// apply the panel's row interchanges to the columns outside the panel (one thread per column: coalesced along the rows) and to rhs
__global__ void __launch_bounds__(256) lu_apply_swaps_kernel(double* __restrict__ A, long n, long j0, int nb, const int* __restrict__ piv, double* __restrict__ rhs) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n) return;
    if (c == n) {                                              // the extra thread: right-hand side
        for (int j = 0; j < nb; ++j) { const long p = j0 + piv[j]; if (p != j0 + j) { const double t = rhs[j0 + j]; rhs[j0 + j] = rhs[p]; rhs[p] = t; } }
        return;
    }
    if (c >= j0 && c < j0 + nb) return;
    for (int j = 0; j < nb; ++j) {
        const long p = j0 + piv[j];
        if (p != j0 + j) { const double t = A[(j0 + j) * n + c]; A[(j0 + j) * n + c] = A[p * n + c]; A[p * n + c] = t; }
    }
}

// U12 = L11^{-1} A12 : one thread per column c >= j0+nb
__global__ void __launch_bounds__(256) lu_trsm_kernel(double* __restrict__ A, long n, int j0, int nb) {
    __shared__ double L11[LB][LB + 1];
    for (int idx = threadIdx.x; idx < nb * nb; idx += 256) L11[idx / nb][idx % nb] = A[(long)(j0 + idx / nb) * n + j0 + idx % nb];
    __syncthreads();
    const long c = (long)j0 + nb + (long)blockIdx.x * 256 + threadIdx.x;
    if (c >= n) return;
    double col[LB];
#pragma unroll
    for (int r = 0; r < LB; ++r) {
        if (r < nb) {
            double v = A[(long)(j0 + r) * n + c];
#pragma unroll
            for (int k = 0; k < LB; ++k) if (k < r) v -= L11[r][k] * col[k];
            col[r] = v;
            A[(long)(j0 + r) * n + c] = v;
        }
    }
}

// solve L U x = rhs (rhs already permuted), single CTA, 32-row blocks; x overwrites rhs
__global__ void __launch_bounds__(1024) lu_solve_kernel(const double* __restrict__ A, long n, double* __restrict__ x) {
    __shared__ double xs[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // forward: unit lower
    for (long i0 = 0; i0 < n; i0 += 32) {
        const long i = i0 + wid;
        double acc = 0.0;
        if (i < n) for (long c = lane; c < i0; c += 32) acc = fma(A[i * n + c], x[c], acc);
        for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) xs[wid] = (i < n) ? x[i] - acc : 0.0;
        __syncthreads();
        if (wid == 0) {
            for (int r = 0; r < 32; ++r) {
                const double xr = xs[r];
                const long ir = i0 + r, il = i0 + lane;
                if (lane > r && il < n && ir < n) xs[lane] -= A[il * n + ir] * xr;
                __syncwarp();
            }
            if (i0 + lane < n) x[i0 + lane] = xs[lane];
        }
        __syncthreads();
    }
    // backward: upper
    const long nblk = (n + 31) / 32;
    for (long bk = nblk - 1; bk >= 0; --bk) {
        const long i0 = bk * 32;
        const long i = i0 + wid;
        const long cstart = i0 + 32;
        double acc = 0.0;
        if (i < n) for (long c = cstart + lane; c < n; c += 32) acc = fma(A[i * n + c], x[c], acc);
        for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) xs[wid] = (i < n) ? x[i] - acc : 0.0;
        __syncthreads();
        if (wid == 0) {
            for (int r = 31; r >= 0; --r) {
                const long ir = i0 + r, il = i0 + lane;
                if (ir < n) {
                    if (lane == r) xs[r] = xs[r] / A[ir * n + ir];
                    __syncwarp();
                    const double xr = xs[r];
                    if (lane < r && il < n) xs[lane] -= A[il * n + ir] * xr;
                }
                __syncwarp();
            }
            if (i0 + lane < n) x[i0 + lane] = xs[lane];
        }
        __syncthreads();
    }
}

// solve (L L^T) x = rhs for the blocked Cholesky factor (row-major, lower) using the stored inverses of the 64x64 diagonal
// blocks, x overwrites rhs.  Column-oriented block substitution over the whole machine: per block column one tiny kernel applies the
// diagonal-block inverse, one grid-wide kernel subtracts the block's contribution from the remaining right-hand side (the single-CTA
// version of this solve cost 5.9 ms per Newton step at n = 3 000: pure latency).
__global__ void __launch_bounds__(64) trsv_diag_kernel(const double* __restrict__ inv, int nb, int transpose, double* __restrict__ xb) {
    __shared__ double t[CB];
    const int tid = threadIdx.x;
    t[tid] = (tid < nb) ? xb[tid] : 0.0;
    __syncthreads();
    if (tid < nb) {
        double acc = 0.0;
        if (!transpose) { for (int c = 0; c <= tid; ++c) acc = fma(inv[tid * CB + c], t[c], acc); }
        else { for (int r = tid; r < nb; ++r) acc = fma(inv[r * CB + tid], t[r], acc); }      // (inv L_bb)^T
        xb[tid] = acc;
    }
}
// forward: x[i] -= L[i, b0 : b0+nb] . x[b0 : b0+nb] for the rows i >= b0 + nb; one warp per row
__global__ void __launch_bounds__(256) trsv_update_fwd_kernel(const double* __restrict__ Lm, long n, long b0, int nb, double* __restrict__ x) {
    const long i = b0 + nb + (((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const double* row = Lm + i * n + b0;
    double acc = 0.0;
    for (int c = lane; c < nb; c += 32) acc = fma(row[c], x[b0 + c], acc);
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) x[i] -= acc;
}
// backward: x[i] -= L[b0 : b0+nb, i]^T . x[b0 : b0+nb] for the columns i < b0; one thread per column (coalesced over i)
__global__ void __launch_bounds__(256) trsv_update_bwd_kernel(const double* __restrict__ Lm, long n, long b0, int nb, double* __restrict__ x) {
    __shared__ double xb[CB];
    if (threadIdx.x < CB) xb[threadIdx.x] = (threadIdx.x < nb) ? x[b0 + threadIdx.x] : 0.0;
    __syncthreads();
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b0) return;
    double acc = 0.0;
#pragma unroll 8
    for (int r = 0; r < nb; ++r) acc = fma(Lm[(b0 + r) * n + i], xb[r], acc);
    x[i] -= acc;
}

int chol_solve(const double* Lm, long n, const double* invdiag, double* x, cudaStream_t st) {
    const long nblk = (n + CB - 1) / CB;
    for (long b = 0; b < nblk; ++b) {                      // L y = rhs
        const long b0 = b * CB;
        const int nb = (int)std::min<long>(CB, n - b0);
        trsv_diag_kernel<<<1, CB, 0, st>>>(invdiag + b * CB * CB, nb, 0, x + b0);
        const long rest = n - b0 - nb;
        if (rest > 0) trsv_update_fwd_kernel<<<(unsigned)cdiv(rest * 32, 256), 256, 0, st>>>(Lm, n, b0, nb, x);
    }
    for (long b = nblk - 1; b >= 0; --b) {                 // L^T x = y
        const long b0 = b * CB;
        const int nb = (int)std::min<long>(CB, n - b0);
        trsv_diag_kernel<<<1, CB, 0, st>>>(invdiag + b * CB * CB, nb, 1, x + b0);
        if (b0 > 0) trsv_update_bwd_kernel<<<(unsigned)cdiv(b0, 256), 256, 0, st>>>(Lm, n, b0, nb, x);
    }
    SC_LAUNCH_CHECK();
    return OK;
}

}  // namespace

// ------------------------------------------------------------------ host drivers --------------------------

int dgemm(int M, int N, int Kd, double alpha, const double* A, long sai, long sak, const double* B, long sbk, long sbj,
          double beta, double* C, long ldc, int lower_only, cudaStream_t st) {
    if (M <= 0 || N <= 0) return OK;
    if (M > 2 * GM && N > 2 * GN) {                        // the O(n^3) products: large tiles
        dim3 gridb((unsigned)cdiv(N, HN), (unsigned)cdiv(M, HM));
        dgemm_big_kernel<<<gridb, 256, 0, st>>>(M, N, Kd, alpha, A, sai, sak, B, sbk, sbj, beta, C, ldc, lower_only);
        SC_LAUNCH_CHECK();
        return OK;
    }
    dim3 grid((unsigned)cdiv(N, GN), (unsigned)cdiv(M, GM));
    dgemm_kernel<<<grid, 256, 0, st>>>(M, N, Kd, alpha, A, sai, sak, B, sbk, sbj, beta, C, ldc, lower_only);
    SC_LAUNCH_CHECK();
    return OK;
}


int gram_assemble(const GpView& gp, double* K, double nugget, int f16_entries, cudaStream_t st) {
    const long phi = 4L * gp.Nd + gp.Nb;
    const int Nc = gp.Nd + gp.Nb;
    dim3 grid((unsigned)cdiv(Nc, GT), (unsigned)cdiv(Nc, GT));
    gram_kernel<<<grid, 256, 0, st>>>(gp, K, phi, nugget, f16_entries);
    SC_LAUNCH_CHECK();
    return OK;
}

// in-place lower Cholesky of row-major A[n x n]; invdiag: [ceil(n/64)][64*64] inverses of the diagonal blocks
int cholesky_lower(double* A, long n, double* invdiag, int* d_fail, cudaStream_t st) {
    SC_CUDA(cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRSM_SMEM));
    SC_CUDA(cudaFuncSetAttribute(diag_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DINV_SMEM));
    for (long j0 = 0; j0 < n; j0 += CB) {
        const int nb = (int)std::min<long>(CB, n - j0);
        chol_diag_kernel<<<1, 256, 0, st>>>(A, n, (int)j0, nb, d_fail);
        SC_LAUNCH_CHECK();
        const long rest = n - j0 - nb;
        if (rest <= 0) break;
        // panel: L21 = A21 L11^-T (fused triangular solve)
        trsm_panel_kernel<<<(unsigned)cdiv(rest, TRSM_ROWS), TRSM_ROWS, TRSM_SMEM, st>>>(A, n, (int)j0, nb, rest);
        SC_LAUNCH_CHECK();
        // trailing: A22 -= L21 L21^T (lower tiles only)
        double* A21 = A + (j0 + nb) * n + j0;
        double* A22 = A + (j0 + nb) * n + (j0 + nb);
        const int rc = dgemm((int)rest, (int)rest, nb, -1.0, A21, n, 1, A21, 1, n, 1.0, A22, n, 1, st);
        if (rc != OK) return rc;
    }
    zero_upper_kernel<<<(unsigned)cdiv(n * n, 256), 256, 0, st>>>(A, n);
    SC_LAUNCH_CHECK();
    diag_inverse_kernel<<<(unsigned)cdiv(n, CB), CB, DINV_SMEM, st>>>(A, n, invdiag);
    SC_LAUNCH_CHECK();
    return OK;
}

// X = L^{-1} (lower, row-major), using the diagonal-block inverses; scratch: [n x n] (only its leading (n/2)^2 corner is used).
// Recursive halving  inv([[A, 0], [C, D]]) = [[A^-1, 0], [-D^-1 C A^-1, D^-1]]  so that almost all of the work is two large
// products per level (the row-panel formulation did n/64 products with M = 64 rows each: 61 ms of the fit at phi = 4 200).
static int tri_inverse_rec(const double* L, long n, const double* invdiag, double* X, double* scratch, long b_lo, long b_hi, cudaStream_t st) {
    if (b_hi - b_lo == 1) {
        const long i0 = b_lo * CB;
        const int nb = (int)std::min<long>(CB, n - i0);
        SC_CUDA(cudaMemcpy2DAsync(X + i0 * n + i0, n * sizeof(double), invdiag + b_lo * CB * CB, CB * sizeof(double),
                                  nb * sizeof(double), nb, cudaMemcpyDeviceToDevice, st));
        return OK;
    }
    const long b_mid = (b_lo + b_hi) / 2;
    int rc = tri_inverse_rec(L, n, invdiag, X, scratch, b_lo, b_mid, st);
    if (rc != OK) return rc;
    rc = tri_inverse_rec(L, n, invdiag, X, scratch, b_mid, b_hi, st);
    if (rc != OK) return rc;
    const long r0 = b_mid * CB, c0 = b_lo * CB;
    const int m = (int)(std::min<long>(b_hi * CB, n) - r0), k = (int)(r0 - c0);
    // T[m x k] = L21 * X11
    rc = dgemm(m, k, k, 1.0, L + r0 * n + c0, n, 1, X + c0 * n + c0, n, 1, 0.0, scratch, n, 0, st);
    if (rc != OK) return rc;
    // X21 = -X22 * T
    return dgemm(m, k, m, -1.0, X + r0 * n + r0, n, 1, scratch, n, 1, 0.0, X + r0 * n + c0, n, 0, st);
}

int tri_inverse_lower(const double* L, long n, const double* invdiag, double* X, double* scratch, cudaStream_t st) {
    SC_CUDA(cudaMemsetAsync(X, 0, (size_t)n * n * sizeof(double), st));
    return tri_inverse_rec(L, n, invdiag, X, scratch, 0, cdiv(n, CB), st);
}

int lu_solve_inplace(double* H, long n, double* rhs, int* d_fail, cudaStream_t st) {
    double* W = nullptr;                                    // column-major panel copy (single-CTA fallback) + pivots
    int* piv = nullptr;
    LuPanelScratch* psc = nullptr;
    { const int prc = ensure_scratch_pool(); if (prc != OK) return prc; }
    SC_CUDA(cudaMallocAsync((void**)&W, (size_t)n * LB * sizeof(double), st));
    SC_CUDA(cudaMallocAsync((void**)&piv, LB * sizeof(int), st));
    SC_CUDA(cudaMallocAsync((void**)&psc, sizeof(LuPanelScratch), st));
    int dev = 0, nsm = 0, coop = 0;
    SC_CUDA(cudaGetDevice(&dev));
    SC_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    SC_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    int rc = OK;
    for (long j0 = 0; j0 < n && rc == OK; j0 += LB) {
        int nb = (int)std::min<long>(LB, n - j0);
        long rows = n - j0;
        // multi-CTA panel: up to min(128, SMs) CTAs, at least 32 rows each; one CTA per SM is co-resident by construction (cooperative launch checks it)
        int G = (int)std::min<long>(std::min(128, nsm), cdiv(rows, 32));
        int rpc = (int)cdiv(rows, G);
        G = (int)cdiv(rows, rpc);
        const size_t slab = (size_t)rpc * (LB + 1) * sizeof(double);
        if (coop && G > 1 && slab <= 160 * 1024) {
            SC_CUDA(cudaMemsetAsync(&psc->barrier, 0, sizeof(unsigned), st));
            SC_CUDA(cudaFuncSetAttribute(lu_panel_factor_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slab));
            double* Hp = H; long nn = n, jj = j0;
            void* args[] = {&Hp, &nn, &jj, &nb, &rows, &rpc, &psc, &piv, &d_fail};
            SC_CUDA(cudaLaunchCooperativeKernel((const void*)lu_panel_factor_mc_kernel, dim3((unsigned)G), dim3(LUP_THREADS), args, slab, st));
        } else {
            const unsigned g = (unsigned)cdiv(rows * nb, 256);
            lu_panel_gather_kernel<<<g, 256, 0, st>>>(H, n, j0, nb, rows, W);
            lu_panel_factor_kernel<<<1, 1024, 0, st>>>(W, rows, nb, piv, d_fail);
            lu_panel_scatter_kernel<<<g, 256, 0, st>>>(H, n, j0, nb, rows, W);
        }
        lu_apply_swaps_kernel<<<(unsigned)cdiv(n + 1, 256), 256, 0, st>>>(H, n, j0, nb, piv, rhs);
        const long rest = n - j0 - nb;
        if (rest <= 0) break;
        lu_trsm_kernel<<<(unsigned)cdiv(rest, 256), 256, 0, st>>>(H, n, (int)j0, nb);
        rc = dgemm((int)rest, (int)rest, nb, -1.0, H + (j0 + nb) * n + j0, n, 1, H + j0 * n + j0 + nb, n, 1,
                   1.0, H + (j0 + nb) * n + j0 + nb, n, 0, st);
    }
    if (rc == OK) lu_solve_kernel<<<1, 1024, 0, st>>>(H, n, rhs);
    cudaFreeAsync(W, st);
    cudaFreeAsync(piv, st);
    cudaFreeAsync(psc, st);
    if (rc != OK) return rc;
    SC_LAUNCH_CHECK();
    return OK;
}

size_t fit_workspace_bytes(int Nd, int Nb) {
    const size_t phi = 4 * (size_t)Nd + Nb, n3 = 3 * (size_t)Nd;
    const size_t nblk = (phi + CB - 1) / CB;
    size_t bytes = 0;
    bytes += phi * phi * 8;            // K / L
    bytes += phi * phi * 8;            // X = L^-1, then reused as H (3N)^2 <= phi^2
    bytes += phi * phi * 8;            // P
    bytes += nblk * CB * CB * 8;       // diagonal-block inverses
    bytes += CB * phi * 8;             // tmp
    bytes += (4 * phi + 4 * n3 + 64) * 8;
    bytes += ((n3 + CB - 1) / CB) * CB * CB * 8;   // diagonal-block inverses of the Newton system
    return bytes + 8192;
}

// The whole fit.  Returns alpha (device, [phi]) and the loss history (host).
int gp_fit_device(const GpView& gp, const double* g_bdy, const double* sol0, int gn_steps, double damping, double tol,
                  double nugget, int f16_entries, void* workspace, size_t ws_bytes, double* alpha_out, double* sol_out,
                  double* loss_hist_host, int* steps_done, cudaStream_t st) {
    const int N = gp.Nd, Nb = gp.Nb;
    const long phi = 4L * N + Nb, n3 = 3L * N;
    SC_REQUIRE(ws_bytes >= fit_workspace_bytes(N, Nb), "fit: workspace too small");
    char* ws = (char*)workspace;
    auto take = [&](size_t b) { char* p = ws; ws += (b + 255) & ~size_t(255); return (double*)p; };
    const long nblk = cdiv(phi, CB);
    double* K = take((size_t)phi * phi * 8);
    double* X = take((size_t)phi * phi * 8);
    double* P = take((size_t)phi * phi * 8);
    double* invd = take((size_t)nblk * CB * CB * 8);
    double* tmp = take((size_t)CB * phi * 8);
    double* b = take(phi * 8);
    double* w = take(phi * 8);
    double* sol = take(n3 * 8);
    double* grad = take(n3 * 8);
    double* rhs = take(n3 * 8);
    double* scal = take(64);
    int* d_fail = (int*)take(64);
    int* d_fail2 = (int*)take(64);
    double* invh = take((size_t)cdiv(n3, CB) * CB * CB * 8);
    double* H = X;                                  // reuse after P is formed
    SC_CUDA(cudaMemsetAsync(d_fail, 0, sizeof(int), st));

    int rc = gram_assemble(gp, K, nugget, f16_entries, st);
    if (rc != OK) return rc;
    rc = cholesky_lower(K, phi, invd, d_fail, st);
    if (rc != OK) return rc;
    int h_fail = 0;
    SC_CUDA(cudaMemcpyAsync(&h_fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, st));
    SC_CUDA(cudaStreamSynchronize(st));
    if (h_fail && f16_entries) {
        // The float16 rounding of the Gram entries (models/GP.py:258) is a symmetric perturbation of spectral norm ~1e-2 at phi = 4 200 --
        // the size of the nugget -- so a near-singular Gram can turn indefinite.  The reference never sees that: its factor comes from the
        // SVD (|K| + nugget I, always positive definite, models/GP.py:260-268).  Retry once with the un-rounded entries (K itself is PSD).
        SC_CUDA(cudaMemsetAsync(d_fail, 0, sizeof(int), st));
        rc = gram_assemble(gp, K, nugget, 0, st);
        if (rc != OK) return rc;
        rc = cholesky_lower(K, phi, invd, d_fail, st);
        if (rc != OK) return rc;
        SC_CUDA(cudaMemcpyAsync(&h_fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, st));
        SC_CUDA(cudaStreamSynchronize(st));
    }
    if (h_fail) {   // mirrors the ValueError of models/GP.py:264-265
        set_error("Cholesky decomposition of K + nugget*I failed (non-positive pivot / NaN)");
        return ERR_NUMERIC;
    }
    rc = tri_inverse_lower(K, phi, invd, X, P, st);      // P is formed afterwards: scratch until then
    if (rc != OK) return rc;
    // P = X^T X (lower tiles, then mirrored)
    rc = dgemm((int)phi, (int)phi, (int)phi, 1.0, X, 1, phi, X, phi, 1, 0.0, P, phi, 1, st);
    if (rc != OK) return rc;
    mirror_lower_kernel<<<(unsigned)cdiv(phi * phi, 256), 256, 0, st>>>(P, phi);
    SC_LAUNCH_CHECK();

    FitDims f{N, Nb, phi, gp.sig2, 1.0 / gp.d + 0.5 * gp.sig2};
    SC_CUDA(cudaMemcpyAsync(sol, sol0, n3 * 8, cudaMemcpyDeviceToDevice, st));
    auto loss_now = [&](double* host_out) -> int {
        build_b_kernel<<<(unsigned)cdiv(phi, 256), 256, 0, st>>>(f, sol, g_bdy, b);
        gemv_kernel<<<(unsigned)cdiv(phi * 32, 256), 256, 0, st>>>(phi, P, b, w);
        dot_kernel<<<1, 256, 0, st>>>(phi, b, w, scal);
        SC_LAUNCH_CHECK();
        SC_CUDA(cudaMemcpyAsync(host_out, scal, 8, cudaMemcpyDeviceToHost, st));
        return OK;
    };
    int nh = 0, steps = 0;
    rc = loss_now(&loss_hist_host[nh++]);
    if (rc != OK) return rc;
    bool chol_gave_up = false;
    for (int it = 0; it < gn_steps; ++it) {
        // b, w are current (computed by loss_now)
        grad_kernel<<<(unsigned)cdiv(n3, 256), 256, 0, st>>>(f, sol, w, grad, rhs);
        dot_kernel<<<1, 256, 0, st>>>(n3, grad, grad, scal + 1);
        SC_LAUNCH_CHECK();
        double g2 = 0.0;
        SC_CUDA(cudaMemcpyAsync(&g2, scal + 1, 8, cudaMemcpyDeviceToHost, st));
        SC_CUDA(cudaStreamSynchronize(st));
        if (std::sqrt(g2) < tol) break;                                  // models/GP.py:521
        hessian_kernel<<<(unsigned)cdiv(n3 * n3, 256), 256, 0, st>>>(f, sol, w, P, damping, H);
        SC_LAUNCH_CHECK();
        // The damped Hessian is symmetric and, near the solution, positive definite: blocked Cholesky first.
        // A non-positive pivot falls back to LU with partial pivoting (what the reference's jnp.linalg.solve does).
        // Once a Hessian of this fit was not SPD the later ones go to LU directly (a failed attempt costs a full factorisation).
        int h_fail2 = chol_gave_up ? 1 : 0;
        if (!chol_gave_up) {
            SC_CUDA(cudaMemsetAsync(d_fail2, 0, sizeof(int), st));
            rc = cholesky_lower(H, n3, invh, d_fail2, st);
            if (rc != OK) return rc;
            SC_CUDA(cudaMemcpyAsync(&h_fail2, d_fail2, sizeof(int), cudaMemcpyDeviceToHost, st));
            SC_CUDA(cudaStreamSynchronize(st));
            if (h_fail2) chol_gave_up = true;
        }
        if (!h_fail2) {
            { const int rcs = chol_solve(H, n3, invh, rhs, st); if (rcs != OK) return rcs; }
            SC_LAUNCH_CHECK();
        } else {
            hessian_kernel<<<(unsigned)cdiv(n3 * n3, 256), 256, 0, st>>>(f, sol, w, P, damping, H);
            SC_LAUNCH_CHECK();
            rc = lu_solve_inplace(H, n3, rhs, d_fail, st);
            if (rc != OK) return rc;
        }
        axpy_kernel<<<(unsigned)cdiv(n3, 256), 256, 0, st>>>(n3, 1.0, rhs, sol);
        SC_LAUNCH_CHECK();
        rc = loss_now(&loss_hist_host[nh++]);
        if (rc != OK) return rc;
        ++steps;
    }
    SC_CUDA(cudaMemcpyAsync(&h_fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, st));
    // alpha = P z with z = b(sol)   (models/GP.py:593-600); b, w are current
    SC_CUDA(cudaMemcpyAsync(alpha_out, w, phi * 8, cudaMemcpyDeviceToDevice, st));
    if (sol_out) SC_CUDA(cudaMemcpyAsync(sol_out, sol, n3 * 8, cudaMemcpyDeviceToDevice, st));
    SC_CUDA(cudaStreamSynchronize(st));
    if (h_fail) { set_error("Newton system is singular (zero pivot in LU)"); return ERR_NUMERIC; }
    *steps_done = steps;
    for (int i = nh; i <= gn_steps; ++i) loss_hist_host[i] = NAN;
    return OK;
}

}  // namespace scasml
