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

// ------------------------------------------------------------------ Cholesky ------------------------------
constexpr int CB = 64;

// factor the nb x nb (nb <= 64) diagonal block at A[j0][j0] in place (lower), write its inverse to invL[64][64]
__global__ void __launch_bounds__(256) chol_diag_kernel(double* __restrict__ A, long n, int j0, int nb,
                                                        double* __restrict__ invL, int* __restrict__ fail) {
    __shared__ double Ls[CB][CB + 1];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < CB * CB; idx += 256) {
        const int r = idx / CB, c = idx % CB;
        Ls[r][c] = (r < nb && c < nb && c <= r) ? A[(long)(j0 + r) * n + j0 + c] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
        if (tid == 0) {
            const double p = Ls[j][j];
            if (!(p > 0.0)) { *fail = 1; Ls[j][j] = 1.0; } else Ls[j][j] = sqrt(p);
        }
        __syncthreads();
        const double ljj = Ls[j][j];
        for (int r = j + 1 + tid; r < nb; r += 256) Ls[r][j] /= ljj;
        __syncthreads();
        for (int idx = tid; idx < (nb - j - 1) * (nb - j - 1); idx += 256) {
            const int r = j + 1 + idx / (nb - j - 1), c = j + 1 + idx % (nb - j - 1);
            if (c <= r) Ls[r][c] -= Ls[r][j] * Ls[c][j];
        }
        __syncthreads();
    }
    // inverse by forward substitution, one column per thread (column c of invL is private to thread c)
    if (tid < CB) {
        const int c = tid;
        for (int r = 0; r < CB; ++r) {
            double s = (r == c) ? 1.0 : 0.0;
            for (int k = c; k < r; ++k) s -= Ls[r][k] * invL[k * CB + c];
            invL[r * CB + c] = (r < c) ? 0.0 : s / Ls[r][r];
        }
    }
    for (int idx = tid; idx < CB * CB; idx += 256) {
        const int r = idx / CB, c = idx % CB;
        if (r < nb && c < nb) A[(long)(j0 + r) * n + j0 + c] = (c <= r) ? Ls[r][c] : 0.0;
    }
}

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
constexpr int LB = 32;

// factor panel columns [j0, j0+nb) of row-major A[n x n]; row swaps applied to whole rows and to rhs
__global__ void __launch_bounds__(1024) lu_panel_kernel(double* __restrict__ A, long n, int j0, int nb,
                                                        double* __restrict__ rhs, int* __restrict__ fail) {
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_piv;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int j = j0; j < j0 + nb; ++j) {
        double best = -1.0; int bi = j;
        for (long i = j + tid; i < n; i += 1024) {
            const double v = fabs(A[i * n + j]);
            if (v > best) { best = v; bi = (int)i; }
        }
        for (int o = 16; o >= 1; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
        __syncthreads();
        if (tid == 0) {
            double b = s_val[0]; int ix = s_idx[0];
            for (int w = 1; w < 32; ++w)
                if (s_val[w] > b || (s_val[w] == b && s_idx[w] < ix)) { b = s_val[w]; ix = s_idx[w]; }
            if (!(b > 0.0)) { *fail = 1; ix = j; }
            s_piv = ix;
            if (ix != j) { const double t = rhs[j]; rhs[j] = rhs[ix]; rhs[ix] = t; }
        }
        __syncthreads();
        const int p = s_piv;
        if (p != j)
            for (long c = tid; c < n; c += 1024) {
                const double t = A[(long)j * n + c]; A[(long)j * n + c] = A[(long)p * n + c]; A[(long)p * n + c] = t;
            }
        __syncthreads();
        const double piv = A[(long)j * n + j];
        const int w = j0 + nb - j - 1;               // remaining panel columns
        // each warp takes rows; lanes take panel columns (coalesced along the row)
        for (long i = j + 1 + wid; i < n; i += 32) {
            double lij = 0.0;
            if (lane == 0) { lij = A[i * n + j] / piv; A[i * n + j] = lij; }
            lij = __shfl_sync(0xffffffffu, lij, 0);
            if (lane < w) A[i * n + j + 1 + lane] -= lij * A[(long)j * n + j + 1 + lane];
        }
        __syncthreads();
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

// solve (L L^T) x = rhs for the blocked Cholesky factor (row-major, lower) using the stored inverses of the
// 64x64 diagonal blocks; single CTA, x overwrites rhs.  Used for the SPD Newton systems.
__global__ void __launch_bounds__(256) chol_solve_kernel(const double* __restrict__ Lm, long n,
                                                          const double* __restrict__ invdiag, double* __restrict__ x) {
    __shared__ double tvec[CB];
    __shared__ double part[4][CB];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long nblk = (n + CB - 1) / CB;
    // forward: L y = rhs
    for (long b = 0; b < nblk; ++b) {
        const long b0 = b * CB;
        const int nb = (int)((n - b0 < CB) ? (n - b0) : CB);
        for (int r = wid; r < nb; r += 8) {
            const double* row = Lm + (b0 + r) * n;
            double acc = 0.0;
            for (long j = lane; j < b0; j += 32) acc = fma(row[j], x[j], acc);
            for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) tvec[r] = x[b0 + r] - acc;
        }
        __syncthreads();
        const double* inv = invdiag + b * CB * CB;
        if (tid < nb) {
            double acc = 0.0;
            for (int c = 0; c <= tid; ++c) acc = fma(inv[tid * CB + c], tvec[c], acc);
            x[b0 + tid] = acc;
        }
        __syncthreads();
    }
    // backward: L^T x = y
    for (long b = nblk - 1; b >= 0; --b) {
        const long b0 = b * CB;
        const int nb = (int)((n - b0 < CB) ? (n - b0) : CB);
        const int g = tid >> 6, c = tid & 63;
        double acc = 0.0;
        if (c < nb) for (long i = b0 + CB + g; i < n; i += 4) acc = fma(Lm[i * n + b0 + c], x[i], acc);
        part[g][c] = acc;
        __syncthreads();
        if (tid < nb) tvec[tid] = x[b0 + tid] - (part[0][tid] + part[1][tid] + part[2][tid] + part[3][tid]);
        __syncthreads();
        const double* inv = invdiag + b * CB * CB;
        if (tid < nb) {
            double a2 = 0.0;
            for (int r = tid; r < nb; ++r) a2 = fma(inv[r * CB + tid], tvec[r], a2);   // (inv L_bb)^T
            x[b0 + tid] = a2;
        }
        __syncthreads();
    }
}

}  // namespace

// ------------------------------------------------------------------ host drivers --------------------------

int dgemm(int M, int N, int Kd, double alpha, const double* A, long sai, long sak, const double* B, long sbk, long sbj,
          double beta, double* C, long ldc, int lower_only, cudaStream_t st) {
    if (M <= 0 || N <= 0) return OK;
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
    for (long j0 = 0; j0 < n; j0 += CB) {
        const int nb = (int)std::min<long>(CB, n - j0);
        double* inv = invdiag + (j0 / CB) * CB * CB;
        chol_diag_kernel<<<1, 256, 0, st>>>(A, n, (int)j0, nb, inv, d_fail);
        SC_LAUNCH_CHECK();
        const long rest = n - j0 - nb;
        if (rest <= 0) break;
        double* A21 = A + (j0 + nb) * n + j0;
        // panel: L21 = A21 * inv(L11)^T   (triangular solve fused as a product with the block inverse)
        int rc = dgemm((int)rest, nb, nb, 1.0, A21, n, 1, inv, 1, CB, 0.0, A21, n, 0, st);
        if (rc != OK) return rc;
        // trailing: A22 -= L21 L21^T (lower tiles only)
        double* A22 = A + (j0 + nb) * n + (j0 + nb);
        rc = dgemm((int)rest, (int)rest, nb, -1.0, A21, n, 1, A21, 1, n, 1.0, A22, n, 1, st);
        if (rc != OK) return rc;
    }
    zero_upper_kernel<<<(unsigned)cdiv(n * n, 256), 256, 0, st>>>(A, n);
    SC_LAUNCH_CHECK();
    return OK;
}

// X = L^{-1} (lower, row-major), using the diagonal-block inverses; tmp: [64 x n]
int tri_inverse_lower(const double* L, long n, const double* invdiag, double* X, double* tmp, cudaStream_t st) {
    SC_CUDA(cudaMemsetAsync(X, 0, (size_t)n * n * sizeof(double), st));
    for (long i0 = 0; i0 < n; i0 += CB) {
        const int nb = (int)std::min<long>(CB, n - i0);
        const double* inv = invdiag + (i0 / CB) * CB * CB;
        // diagonal block: X[i0:i0+nb, i0:i0+nb] = inv(L_ii)
        SC_CUDA(cudaMemcpy2DAsync(X + i0 * n + i0, n * sizeof(double), inv, CB * sizeof(double),
                                  nb * sizeof(double), nb, cudaMemcpyDeviceToDevice, st));
        if (i0 == 0) continue;
        // tmp[nb x i0] = L[i0:i0+nb, 0:i0] * X[0:i0, 0:i0]
        int rc = dgemm(nb, (int)i0, (int)i0, 1.0, L + i0 * n, n, 1, X, n, 1, 0.0, tmp, n, 0, st);
        if (rc != OK) return rc;
        // X[i0:i0+nb, 0:i0] = -inv * tmp
        rc = dgemm(nb, (int)i0, nb, -1.0, inv, CB, 1, tmp, n, 1, 0.0, X + i0 * n, n, 0, st);
        if (rc != OK) return rc;
    }
    return OK;
}

int lu_solve_inplace(double* H, long n, double* rhs, int* d_fail, cudaStream_t st) {
    for (long j0 = 0; j0 < n; j0 += LB) {
        const int nb = (int)std::min<long>(LB, n - j0);
        lu_panel_kernel<<<1, 1024, 0, st>>>(H, n, (int)j0, nb, rhs, d_fail);
        SC_LAUNCH_CHECK();
        const long rest = n - j0 - nb;
        if (rest <= 0) break;
        lu_trsm_kernel<<<(unsigned)cdiv(rest, 256), 256, 0, st>>>(H, n, (int)j0, nb);
        SC_LAUNCH_CHECK();
        int rc = dgemm((int)rest, (int)rest, nb, -1.0, H + (j0 + nb) * n + j0, n, 1, H + j0 * n + j0 + nb, n, 1,
                       1.0, H + (j0 + nb) * n + j0 + nb, n, 0, st);
        if (rc != OK) return rc;
    }
    lu_solve_kernel<<<1, 1024, 0, st>>>(H, n, rhs);
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
    if (h_fail) {   // mirrors the ValueError of models/GP.py:264-265
        set_error("Cholesky decomposition of K + nugget*I failed (non-positive pivot / NaN)");
        return ERR_NUMERIC;
    }
    rc = tri_inverse_lower(K, phi, invd, X, tmp, st);
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
        SC_CUDA(cudaMemsetAsync(d_fail2, 0, sizeof(int), st));
        rc = cholesky_lower(H, n3, invh, d_fail2, st);
        if (rc != OK) return rc;
        int h_fail2 = 0;
        SC_CUDA(cudaMemcpyAsync(&h_fail2, d_fail2, sizeof(int), cudaMemcpyDeviceToHost, st));
        SC_CUDA(cudaStreamSynchronize(st));
        if (!h_fail2) {
            chol_solve_kernel<<<1, 256, 0, st>>>(H, n3, invh, rhs);
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
