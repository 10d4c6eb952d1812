// FP64 route of the fused surrogate evaluation (sm_100a).
//
// Replaces the reference's per-call XLA fusions for GP.predict / compute_gradient / compute_PDE_loss
// (models/GP.py:630-687, 326-411, 746-769): the [R x phi] kernel rows are never materialised.  One CTA
// owns a 64-point tile, streams centre tiles through shared memory, forms the three distance dot
// products  x.y, x.roll(y), roll(x).y  as a register-tiled FP64 contraction and applies the closed-form
// functionals and the GP weights alpha in the epilogue.
// This is the parity anchor (errors ~1e-13 vs the oracle); the tcgen05 route lives in gp_eval_tc.cu.
#include <algorithm>
#include "gp.cuh"

namespace scasml {

namespace {

constexpr int BM = 64;    // points per CTA
constexpr int BK = 8;     // contraction chunk
constexpr int TM = 4;     // points per thread
constexpr int NT = 256;   // threads: 16 (ty: points) x 16 (tx: centres)

enum XFeat : int { XF_NX = 0, XF_SX = 1, XF_XT = 2, XF_X0 = 3, XF_SXROLL = 4, XF_XI = 5, XF_XIR = 10, XF_STRIDE = 15 };

// CLASS 0: u only (EVAL_U / EVAL_TERMINAL); 1: u + div_x u; 2: PDE residual (u, div_x, lap_x, dt_x)
template <int CLASS, int TN>
__global__ void __launch_bounds__(NT, 1)
eval_f64_kernel(GpView gp, const double* __restrict__ X, long R, int mode,
                double* __restrict__ out0, double* __restrict__ out1,
                double* __restrict__ out2, double* __restrict__ out3, double* __restrict__ part) {
    // part != null (small batches: the top-level u_hat of a u_solve is 1 200 points = 19 CTAs): the centre tiles are dealt out over
    // gridDim.y CTAs per point tile, each leaves its partial sums in part[(y * 4 + quantity) * R + row], eval_f64_finish_kernel adds them
    // in a fixed order
    constexpr int BN = 16 * TN;
    constexpr bool PDE = (CLASS == 2);
    __shared__ double Xs[BK][BM + 2];
    __shared__ double Xr[PDE ? BK : 1][BM + 2];
    __shared__ double Ys[BK][BN + 2];
    __shared__ double Yr[BK][BN + 2];
    __shared__ double xf[BM][XF_STRIDE];
    __shared__ double cf[BN][CF_STRIDE];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int D = gp.D, d = gp.d;
    const long row0 = (long)blockIdx.x * BM;

    // ---- per-point features (one pass over the tile's rows) ----
    {
        const int r = tid >> 2, part = tid & 3;           // 4 threads per point
        const long row = row0 + r;
        double nx = 0.0, sx = 0.0;
        if (row < R) {
            const double* xr = X + row * (long)D;
            for (int i = part; i < D; i += 4) {
                const double v = xr[i];
                nx += v * v;
                if (i < d) sx += v;
            }
        }
        nx += __shfl_xor_sync(0xffffffffu, nx, 1); nx += __shfl_xor_sync(0xffffffffu, nx, 2);
        sx += __shfl_xor_sync(0xffffffffu, sx, 1); sx += __shfl_xor_sync(0xffffffffu, sx, 2);
        if (part == 0) {
            double xt = 0.0, x0 = 0.0;
            if (row < R) { xt = X[row * (long)D + d]; x0 = X[row * (long)D]; }
            xf[r][XF_NX] = nx; xf[r][XF_SX] = sx; xf[r][XF_XT] = xt; xf[r][XF_X0] = x0;
            xf[r][XF_SXROLL] = sx - x0 + xt;
#pragma unroll
            for (int m = 0; m < MC_IDX; ++m) {
                xf[r][XF_XI + m] = (row < R) ? X[row * (long)D + gp.I[m]] : 0.0;
                xf[r][XF_XIR + m] = (row < R) ? X[row * (long)D + gp.I[m] + 1] : 0.0;
            }
        }
    }
    __syncthreads();

    const double a = gp.a, a2 = a * a, a3 = a2 * a, dd = (double)d, inv5 = 1.0 / MC_IDX;
    double accU[TM], accG[TM], accL[TM], accT[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) { accU[i] = 0.0; accG[i] = 0.0; accL[i] = 0.0; accT[i] = 0.0; }

    const int ntile_dom = gp.NdPad / BN, ntile = (gp.NdPad + gp.NbPad) / BN;
    for (int t = (int)blockIdx.y; t < ntile; t += (int)gridDim.y) {
        const bool is_dom = t < ntile_dom;
        const int c0 = t * BN;
        __syncthreads();                                   // previous tile's epilogue done with cf / tiles
        for (int idx = tid; idx < BN * CF_STRIDE; idx += NT)
            (&cf[0][0])[idx] = gp.feat[(long)c0 * CF_STRIDE + idx];

        double dot1[TM][TN], dot2[TM][TN], dot3[PDE ? TM : 1][PDE ? TN : 1];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) { dot1[i][j] = 0.0; dot2[i][j] = 0.0; if (PDE) dot3[i][j] = 0.0; }

        for (int k0 = 0; k0 < D; k0 += BK) {
            __syncthreads();
            for (int idx = tid; idx < BM * BK; idx += NT) {
                const int r = idx / BK, kk = idx % BK;
                const long row = row0 + r;
                const int k = k0 + kk;
                double v = 0.0, vr = 0.0;
                if (row < R && k < D) {
                    v = X[row * (long)D + k];
                    if (PDE) vr = X[row * (long)D + ((k + 1 == D) ? 0 : k + 1)];
                }
                Xs[kk][r] = v;
                if (PDE) Xr[kk][r] = vr;
            }
            for (int idx = tid; idx < BN * BK; idx += NT) {
                const int r = idx / BK, kk = idx % BK;
                const int k = k0 + kk;
                double v = 0.0, vr = 0.0;
                if (k < D) {
                    const double* yr = gp.C + (long)(c0 + r) * D;
                    v = yr[k];
                    if (is_dom) vr = yr[(k + 1 == D) ? 0 : k + 1];
                }
                Ys[kk][r] = v;
                Yr[kk][r] = vr;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                double xa[TM], xb[PDE ? TM : 1], ya[TN], yb[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) { xa[i] = Xs[kk][ty * TM + i]; if (PDE) xb[i] = Xr[kk][ty * TM + i]; }
#pragma unroll
                for (int j = 0; j < TN; ++j) { ya[j] = Ys[kk][tx * TN + j]; yb[j] = Yr[kk][tx * TN + j]; }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) {
                        dot1[i][j] = fma(xa[i], ya[j], dot1[i][j]);
                        dot2[i][j] = fma(xa[i], yb[j], dot2[i][j]);
                        if (PDE) dot3[i][j] = fma(xb[i], ya[j], dot3[i][j]);
                    }
            }
        }

        // ---- epilogue: closed-form functionals x alpha ----
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const double* c = cf[tx * TN + j];
            const double ny = c[CF_NY], sy = c[CF_SY], yt = c[CF_YT];
            const double A1 = c[CF_A1], A3 = c[CF_A3], A4 = c[CF_A4], A5 = c[CF_A5];
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const double* p = xf[ty * TM + i];
                const double nn = p[XF_NX] + ny;
                const double S = p[XF_SX] - sy, rt = p[XF_XT] - yt;
                const double k = exp(-0.5 * a * (nn - 2.0 * dot1[i][j]));
                double u = k * (A1 + a * (A4 * rt + A5 * S));
                double g = 0.0, tt = 0.0, lp = 0.0;
                if (CLASS >= 1) g = k * (-a * S * A1 - a2 * rt * S * A4 + (a * dd - a2 * S * S) * A5);
                if (PDE) tt = k * (-a * rt * A1 + (a - a2 * rt * rt) * A4 - a2 * rt * S * A5);
                if (is_dom) {
                    const double ky = exp(-0.5 * a * (nn - 2.0 * dot2[i][j]));
                    double m1 = 0.0, m2 = 0.0;
#pragma unroll
                    for (int m = 0; m < MC_IDX; ++m) {
                        const double ry = p[XF_XI + m] - c[CF_YIR + m];
                        m1 += ry; m2 = fma(ry, ry, m2);
                    }
                    const double MH = a2 * m2 * inv5 - a;
                    const double w3 = ky * A3 * dd;
                    u += w3 * MH;
                    if (CLASS >= 1) {
                        const double Sy = p[XF_SX] - c[CF_SYROLL];
                        g += w3 * (2.0 * a2 * m1 * inv5 + a2 * Sy - a3 * Sy * m2 * inv5);
                    }
                    if (PDE) tt += w3 * (-a * (p[XF_XT] - c[CF_Y0])) * MH;
                }
                if (PDE) {
                    const double kx = exp(-0.5 * a * (nn - 2.0 * dot3[i][j]));
                    double n1 = 0.0, n2 = 0.0, q2 = 0.0;
#pragma unroll
                    for (int m = 0; m < MC_IDX; ++m) {
                        const double rx = p[XF_XIR + m] - c[CF_YI + m];
                        n1 += rx; n2 = fma(rx, rx, n2);
                        const double q = p[XF_XIR + m] - c[CF_YIR + m];
                        q2 = fma(q, q, q2);
                    }
                    const double MHx = a2 * n2 * inv5 - a;
                    const double Sx = p[XF_SXROLL] - sy, rxd = p[XF_X0] - yt;
                    lp = kx * dd * (MHx * (A1 + A4 * a * rxd)
                                    + A5 * (-2.0 * a2 * n1 * inv5 - a2 * Sx + a3 * Sx * n2 * inv5));
                    if (is_dom) {
                        const double A = a2 * q2 - MC_IDX * a;
                        lp += A3 * (dd * dd / (MC_IDX * MC_IDX)) * k * (A * A + 2.0 * MC_IDX * a2 - 4.0 * a3 * q2);
                    }
                }
                accU[i] += u;
                if (CLASS >= 1) accG[i] += g;
                if (PDE) { accL[i] += lp; accT[i] += tt; }
            }
        }
    }

    // ---- reduce over the 16 centre-threads of each point, write ----
#pragma unroll
    for (int i = 0; i < TM; ++i) {
#pragma unroll
        for (int o = 8; o >= 1; o >>= 1) {
            accU[i] += __shfl_xor_sync(0xffffffffu, accU[i], o);
            if (CLASS >= 1) accG[i] += __shfl_xor_sync(0xffffffffu, accG[i], o);
            if (PDE) {
                accL[i] += __shfl_xor_sync(0xffffffffu, accL[i], o);
                accT[i] += __shfl_xor_sync(0xffffffffu, accT[i], o);
            }
        }
    }
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int r = ty * TM + i;
            const long row = row0 + r;
            if (row >= R) continue;
            const double u = accU[i];
            if (part != nullptr) {
                double* pp = part + (size_t)blockIdx.y * 4 * R + row;
                pp[0] = u;
                if (CLASS >= 1) pp[R] = accG[i];
                if (PDE) { pp[2 * R] = accL[i]; pp[3 * R] = accT[i]; }
                continue;
            }
            if (CLASS == 0) {
                if (mode == EVAL_TERMINAL) {
                    const double gx = 1.0 - 1.0 / (1.0 + exp(xf[r][XF_XT] + xf[r][XF_SX]));   // equations.py:259
                    out0[row] = gx - u;
                } else {
                    out0[row] = u;
                }
            } else if (CLASS == 1) {
                out0[row] = u;
                out1[row] = accG[i];
            } else {
                const double s2 = gp.sig2;
                out0[row] = accT[i] + (s2 * u - 1.0 / (double)d - 0.5 * s2) * accG[i] + 0.5 * s2 * accL[i];  // GP.py:767-768
                if (out1) out1[row] = accG[i];
                if (out2) out2[row] = accL[i];
                if (out3) out3[row] = accT[i];
            }
        }
    }
}

// ---- centre features ----
__global__ void centre_features_kernel(GpView gp, const double* __restrict__ alpha, double* __restrict__ feat) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = gp.NdPad + gp.NbPad;
    if (c >= total) return;
    const int D = gp.D, d = gp.d, Nd = gp.Nd, Nb = gp.Nb;
    double* f = feat + (long)c * CF_STRIDE;
    for (int i = 0; i < CF_STRIDE; ++i) f[i] = 0.0;
    const bool dom = c < gp.NdPad;
    const int j = dom ? c : c - gp.NdPad;
    if ((dom && j >= Nd) || (!dom && j >= Nb)) return;
    const double* y = gp.C + (long)c * D;
    double ny = 0.0, sy = 0.0;
    for (int i = 0; i < D; ++i) { ny += y[i] * y[i]; if (i < d) sy += y[i]; }
    f[CF_NY] = ny; f[CF_SY] = sy; f[CF_YT] = y[d]; f[CF_Y0] = y[0]; f[CF_SYROLL] = sy - y[0] + y[d];
    for (int m = 0; m < MC_IDX; ++m) { f[CF_YI + m] = y[gp.I[m]]; f[CF_YIR + m] = y[gp.I[m] + 1]; }
    if (alpha) {
        if (dom) {
            f[CF_A1] = alpha[j];
            f[CF_A3] = alpha[Nd + Nb + j];
            f[CF_A4] = alpha[2 * Nd + Nb + j];
            f[CF_A5] = alpha[3 * Nd + Nb + j];
        } else {
            f[CF_A1] = alpha[Nd + j];
        }
    }
}

// ---- gradient vector (public API only; not on the solver hot path) ----
// stage 1: per-pair coefficients; stage 2: grad = -a (x * rowsum - coef @ centres) + sparse terms.
template <int TN>
__global__ void __launch_bounds__(NT, 1)
grad_coef_kernel(GpView gp, const double* __restrict__ X, long R,
                 double* __restrict__ coefR /*[R][NdPad+NbPad]*/, double* __restrict__ coefL /*[R][NdPad]*/) {
    constexpr int BN = 16 * TN;
    __shared__ double Xs[BK][BM + 2];
    __shared__ double Ys[BK][BN + 2];
    __shared__ double Yr[BK][BN + 2];
    __shared__ double xf[BM][XF_STRIDE];
    __shared__ double cf[BN][CF_STRIDE];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int D = gp.D, d = gp.d;
    const long row0 = (long)blockIdx.x * BM;
    const int c0 = blockIdx.y * BN;
    const bool is_dom = c0 < gp.NdPad;
    const int NcPad = gp.NdPad + gp.NbPad;
    {
        const int r = tid >> 2, part = tid & 3;
        const long row = row0 + r;
        double nx = 0.0, sx = 0.0;
        if (row < R) {
            const double* xr = X + row * (long)D;
            for (int i = part; i < D; i += 4) { const double v = xr[i]; nx += v * v; if (i < d) sx += v; }
        }
        nx += __shfl_xor_sync(0xffffffffu, nx, 1); nx += __shfl_xor_sync(0xffffffffu, nx, 2);
        sx += __shfl_xor_sync(0xffffffffu, sx, 1); sx += __shfl_xor_sync(0xffffffffu, sx, 2);
        if (part == 0) {
            xf[r][XF_NX] = nx; xf[r][XF_SX] = sx;
            xf[r][XF_XT] = (row < R) ? X[row * (long)D + d] : 0.0;
            for (int m = 0; m < MC_IDX; ++m) xf[r][XF_XI + m] = (row < R) ? X[row * (long)D + gp.I[m]] : 0.0;
        }
    }
    for (int idx = tid; idx < BN * CF_STRIDE; idx += NT) (&cf[0][0])[idx] = gp.feat[(long)c0 * CF_STRIDE + idx];
    double dot1[TM][TN], dot2[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) { dot1[i][j] = 0.0; dot2[i][j] = 0.0; }
    for (int k0 = 0; k0 < D; k0 += BK) {
        __syncthreads();
        for (int idx = tid; idx < BM * BK; idx += NT) {
            const int r = idx / BK, kk = idx % BK;
            const long row = row0 + r; const int k = k0 + kk;
            Xs[kk][r] = (row < R && k < D) ? X[row * (long)D + k] : 0.0;
        }
        for (int idx = tid; idx < BN * BK; idx += NT) {
            const int r = idx / BK, kk = idx % BK; const int k = k0 + kk;
            double v = 0.0, vr = 0.0;
            if (k < D) { const double* yr = gp.C + (long)(c0 + r) * D; v = yr[k]; if (is_dom) vr = yr[(k + 1 == D) ? 0 : k + 1]; }
            Ys[kk][r] = v; Yr[kk][r] = vr;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            double xa[TM], ya[TN], yb[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) xa[i] = Xs[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) { ya[j] = Ys[kk][tx * TN + j]; yb[j] = Yr[kk][tx * TN + j]; }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) { dot1[i][j] = fma(xa[i], ya[j], dot1[i][j]); dot2[i][j] = fma(xa[i], yb[j], dot2[i][j]); }
        }
    }
    const double a = gp.a, a2 = a * a, dd = (double)d, inv5 = 1.0 / MC_IDX;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long row = row0 + ty * TM + i;
        if (row >= R) continue;
        const double* p = xf[ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const double* c = cf[tx * TN + j];
            const double nn = p[XF_NX] + c[CF_NY];
            const double k = exp(-0.5 * a * (nn - 2.0 * dot1[i][j]));
            const double S = p[XF_SX] - c[CF_SY], rt = p[XF_XT] - c[CF_YT];
            coefR[row * (long)NcPad + c0 + tx * TN + j] = k * (c[CF_A1] + a * (c[CF_A4] * rt + c[CF_A5] * S));
            if (is_dom) {
                const double ky = exp(-0.5 * a * (nn - 2.0 * dot2[i][j]));
                double m2 = 0.0;
#pragma unroll
                for (int m = 0; m < MC_IDX; ++m) { const double ry = p[XF_XI + m] - c[CF_YIR + m]; m2 = fma(ry, ry, m2); }
                coefL[row * (long)gp.NdPad + c0 + tx * TN + j] = ky * c[CF_A3] * dd * (a2 * m2 * inv5 - a);
            }
        }
    }
}

// grad[row][i] for one (row, i): -a (x_i * sum_j c_j - sum_j c_j y_ji) - a (x_i * sum_j cl_j - sum_j cl_j roll(y_j)_i) + sparse
__global__ void grad_contract_kernel(GpView gp, const double* __restrict__ X, long R,
                                     const double* __restrict__ coefR, const double* __restrict__ coefL,
                                     double* __restrict__ grad) {
    const long row = blockIdx.x;
    const int D = gp.D, d = gp.d, NcPad = gp.NdPad + gp.NbPad, NdPad = gp.NdPad;
    const double a = gp.a, a2 = a * a, dd = (double)d;
    const double* cr = coefR + row * (long)NcPad;
    const double* cl = coefL + row * (long)NdPad;
    const double* x = X + row * (long)D;
    __shared__ double s_sum, s_k4, s_k5;
    // scalar sums: rowsum of coefficients, sum_j a*alpha4*k_j (e_t term), sum_j a*alpha5*k_j (1_s term)
    double ps = 0.0, p4 = 0.0, p5 = 0.0;
    for (int j = threadIdx.x; j < NcPad; j += blockDim.x) {
        ps += cr[j];
        if (j < NdPad) {
            ps += cl[j];
            const double* f = gp.feat + (long)j * CF_STRIDE;
            // recover k_j from coefR is not possible in general -> recompute from the distance
            const double* y = gp.C + (long)j * D;
            double r2 = 0.0;
            for (int i = 0; i < D; ++i) { const double t = x[i] - y[i]; r2 = fma(t, t, r2); }
            const double k = exp(-0.5 * a * r2);
            p4 += f[CF_A4] * k; p5 += f[CF_A5] * k;
        }
    }
    __shared__ double red[3][32];
    for (int o = 16; o >= 1; o >>= 1) {
        ps += __shfl_xor_sync(0xffffffffu, ps, o); p4 += __shfl_xor_sync(0xffffffffu, p4, o); p5 += __shfl_xor_sync(0xffffffffu, p5, o);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) { red[0][wid] = ps; red[1][wid] = p4; red[2][wid] = p5; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int w = 0; w < nw; ++w) { s0 += red[0][w]; s1 += red[1][w]; s2 += red[2][w]; }
        s_sum = s0; s_k4 = s1; s_k5 = s2;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const int ir = (i + 1 == D) ? 0 : i + 1;
        double acc = 0.0;
        for (int j = 0; j < NcPad; ++j) acc = fma(cr[j], gp.C[(long)j * D + i], acc);
        for (int j = 0; j < NdPad; ++j) acc = fma(cl[j], gp.C[(long)j * D + ir], acc);
        double g = -a * (x[i] * s_sum - acc);
        if (i == d) g += a * s_k4; else g += a * s_k5;
        // 1_I term of grad lap_y: sum_j alpha3 d ky (2a^2/5) ry_m, ry_m = x_{I_m} - y_{I_m+1}
        for (int m = 0; m < MC_IDX; ++m) {
            if (gp.I[m] != i) continue;
            double t = 0.0;
            for (int j = 0; j < NdPad; ++j) {
                const double* f = gp.feat + (long)j * CF_STRIDE;
                const double* y = gp.C + (long)j * D;
                double r2 = 0.0;
                for (int q = 0; q < D; ++q) { const double u = x[q] - y[(q + 1 == D) ? 0 : q + 1]; r2 = fma(u, u, r2); }
                const double ky = exp(-0.5 * a * r2);
                t += f[CF_A3] * dd * ky * (2.0 * a2 / MC_IDX) * (x[i] - f[CF_YIR + m]);
            }
            g += t;
        }
        grad[row * (long)D + i] = g;
    }
}

// sums the per-centre-range partial sums of eval_f64_kernel (fixed order) and applies the output formulas
__global__ void __launch_bounds__(128) eval_f64_finish_kernel(GpView gp, const double* __restrict__ X, long R, int mode,
                                                              const double* __restrict__ part, int nsplit,
                                                              double* __restrict__ out0, double* __restrict__ out1,
                                                              double* __restrict__ out2, double* __restrict__ out3) {
    const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= R) return;
    double u = 0.0, g = 0.0, l = 0.0, t = 0.0;
    for (int y = 0; y < nsplit; ++y) {
        const double* pp = part + (size_t)y * 4 * R + row;
        u += pp[0];
        if (mode == EVAL_UG || mode == EVAL_PDE) g += pp[R];
        if (mode == EVAL_PDE) { l += pp[2 * R]; t += pp[3 * R]; }
    }
    if (mode == EVAL_U) {
        out0[row] = u;
    } else if (mode == EVAL_TERMINAL) {
        const double* xr = X + row * (long)gp.D;
        double sx = 0.0;
        for (int i = 0; i < gp.d; ++i) sx += xr[i];
        out0[row] = (1.0 - 1.0 / (1.0 + exp(xr[gp.d] + sx))) - u;                       // equations.py:259
    } else if (mode == EVAL_UG) {
        out0[row] = u; out1[row] = g;
    } else {
        const double s2 = gp.sig2;
        out0[row] = t + (s2 * u - 1.0 / (double)gp.d - 0.5 * s2) * g + 0.5 * s2 * l;    // GP.py:767-768
        if (out1) out1[row] = g;
        if (out2) out2[row] = l;
        if (out3) out3[row] = t;
    }
}

}  // namespace

int launch_eval_f64(const GpView& gp, const double* X, long R, int mode,
                    double* out0, double* out1, double* out2, double* out3, cudaStream_t stream, bool split_small) {
    if (R <= 0) return OK;
    SC_REQUIRE(X && out0, "eval: null pointer");
    SC_REQUIRE(gp.NdPad % CENTRE_PAD == 0 && gp.NbPad % CENTRE_PAD == 0, "eval: centre padding");
    SC_REQUIRE(mode == EVAL_U || mode == EVAL_TERMINAL || mode == EVAL_UG || mode == EVAL_PDE, "eval: unknown mode");
    if (mode == EVAL_UG) SC_REQUIRE(out1, "eval UG: out1 is null");
    // small batches: deal the centre tiles out over several CTAs per point tile so that the launch covers the machine
    int dev = 0, nsm = 0;
    SC_CUDA(cudaGetDevice(&dev));
    SC_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    const long gx = cdiv(R, BM);
    const int ntile = (gp.NdPad + gp.NbPad) / (16 * (mode == EVAL_PDE ? 2 : 4));
    int nsplit = split_small ? (int)std::min<long>(ntile, (2L * nsm) / gx) : 1;
    if (nsplit < 2) nsplit = 1;
    double* part = nullptr;
    { const int prc = ensure_scratch_pool(); if (prc != OK) return prc; }
    if (nsplit > 1) SC_CUDA(cudaMallocAsync((void**)&part, (size_t)nsplit * 4 * R * sizeof(double), stream));
    const dim3 grid((unsigned)gx, (unsigned)nsplit);
    switch (mode) {
        case EVAL_U:
        case EVAL_TERMINAL:
            eval_f64_kernel<0, 4><<<grid, NT, 0, stream>>>(gp, X, R, mode, out0, out1, out2, out3, part);
            break;
        case EVAL_UG:
            eval_f64_kernel<1, 4><<<grid, NT, 0, stream>>>(gp, X, R, mode, out0, out1, out2, out3, part);
            break;
        default:
            eval_f64_kernel<2, 2><<<grid, NT, 0, stream>>>(gp, X, R, mode, out0, out1, out2, out3, part);
            break;
    }
    int rc = OK;
    if (cudaGetLastError() != cudaSuccess) { rc = ERR_CUDA; set_error("eval_f64_kernel launch failed"); }
    if (part != nullptr) {
        if (rc == OK) {
            eval_f64_finish_kernel<<<(unsigned)cdiv(R, 128), 128, 0, stream>>>(gp, X, R, mode, part, nsplit, out0, out1, out2, out3);
            if (cudaGetLastError() != cudaSuccess) { rc = ERR_CUDA; set_error("eval_f64_finish_kernel launch failed"); }
        }
        cudaFreeAsync(part, stream);
    }
    return rc;
}

size_t gradient_scratch_bytes(const GpView& gp, long R) {
    return (size_t)R * (size_t)(gp.NdPad + gp.NbPad + gp.NdPad) * sizeof(double);
}

int launch_gradient_f64(const GpView& gp, const double* X, long R, double* grad,
                        double* scratch, size_t scratch_bytes, cudaStream_t stream) {
    if (R <= 0) return OK;
    SC_REQUIRE(scratch_bytes >= gradient_scratch_bytes(gp, R), "gradient: scratch too small");
    double* coefR = scratch;
    double* coefL = scratch + (size_t)R * (gp.NdPad + gp.NbPad);
    SC_CUDA(cudaMemsetAsync(scratch, 0, gradient_scratch_bytes(gp, R), stream));
    const dim3 grid((unsigned)cdiv(R, BM), (unsigned)((gp.NdPad + gp.NbPad) / 64));
    grad_coef_kernel<4><<<grid, NT, 0, stream>>>(gp, X, R, coefR, coefL);
    SC_LAUNCH_CHECK();
    grad_contract_kernel<<<(unsigned)R, 128, 0, stream>>>(gp, X, R, coefR, coefL, grad);
    SC_LAUNCH_CHECK();
    return OK;
}

int build_centre_features(const GpView& gp, const double* alpha, double* feat_out, cudaStream_t stream) {
    const int total = gp.NdPad + gp.NbPad;
    centre_features_kernel<<<(unsigned)cdiv(total, 128), 128, 0, stream>>>(gp, alpha, feat_out);
    SC_LAUNCH_CHECK();
    return OK;
}

}  // namespace scasml
