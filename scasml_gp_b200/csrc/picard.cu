// Flattened multilevel-Picard recursion on the device (sm_100a).
//
// The reference walks the Picard tree depth-first in Python, one batch of XLA ops per uz_solve call
// (solvers/ScaSML.py:149-305; 473 calls at n = rho = 4).  Here the tree is enumerated once on the host
// (PicardPlan::build), every call of one level is packed into a single level-wise batch, and a level costs
//   row records      : parent point / global id / time of every row, once (row_setup_kernel)
//   sampler kernels  : Philox increments + path/terminal points (sample_terminal_kernel, sample_paths_kernel)
//   evaluation       : fused surrogate kernels over the level's whole point buffer (gp_eval*.cu)
//   reduction        : per-point weights (point_weights_kernel), then warp-per-row Monte-Carlo means, clip (reduce_rows_kernel)
// Levels are sampled top-down (children's rows are the parents' sample points) and reduced bottom-up.
#include <algorithm>
#include <cstring>
#include "picard.cuh"
#include "gp_tc.cuh"

namespace scasml {

// ------------------------------------------------------------------ plan (host) --------------------------

int PicardPlan::q_of(int L, int l) const { return p_.variant == 0 ? p_.Qrow[L - l - 1] : 1; }
int PicardPlan::mcf_of(int L, int l) const {
    if (p_.variant == 0) return p_.Mfrow[L - l - 1];
    long long v = 1;
    for (int i = 0; i < L - l; ++i) v *= p_.M;
    return (int)v;
}
int PicardPlan::mcg_of(int L) const {
    if (p_.variant == 0) return p_.Mgrow[L];
    long long v = 1;
    for (int i = 0; i < L; ++i) v *= p_.M;
    return (int)v;
}
long long PicardPlan::owned(long long units, bool strided) const {
    if (!strided) return units;
    return units > p_.rank ? (units - p_.rank + p_.world - 1) / p_.world : 0;
}

namespace {
struct Mult { long long v; };
}

// mult = rows per test point in the unsharded tree (for the reference-equivalent sample-point count)
static void level0_counters(const PicardParams& p, PicardStats& st, long long mult, int mcg0) {
    st.eval_counter += (p.scasml ? 1 : 0) + mcg0;          // g() call + "+= MC_g" (solvers/ScaSML.py:59,205)
    st.sample_points += mult * mcg0;                        // drawn and discarded by the reference (:190-219)
}

int PicardPlan::add_call(int level, long long nrows, int parent, int pl, int pk) {
    const int idx = (int)calls_.size();
    calls_.emplace_back();
    {
        CallRec& r = calls_[idx];
        std::memset(&r, 0, sizeof(r));
        r.level = level; r.nrows = nrows; r.parent = parent; r.pl = pl; r.pk = pk;
        r.rowbase = levels_[level].NR;
        for (int i = 0; i < MAXLK; ++i) { r.child[i][0] = -1; r.child[i][1] = -1; }
    }
    levels_[level].NR += nrows;
    levels_[level].calls.push_back(idx);
    return idx;
}

int PicardPlan::build(const PicardParams& p, long B) {
    p_ = p; B_ = B;
    SC_REQUIRE(p.n >= 0 && p.n <= MAX_LEVEL, "picard: level n out of range");
    SC_REQUIRE(p.d >= 1, "picard: d");
    SC_REQUIRE(p.world >= 1 && p.rank >= 0 && p.rank < p.world, "picard: rank/world");
    SC_REQUIRE(B >= 0, "picard: B");
    if (p.variant == 0) {
        SC_REQUIRE(p.qmax >= 1 && p.qmax <= MAX_Q, "picard: qmax out of range");
        for (int i = 0; i < p.n; ++i) {
            SC_REQUIRE(p.Qrow[i] >= 1 && p.Qrow[i] <= p.qmax, "picard: Q entry out of range");
            SC_REQUIRE(p.Mfrow[i] >= 1, "picard: Mf entry");
        }
        for (int i = 0; i <= p.n; ++i) SC_REQUIRE(p.Mgrow[i] >= 1, "picard: Mg entry");
    } else {
        SC_REQUIRE(p.M >= 1, "picard: M");
    }
    calls_.clear();
    levels_.assign(MAX_LEVEL + 1, LevelRec());
    stats_ = PicardStats();
    keyctr_ = p.key_counter;
    if (p.n == 0) {                       // uz_solve(0, ...) returns zeros after the discarded terminal work
        level0_counters(p_, stats_, 1, mcg_of(0));
        ws_bytes_ = 256;
        return OK;
    }

    // depth-first enumeration in the reference's execution order (key assignment = order of random.split)
    struct Frame { int call; long long mult; };
    // recursion via explicit lambda
    struct Rec {
        PicardPlan* self;
        void go(int idx, long long mult) {
            PicardPlan& P = *self;
            const PicardParams& p = P.p_;
            const int L = P.calls_[idx].level;
            const long long nrows = P.calls_[idx].nrows;
            const int mcg = P.mcg_of(L);
            const bool strided = (L == p.n) && p.world > 1;
            P.stats_.eval_counter += (p.scasml ? 1 : 0) + mcg;
            P.stats_.sample_points += mult * mcg;
            P.stats_.n_calls += 1;
            for (int l = 0; l < L; ++l) {
                const int q = P.q_of(L, l), mcf = P.mcf_of(L, l);
                const long long npts = P.owned(nrows * mcf, strided);
                for (int k = 0; k < q; ++k) {
                    const int lk = l * MAX_Q + k;
                    if (p.variant == 0) P.calls_[idx].key[lk] = P.keyctr_++;
                    P.stats_.sample_points += mult * mcf;
                    const long long fcount = (p.variant == 0) ? mcf : (p.scasml ? mcg : mcf);  // quirk A.3-7
                    for (int role = 0; role < 2; ++role) {
                        const int cl = l - role;                 // child level: l, then l-1
                        if (role == 1 && l == 0) break;
                        if (cl >= 1) {
                            const int ci = P.add_call(cl, npts, idx, l, k);
                            P.calls_[idx].child[lk][role] = ci;
                            go(ci, mult * mcf);
                        } else {
                            level0_counters(p, P.stats_, mult * mcf, P.mcg_of(0));
                        }
                        P.stats_.eval_counter += (p.scasml ? 1 : 0) + fcount;   // f() call + "+= MC_f"
                    }
                }
            }
        }
    } rec{this};
    const int top = add_call(p.n, B, -1, 0, 0);
    rec.go(top, 1);
    stats_.keys_used = (long long)(keyctr_ - p.key_counter);

    // layout
    const int D = p.d + 1;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    long long executed = 0;
    for (int L = 1; L <= p.n; ++L) {
        LevelRec& lv = levels_[L];
        const bool strided = (L == p.n) && p.world > 1;
        lv.NT = owned(lv.NR * mcg_of(L), strided);
        long long pt = 0;
        lv.term_off = pt; pt += lv.NT;
        lv.ug_off = pt;
        for (int l = 1; l < L; ++l) {
            lv.NP[l] = owned(lv.NR * mcf_of(L, l), strided);
            for (int k = 0; k < q_of(L, l); ++k) { lv.set_off[l * MAX_Q + k] = pt; pt += lv.NP[l]; }
        }
        lv.n_ug = pt - lv.ug_off;
        lv.pde_off = pt;
        // MLP never reads its level-0 step points (f(x,0,0) = 0 and the level-0 children return zeros)
        lv.NP[0] = p.scasml ? owned(lv.NR * mcf_of(L, 0), strided) : 0;
        for (int k = 0; k < q_of(L, 0); ++k) { lv.set_off[k] = pt; pt += lv.NP[0]; }
        lv.n_pde = pt - lv.pde_off;
        lv.npoints = pt;
        executed += pt;
        lv.off_P = take((size_t)pt * D * sizeof(double));
        lv.off_gid = take((size_t)pt * sizeof(long long));
        // operand records of the tcgen05 evaluation kernel (written by the samplers; unused on the FP64 route)
        lv.off_rec = take((p.scasml && tc_rec_nstep(D) > 0) ? (size_t)pt * tc_rec_bytes(tc_rec_nstep(D)) : 0);
        lv.off_recf = take((p.scasml && tc_rec_nstep(D) > 0) ? (size_t)pt * TC_REC_NFEAT * sizeof(double) : 0);
        lv.off_rows = take((size_t)lv.NR * sizeof(RowRec));
        lv.off_ev0 = take((size_t)pt * sizeof(double));
        lv.off_ev1 = take((size_t)pt * sizeof(double));
        lv.off_us = take((size_t)lv.NR * 2 * sizeof(double));
        lv.off_calls = take(lv.calls.size() * sizeof(CallDev));
        lv.off_rowbase = take(lv.calls.size() * sizeof(long long));
        lv.off_lvdev = take(sizeof(LevelDev));
    }
    stats_.executed_points = executed;
    ws_bytes_ = off + 256;
    return OK;
}

// ------------------------------------------------------------------ kernels -------------------------------

namespace {

__device__ __forceinline__ int find_call(const long long* __restrict__ rowbase, int ncalls, long long R) {
    int lo = 0, hi = ncalls - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (rowbase[mid] <= R) lo = mid; else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ int find_call(const LevelDev& lv, long long R) { return find_call(lv.rowbase, lv.ncalls, R); }

__device__ __forceinline__ long long ceil_div_pos(long long a, long long b) { return a <= 0 ? 0 : (a + b - 1) / b; }

// Row records of a level: one thread per row resolves (call, parent point, global id, time) once; the samplers and the
// reduction then start from a single 32-byte load.
__global__ void __launch_bounds__(256) row_setup_kernel(LevelDev lv) {
    const long long R = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (R >= lv.NR) return;
    const int ci = find_call(lv, R);
    const CallDev& c = lv.calls[ci];
    const long long i = R - c.rowbase;
    RowRec r;
    r.x = c.xsrc + i * lv.D;
    r.gid = c.gidsrc ? c.gidsrc[i] : lv.gid0 + i;
    r.t = r.x[lv.d];
    r.call = ci; r.pad = 0;
    lv.rows[R] = r;
}

// Sampler kernels.  Persistent CTAs keep the 64 KB inverse-CDF table in shared memory (the 16-bit gathers were the dominant
// L1 wavefront source when the table was read through the global path).  A warp owns chunks of 32 consecutive points:
//   phase 1  lane <-> point: row record, flat normal index, step scalars (one load chain per 32 points);
//   phase 2  two points at a time: each half-warp generates the Philox blocks of one point into the warp's buffer, then all
//            32 lanes walk the coordinates of a point (lane <-> coordinate j, j + 32, ...): 16-bit chunk -> table -> point,
//            so every store instruction writes 256 contiguous bytes of the level's point buffer.
constexpr int SMP_MAX_WARPS = 16;
constexpr int SMP_TABLE_BYTES = 65536;

__device__ __forceinline__ void load_normal_table(__half* stab, const __half* __restrict__ gtab) {
    const uint4* src = (const uint4*)gtab;
    uint4* dst = (uint4*)stab;
    for (int i = threadIdx.x; i < SMP_TABLE_BYTES / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
}
// same map as chunk_to_normal (common.cuh): c >= 32768 -> +tab[c - 32768], c < 32768 -> -tab[32767 - c], branch-free
__device__ __forceinline__ double chunk_to_normal_s(const __half* stab, uint32_t c) {
    const uint32_t low = ((c >> 15) & 1u) - 1u;                       // all ones iff c < 32768
    const uint32_t idx = (c ^ low) & 0x7FFFu;
    const unsigned short bits = (unsigned short)(((const unsigned short*)stab)[idx] ^ (low & 0x8000u));
    return (double)__half2float(__ushort_as_half(bits));
}
// X + (drift + sigma (sq N)) with every rounding explicit (no FMA contraction: NumPy's operation order for any sigma)
__device__ __forceinline__ double step_add(double X, double drift, double sigma, double sq, double N) {
    return __dadd_rn(X, __dadd_rn(drift, __dmul_rn(sigma, __dmul_rn(sq, N))));
}
// predicated 8-byte store (keeps the column loops branch-free)
__device__ __forceinline__ void store_pred(double* p, double v, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.f64 [%0], %1;\n\t}" ::"l"(p), "d"(v), "r"((int)pred) : "memory");
}
__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// half-warp h generates the Philox blocks covering the flat indices [f0, f0 + d) of its point into buf
__device__ __forceinline__ void philox_blocks(uint4* buf, unsigned long long f0, int d, PhiloxKey key, int hl) {
    const unsigned long long B0 = f0 >> 3;
    const int nblk = (int)(((f0 + (unsigned long long)d - 1ull) >> 3) - B0) + 1;
#pragma unroll 1
    for (int b = hl; b < nblk; b += 16) buf[b] = philox4x32_10(B0 + (unsigned long long)b, key);
}

// per-point record of a chunk (phase 1 -> phase 2 through the warp's shared memory: broadcast loads instead of shuffles)
struct PtRec {
    unsigned long long xp;               // parent row
    unsigned long long f0;               // flat index of the point's first normal
    double sq, drift;                    // terminal: sqrt(T-t), mu (T-t);  paths: unused
    int call, pad;
};
constexpr int SMP_WARP_FIXED = 32 * (int)sizeof(PtRec);     // bytes per warp besides the Philox buffer / step scalars

// ---- operand records of the tcgen05 evaluation kernel (gp_tc.cuh), written while a point's coordinates are in registers ----
// Column j of a point (lane <-> column): one 32-bit word (hi | lo << 16), the f16 split of a log2(e) x_j (columns past the row are zero), and
// this lane's share of |x|^2 and of sum x over ALL D columns (the writer takes the time column out of the sum again).  EDGE: the pass may reach
// past the row (only a point's last pass of 32 columns does).  The same arithmetic as gp_eval_tc.cu::rec_image_kernel.
template <bool EDGE>
__device__ __forceinline__ void rec_emit(uint32_t* rw, int j, int NC, double v, bool inrow, float asc, double& nx, double& sx) {
    const double vv = (!EDGE || inrow) ? v : 0.0;
    nx = fma(vv, vv, nx);
    sx += vv;
    const float sv = (float)vv * asc;
    // hi + lo split in FP32 (sv rounded to 24 bits; sv - hi is exact in FP32): |error| <= 2^-22 |sv|
    const float hf = __half2float(__float2half_rn(sv));
    uint32_t w;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(sv - hf), "f"(hf));      // upper half: lo, lower half: hi (hf is exact in f16)
    if (!EDGE || j < NC) rw[j] = w;
}
// columns the sampler's passes do not reach (the k-step count is rounded up to an instantiated one)
__device__ __forceinline__ void rec_zero_tail(uint32_t* rw, int jfirst, int NC) {
    for (int j = jfirst; j < NC; j += 32) rw[j] = 0u;
}
// (|x|^2, sum_{i<d} x_i) of TWO points from the lanes' shares: the xor-16-8-4-2-1 butterfly of each of the four sums, packed -- after the first
// level the lower half-warp carries point 0 and the upper one point 1, after the second the 8-lane groups carry (point, quantity): 7 FP64 adds
// and 14 shuffles for two points instead of 20 and 40 (the FP64 warp reductions were what made sampler-side statistics a loss before).
// tm0 / tm1: the points' time columns, which the lanes' sums include.
__device__ __forceinline__ void rec_reduce2(uint32_t* r0, uint32_t* r1, int NC, int lane, double nx0, double sx0, double nx1, double sx1,
                                            double tm0, double tm1) {
    const bool up = lane >= 16, b3 = (lane & 8) != 0;
    double an = (up ? nx1 : nx0) + __shfl_xor_sync(0xffffffffu, up ? nx0 : nx1, 16);
    double as = (up ? sx1 : sx0) + __shfl_xor_sync(0xffffffffu, up ? sx0 : sx1, 16);
    double v = (b3 ? as : an) + __shfl_xor_sync(0xffffffffu, b3 ? an : as, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    if ((lane & 7) == 0) ((double*)((up ? r1 : r0) + NC))[b3 ? 1 : 0] = b3 ? v - (up ? tm1 : tm0) : v;
}

// Feature blocks of the TWO points a warp has just produced (12 coordinates each: x_0, t, x_I, x_{I+1}).  Lane k < 12 writes feature k of point
// a, lane 16 + k of point b: one coalesced 96-byte store per point.  The coordinate sits in lane (c mod 32), pass (c / 32) of the lane <-> column
// registers va / vb: two shuffles per pass and a select -- no divergence in the column loop, no read-back of the row (both were tried: +1.5 ms).
struct RecFeatLane { int src_lane, src_q; bool is_time; };
__device__ __forceinline__ RecFeatLane rec_feature_lane(const LevelDev& lv, int lane) {
    const int k = lane & 15;
    const int c = lv.rec_col[k < 12 ? k : 0];
    RecFeatLane f; f.src_lane = c & 31; f.src_q = c >> 5; f.is_time = (c == lv.d);
    return f;
}
template <int JP>
__device__ __forceinline__ void rec_features2(double* recfl, int pa, int pb, int lane, const RecFeatLane& f, const double (&va)[JP], const double (&vb)[JP],
                                              double ta, double tb) {
    const bool up = lane >= 16;
    double val = up ? tb : ta;
#pragma unroll
    for (int q = 0; q < JP; ++q) {
        const double a = __shfl_sync(0xffffffffu, va[q], f.src_lane);
        const double b = __shfl_sync(0xffffffffu, vb[q], f.src_lane);
        if (!f.is_time && f.src_q == q) val = up ? b : a;
    }
    if ((lane & 15) < 12) recfl[(up ? pb : pa) * 12 + (lane & 15)] = val;
}

// terminal points X_T = (x + mu (T-t)) + sigma (sqrt(T-t) N)   (solvers/ScaSML.py:190-198)
template <int JP, bool REC>
__global__ void __launch_bounds__(SMP_MAX_WARPS * 32, REC ? 1 : 2) sample_terminal_kernel(LevelDev lv, int nslot, int cpts) {
    extern __shared__ __align__(16) uint8_t smp_smem[];
    __half* stab = (__half*)smp_smem;
    load_normal_table(stab, lv.ntab);
    constexpr bool REG = JP <= 4;                            // parent coordinates prefetched into registers
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, h = lane >> 4, hl = lane & 15;
    uint8_t* wbase = smp_smem + SMP_TABLE_BYTES + (size_t)warp * (SMP_WARP_FIXED + (size_t)2 * nslot * 16);
    PtRec* rec = (PtRec*)wbase;
    uint4* pbuf = (uint4*)(wbase + SMP_WARP_FIXED);
    const unsigned short* pch = (const unsigned short*)pbuf;
    const int d = lv.d, D = lv.D, MCg = lv.MCg;
    const double sigma = lv.sigma, T = lv.T, mu = lv.mu;
    const PhiloxKey key = make_key(0u, 0u, lv.seed);
    const RecFeatLane flane = rec_feature_lane(lv, lane);
    const long long NT = lv.NT, nchunk = (NT + cpts - 1) / cpts;     // cpts <= 32 points per chunk (smaller for small launches)
    for (long long ch = (long long)blockIdx.x * nwarp + warp; ch < nchunk; ch += (long long)gridDim.x * nwarp) {
        const long long s = ch * cpts + lane;
        __syncwarp();                                                // the previous chunk's records have been consumed
        if (lane < cpts && s < NT) {
            const long long u = lv.rank + (long long)lv.world * s;
            const long long R = u / MCg;
            const int m = (int)(u - R * MCg);
            const RowRec rr = lv.rows[R];
            const double Tt = T - rr.t;
            PtRec r;
            r.xp = (unsigned long long)rr.x;
            r.f0 = (unsigned long long)(rr.gid * MCg + m) * (unsigned long long)d;
            r.sq = sqrt(Tt); r.drift = mu * Tt; r.call = 0; r.pad = 0;
            rec[lane] = r;
        }
        __syncwarp();
        const int npt = (int)((NT - ch * cpts < cpts) ? (NT - ch * cpts) : cpts);
        double* const outl = lv.P + (lv.term_off + ch * cpts) * D + lane;      // this lane's column of the chunk's first point
        const int NC = REC ? lv.rec_nstep * 16 : 0, recb = REC ? NC * 4 + 16 : 0;
        uint32_t* const recl = REC ? (uint32_t*)(lv.rec + (size_t)(lv.term_off + ch * cpts) * recb) : nullptr;   // 32-bit words: recb / 4 per point
        const int recw = recb >> 2;
        const float asc = lv.rec_ascale;
        double* const recfl = REC ? lv.recf + (size_t)(lv.term_off + ch * cpts) * 12 : nullptr;
        // two points per iteration; an odd tail repeats its last point (same values written twice) so nothing below is conditional
        for (int pp = 0; pp < npt; pp += 2) {
            const int pi[2] = {pp, (pp + 1 < npt) ? pp + 1 : pp};
            double rnx[2] = {0.0, 0.0}, rsx[2] = {0.0, 0.0};
            double tval[2][REG ? JP : 1];                    // the points' coordinates (lane <-> column), for their feature blocks
            const PtRec ra = rec[pi[0]], rb = rec[pi[1]];
            const double* xs[2] = {(const double*)ra.xp, (const double*)rb.xp};
            double xv[2][REG ? JP : 1];
            if (REG) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                    for (int q = 0; q < JP; ++q) { const int j = lane + 32 * q; xv[hh][q] = (j < d) ? __ldg(xs[hh] + j) : 0.0; }
            }
            philox_blocks(pbuf + h * nslot, h ? rb.f0 : ra.f0, d, key, hl);
            __syncwarp();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const unsigned long long f0p = hh ? rb.f0 : ra.f0;
                const double sqp = hh ? rb.sq : ra.sq, drp = hh ? rb.drift : ra.drift;
                const unsigned short* cp = pch + (size_t)hh * nslot * 8 + (int)(f0p & 7ull);
                double* const dst = outl + pi[hh] * D;               // 32-bit offset: one wide multiply-add
                constexpr int QB = JP < 8 ? JP : 8;                  // column passes per load round (large d: the loads of a round are
                                                                     // issued together, ahead of the round's stores)
#pragma unroll 1
                for (int q0 = 0; q0 < JP; q0 += QB) {
                    if (!REG && 32 * q0 > d) break;                 // warp-uniform: nothing left in this row
                    double xin[QB];
#pragma unroll
                    for (int qq = 0; qq < QB; ++qq) { const int j = lane + 32 * (q0 + qq); xin[qq] = REG ? xv[hh][q0 + qq] : ((j < d) ? __ldg(xs[hh] + j) : 0.0); }
#pragma unroll
                    for (int qq = 0; qq < QB; ++qq) {                // predicated, no branches: lanes past the row compute garbage and drop it
                        const int j = lane + 32 * (q0 + qq);
                        const double N = chunk_to_normal_s(stab, cp[(JP <= 4 || j < d) ? j : 0]);
                        const double val = __dadd_rn(__dadd_rn(xin[qq], drp), __dmul_rn(sigma, __dmul_rn(sqp, N)));
                        store_pred(dst + 32 * (q0 + qq), (j == d) ? T : val, j <= d);
                        if (REC) {
                            if (REG) tval[hh][q0 + qq] = val;
                            if (q0 + qq == JP - 1) rec_emit<true>(recl + pi[hh] * recw, j, NC, (j == d) ? T : val, j <= d, asc, rnx[hh], rsx[hh]);
                            else rec_emit<false>(recl + pi[hh] * recw, j, NC, val, true, asc, rnx[hh], rsx[hh]);
                        }
                    }
                }
                if (REC) rec_zero_tail(recl + pi[hh] * recw, JP * 32 + lane, NC);
            }
            if (REC) {
                rec_reduce2(recl + pi[0] * recw, recl + pi[1] * recw, NC, lane, rnx[0], rsx[0], rnx[1], rsx[1], T, T);
                rec_features2<REG ? JP : 1>(recfl, pi[0], pi[1], lane, flane, tval[0], tval[1], T, T);
            }
            __syncwarp();
        }
    }
}

// interior path points of step set l (all k), solvers/ScaSML.py:220-238 / ScaSML_full_history.py:142-154
template <int JP, bool REC>
__global__ void __launch_bounds__(SMP_MAX_WARPS * 32, REC ? 1 : 2) sample_paths_kernel(LevelDev lv, int l, int nslot, int cpts) {
    extern __shared__ __align__(16) uint8_t smp_smem[];
    __half* stab = (__half*)smp_smem;
    load_normal_table(stab, lv.ntab);
    constexpr bool REG = JP <= 4;                            // running path state in registers; else re-read from the previous step's point
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, h = lane >> 4, hl = lane & 15;
    const int nq = (lv.variant == 0) ? lv.q[l] : 1;
    uint8_t* wbase = smp_smem + SMP_TABLE_BYTES + (size_t)warp * (SMP_WARP_FIXED + (size_t)2 * nslot * 16 + (size_t)nq * 3 * 32 * 8);
    PtRec* rec = (PtRec*)wbase;
    uint4* pbuf = (uint4*)(wbase + SMP_WARP_FIXED);
    const unsigned short* pch = (const unsigned short*)pbuf;
    double* scal = (double*)(wbase + SMP_WARP_FIXED + (size_t)2 * nslot * 16);   // [nq][3][32]: t_k, sqrt(d_k), mu d_k per (step, point of the chunk)
    const int d = lv.d, D = lv.D, MCf = lv.MCf[l];
    const double sigma = lv.sigma, T = lv.T, mu = lv.mu;
    const long long NP = lv.NP[l];
    const PhiloxKey kT = make_key(0u, 0u, lv.seed);
    const RecFeatLane flane = rec_feature_lane(lv, lane);
    const bool quad = lv.variant == 0;
    const long long nchunk = (NP + cpts - 1) / cpts;
    for (long long ch = (long long)blockIdx.x * nwarp + warp; ch < nchunk; ch += (long long)gridDim.x * nwarp) {
        const long long s = ch * cpts + lane;
        __syncwarp();                                                // the previous chunk's records / scalars have been consumed
        if (lane < cpts && s < NP) {
            const long long u = lv.rank + (long long)lv.world * s;
            const long long R = u / MCf;
            const int m = (int)(u - R * MCf);
            const RowRec rr = lv.rows[R];
            const long long pgid = rr.gid * MCf + m;
            PtRec r;
            r.xp = (unsigned long long)rr.x; r.call = rr.call; r.pad = 0; r.sq = 0.0; r.drift = 0.0;
            r.f0 = (unsigned long long)pgid * (unsigned long long)d;
            rec[lane] = r;
            const double t = rr.t;
            if (quad) {
                double tprev = t;
                for (int k = 0; k < nq; ++k) {
                    const int lk = l * MAX_Q + k;
                    const double tk = cloc_of(T, t, lv.cnode[lk]);
                    const double dk = __dsub_rn(tk, tprev);
                    scal[(k * 3 + 0) * 32 + lane] = tk;
                    scal[(k * 3 + 1) * 32 + lane] = sqrt(dk);
                    scal[(k * 3 + 2) * 32 + lane] = mu * dk;
                    lv.gid[lv.set_off[lk] + s] = pgid;
                    tprev = tk;
                }
            } else {
                const double tau = chunk_to_uniform(chunk16((unsigned long long)pgid, kT));
                const double steps = tau * (T - t);
                scal[0 * 32 + lane] = t + steps;
                scal[1 * 32 + lane] = sqrt(steps);
                scal[2 * 32 + lane] = mu * steps;
                lv.gid[lv.set_off[l * MAX_Q] + s] = pgid;
            }
        }
        __syncwarp();
        const int npt = (int)((NP - ch * cpts < cpts) ? (NP - ch * cpts) : cpts);
        const int NC = REC ? lv.rec_nstep * 16 : 0, recb = REC ? NC * 4 + 16 : 0, recw = recb >> 2;
        const float asc = lv.rec_ascale;
        for (int pp = 0; pp < npt; pp += 2) {                        // an odd tail repeats its last point: nothing below is conditional
            const int pi[2] = {pp, (pp + 1 < npt) ? pp + 1 : pp};
            const PtRec ra = rec[pi[0]], rb = rec[pi[1]];
            const double* xs[2] = {(const double*)ra.xp, (const double*)rb.xp};
            double xv[2][REG ? JP : 1];
            if (REG) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                    for (int q = 0; q < JP; ++q) { const int j = lane + 32 * q; xv[hh][q] = (j < d) ? __ldg(xs[hh] + j) : 0.0; }
            }
            const unsigned long long f0h = h ? rb.f0 : ra.f0;
            const unsigned* keys = lv.calls[h ? rb.call : ra.call].key;
#pragma unroll 1
            for (int k = 0; k < nq; ++k) {
                const int lk = l * MAX_Q + k;
                const PhiloxKey key = quad ? make_key(__ldg(keys + lk), 1u, lv.seed) : kT;
                philox_blocks(pbuf + h * nslot, f0h, d, key, hl);
                __syncwarp();
                double* const outl = lv.P + (lv.set_off[lk] + ch * cpts) * D + lane;
                uint32_t* const recl = REC ? (uint32_t*)(lv.rec + (size_t)(lv.set_off[lk] + ch * cpts) * recb) : nullptr;
                double* const recfl = REC ? lv.recf + (size_t)(lv.set_off[lk] + ch * cpts) * 12 : nullptr;
                double rnx[2] = {0.0, 0.0}, rsx[2] = {0.0, 0.0}, rtm[2] = {0.0, 0.0};
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int p = pi[hh];
                    const unsigned long long f0p = hh ? rb.f0 : ra.f0;
                    const double tk = scal[(k * 3 + 0) * 32 + p];
                    if (REC) rtm[hh] = tk;
                    const double sqp = scal[(k * 3 + 1) * 32 + p], drp = scal[(k * 3 + 2) * 32 + p];
                    const unsigned short* cp = pch + (size_t)hh * nslot * 8 + (int)(f0p & 7ull);
                    double* const dst = outl + p * D;
                    const double* prev = (k == 0) ? xs[hh] : lv.P + (lv.set_off[lk - 1] + ch * cpts + p) * D;
                    constexpr int QB = JP < 8 ? JP : 8;              // column passes per load round
#pragma unroll 1
                    for (int q0 = 0; q0 < JP; q0 += QB) {
                        if (!REG && 32 * q0 > d) break;             // warp-uniform: nothing left in this row
                        double xin[QB];
#pragma unroll
                        for (int qq = 0; qq < QB; ++qq) {            // !REG: prev was written by this very thread one step earlier
                            const int j = lane + 32 * (q0 + qq);
                            xin[qq] = REG ? xv[hh][q0 + qq] : ((j < d) ? prev[j] : 0.0);
                        }
#pragma unroll
                        for (int qq = 0; qq < QB; ++qq) {            // predicated, no branches
                            const int j = lane + 32 * (q0 + qq);
                            const double N = chunk_to_normal_s(stab, cp[(JP <= 4 || j < d) ? j : 0]);
                            const double xn = step_add(xin[qq], drp, sigma, sqp, N);
                            if (REG) xv[hh][q0 + qq] = xn;
                            store_pred(dst + 32 * (q0 + qq), (j == d) ? tk : xn, j <= d);
                            if (REC) {
                                if (q0 + qq == JP - 1) rec_emit<true>(recl + p * recw, j, NC, (j == d) ? tk : xn, j <= d, asc, rnx[hh], rsx[hh]);
                                else rec_emit<false>(recl + p * recw, j, NC, xn, true, asc, rnx[hh], rsx[hh]);
                            }
                        }
                    }
                    if (REC) rec_zero_tail(recl + p * recw, JP * 32 + lane, NC);
                }
                if (REC) {
                    rec_reduce2(recl + pi[0] * recw, recl + pi[1] * recw, NC, lane, rnx[0], rsx[0], rnx[1], rsx[1], rtm[0], rtm[1]);
                    rec_features2<REG ? JP : 1>(recfl, pi[0], pi[1], lane, flane, xv[0], xv[1], rtm[0], rtm[1]);
                }
                __syncwarp();
            }
        }
    }
}

// MLP terminal values g(X_T) (solvers/MLP.py:42-55): one warp per terminal point
__global__ void __launch_bounds__(256) mlp_terminal_kernel(LevelDev lv) {
    const long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= lv.NT) return;
    const double* x = lv.P + (lv.term_off + s) * lv.D;
    double acc = 0.0;
    for (int j = lane; j < lv.D; j += 32) acc += x[j];           // t + sum_i x_i
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) lv.ev0[lv.term_off + s] = 1.0 - 1.0 / (1.0 + exp(acc));
}

__device__ __forceinline__ double clip_keep_nan(double v, double c) { return (v < -c) ? -c : ((v > c) ? c : v); }

// One Philox block serves 8 consecutive flat indices: the lanes of a warp generate the blocks covering
// [fstart, fstart + count) once, park them in shared memory, and every lane picks its 16-bit chunk from there.
__device__ __forceinline__ void fill_chunks(uint4* buf, unsigned long long fstart, int count, PhiloxKey key, int lane) {
    const unsigned long long blk0 = fstart >> 3;
    const int nblk = (int)(((fstart + (unsigned long long)count - 1ull) >> 3) - blk0) + 1;
    __syncwarp();
    for (int b = lane; b < nblk; b += 32) buf[b] = philox4x32_10(blk0 + (unsigned long long)b, key);
    __syncwarp();
}
__device__ __forceinline__ uint32_t read_chunk(const uint4* buf, unsigned long long fstart, unsigned long long f) {
    return (uint32_t)((const unsigned short*)buf)[(int)(f - ((fstart >> 3) << 3))];
}

// ---- Monte-Carlo means of one level (solvers/ScaSML.py:211-215,252-284), two kernels -------------------------------------
// The Brownian increments are not regenerated: a sampled point carries them,
//   N = (X_T - x - mu (T-t)) / (sigma sqrt(T-t)),    W_k = (X_k - x - mu (t_k - t)) / sigma      (ScaSML.py:190-194,228-233)
// so  z_j = sum_p c_p X_pj - x_j sum_p c_p - sum_p c_p drift_p  is a weighted row sum of the level's point buffer.
//   (A) point_weights_kernel: one thread per sample point turns its evaluation outputs / child results into the scalar
//       weight c_p of its row in z and its contribution to u, IN PLACE over ev0 / ev1 (fully parallel, no per-row chains);
//   (B) reduce_rows_kernel: one warp per parent row streams the row's points with those weights.
// Degenerate steps (a sampled point equal to its parent: T == t, or tau == 0 in the full-history variant) have weight 0 in (A)
// and fall back to regenerated Philox normals in (B).

// owned sample indices [lo, hi) of row R when unit u = R MC + m belongs to rank u % world
__device__ __forceinline__ void owned_range(long long R, int MC, long long rank, long long world, long long& lo, long long& hi) {
    if (world == 1) { lo = R * MC; hi = lo + MC; return; }
    lo = ceil_div_pos(R * MC - rank, world); hi = ceil_div_pos((R + 1) * MC - rank, world);
}

__global__ void __launch_bounds__(256) point_weights_kernel(LevelDev lv) {
    const long long pt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= lv.npoints) return;
    const double T = lv.T, sig = lv.sigma;
    if (pt >= lv.term_off && pt < lv.term_off + lv.NT) {
        // terminal condition: u += g / MC_g, z += g N / (MC_g delta0)
        const long long s = pt - lv.term_off;
        const long long R = (lv.rank + (long long)lv.world * s) / lv.MCg;
        const double Tt = T - lv.rows[R].t;
        const double delta0 = (lv.variant == 0) ? (Tt + 1e-6) : Tt;    // ScaSML.py:213 vs ScaSML_full_history.py:133
        const double sqT = sqrt(Tt);
        const double gt = lv.ev0[pt];
        const double invT = (1.0 / ((double)lv.MCg * delta0)) / (sig * sqT);
        lv.ev0[pt] = (sqT > 0.0) ? gt * invT : 0.0;
        lv.ev1[pt] = gt / lv.MCg;
        return;
    }
    int l = -1, k = 0;
    long long s = 0;
    for (int ll = 0; ll < lv.L; ++ll) {
        const long long o0 = lv.set_off[ll * MAX_Q], np = lv.NP[ll];
        if (np > 0 && pt >= o0 && pt < o0 + np * lv.q[ll]) { l = ll; k = (int)((pt - o0) / np); s = pt - o0 - (long long)k * np; }
    }
    if (l < 0) return;
    const int lk = l * MAX_Q + k, MCf = lv.MCf[l];
    const long long u = lv.rank + (long long)lv.world * s;
    const long long R = u / MCf;
    const RowRec rr = lv.rows[R];
    const CallDev& c = lv.calls[rr.call];
    const double t = rr.t, Tt = T - t;
    const long long crow = s - c.rowbase * MCf;               // row of this sample inside the child calls
    double y1 = 0.0, y2 = 0.0;
    if (l >= 1) {
        double uh = 0.0, sG = 0.0;
        if (lv.scasml) { uh = lv.ev0[pt]; sG = sig * lv.ev1[pt]; }
        const double* r1 = lv.us[l] + 2 * (c.childbase[lk][0] + crow);
        y1 = lv.scasml ? sig * ((r1[0] + uh) * (sG + r1[1]) - uh * sG) : sig * r1[0] * r1[1];
        if (l >= 2) {
            const double* r2 = lv.us[l - 1] + 2 * (c.childbase[lk][1] + crow);
            y2 = lv.scasml ? sig * ((r2[0] + uh) * (sG + r2[1]) - uh * sG) : sig * r2[0] * r2[1];
        }
    } else if (lv.scasml) {
        y1 = lv.ev0[pt];                                       // PDE residual of the surrogate (ScaSML.py:275)
    }
    double cw, uc;
    if (lv.variant == 0) {
        const double tk = cloc_of(T, t, lv.cnode[lk]);
        const double wk = wloc_of(T, t, lv.wnode[lk]);
        const double dnew = (tk - t) + 1e-6;
        double dadd = dnew, dsub = dnew;
        if (lv.stale_delta) {
            // solvers/MLP.py:201,249,270: the z term of y1 divides by the delta_t left over from the previous (l >= 1) node,
            // level-0 nodes never update it
            const double d0 = Tt + 1e-6;
            if (l == 0) { dadd = d0; dsub = d0; }
            else {
                int pl = l, pk = k - 1;
                if (pk < 0) { pl = l - 1; pk = (pl >= 1) ? lv.q[pl] - 1 : 0; }
                dadd = (pl >= 1) ? (cloc_of(T, t, lv.cnode[pl * MAX_Q + pk]) - t) + 1e-6 : d0;
            }
        }
        uc = (wk / MCf) * (y1 - y2);
        cw = (y1 * (wk / (MCf * dadd)) - y2 * (wk / (MCf * dsub))) / sig;   // z += (y1 / delta_a - y2 / delta_s) w_k / MC_f * W_k
    } else {
        const int m = (int)(u - R * MCf);
        const long long pgid = rr.gid * MCf + m;
        const double tau = chunk_to_uniform(chunk16((unsigned long long)pgid, make_key(0u, 0u, lv.seed)));
        const double steps = tau * Tt;
        uc = Tt * (y1 - y2) / MCf;
        const double yz = uc / sqrt(steps + 1e-6);               // ScaSML_full_history.py:169
        const double sqs = sqrt(steps);
        cw = (sqs > 0.0) ? yz / (sig * sqs) : 0.0;               // N = (X - x - mu steps) / (sigma sqrt(steps))
        if (!(sqs > 0.0)) atomicOr(&lv.rows[R].pad, 1);           // degenerate step: (B) regenerates its normals
    }
    lv.ev0[pt] = cw;
    lv.ev1[pt] = uc;
}

// entry e of row R's sample list [terminal samples | step set (l, k) samples ...] -> point index (pure arithmetic)
struct RowEntries {
    long long sT_lo, nT, total;
};
__device__ __forceinline__ RowEntries row_entries(const LevelDev& lv, long long R) {
    RowEntries E;
    long long hi;
    owned_range(R, lv.MCg, lv.rank, lv.world, E.sT_lo, hi);
    E.nT = hi - E.sT_lo;
    E.total = E.nT;
    for (int l = 0; l < lv.L; ++l) {
        if (lv.NP[l] == 0) continue;               // MLP: f(x, 0, 0) = 0 at level 0, nothing to add (solvers/MLP.py:243)
        long long a, b;
        owned_range(R, lv.MCf[l], lv.rank, lv.world, a, b);
        E.total += (b - a) * lv.q[l];
    }
    return E;
}
// kind: 0 terminal, 1 step set, -1 none; sidx: owned sample index; MC: samples per row of the entry's set
__device__ __forceinline__ long long entry_point(const LevelDev& lv, long long R, const RowEntries& E, long long e, int& kind, long long& sidx, int& MC) {
    kind = -1; sidx = 0; MC = 1;
    long long pt = 0;
    if (e < E.nT) { kind = 0; sidx = E.sT_lo + e; MC = lv.MCg; return lv.term_off + sidx; }
    long long rem = e - E.nT;
    for (int l = 0; l < lv.L; ++l) {
        if (lv.NP[l] == 0) continue;
        long long a, b;
        owned_range(R, lv.MCf[l], lv.rank, lv.world, a, b);
        const long long n = b - a;
        for (int k = 0; k < lv.q[l]; ++k) {
            if (kind < 0 && rem >= 0 && rem < n) { kind = 1; sidx = a + rem; pt = lv.set_off[l * MAX_Q + k] + sidx; MC = lv.MCf[l]; }
            rem -= n;
        }
    }
    return pt;
}

// Rare path of the reduction (kept out of line so that it costs the streaming kernel no registers): the directly accumulated
// z terms  sum_p w_p N_p  of a row's degenerate sample points, with REGENERATED Philox normals -- terminal samples of a row
// with T == t, full-history steps with tau == 0 (flagged in RowRec::pad by point_weights_kernel).  Result: zdir[32 q + lane].
template <int JCH>
__device__ __noinline__ void reduce_direct_terms(const LevelDev& lv, long long R, int jpass, uint4* wbuf, double* zdir_out) {
    const int lane = threadIdx.x & 31;
    const int d = lv.d;
    const RowRec rr = lv.rows[R];
    const double Tt = lv.T - rr.t;
    const bool term_direct = !(sqrt(Tt) > 0.0);
    const PhiloxKey kT = make_key(0u, 0u, lv.seed);
    const int cnt = (d - jpass < 32 * JCH) ? (d - jpass) : 32 * JCH;
    double zdir[JCH];
#pragma unroll
    for (int i = 0; i < JCH; ++i) zdir[i] = 0.0;
    auto direct_normals = [&](unsigned mask, double w, unsigned long long f0) {
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const double wi = __shfl_sync(0xffffffffu, w, src);
            const unsigned long long f0i = __shfl_sync(0xffffffffu, f0, src);
            const unsigned long long fs = f0i + (unsigned long long)jpass;
            fill_chunks(wbuf, fs, cnt, kT, lane);
#pragma unroll
            for (int q = 0; q < JCH; ++q) {
                const int j = jpass + lane + 32 * q;
                if (j < d) zdir[q] = fma(wi, chunk_to_normal(lv.ntab, read_chunk(wbuf, fs, f0i + j)), zdir[q]);
            }
        }
    };
    const RowEntries E = row_entries(lv, R);
    for (long long base = 0; base < E.total; base += 32) {
        const long long e = base + lane;
        const bool valid = e < E.total;
        int kind, MC; long long sidx;
        const long long pt = entry_point(lv, R, E, valid ? e : 0, kind, sidx, MC);
        const double uc = valid ? lv.ev1[pt] : 0.0;
        if (term_direct) {
            const int m = (int)(lv.rank + (long long)lv.world * sidx - R * MC);
            const unsigned tmask = __ballot_sync(0xffffffffu, valid && kind == 0);
            if (tmask) direct_normals(tmask, uc, (unsigned long long)(rr.gid * MC + m) * (unsigned long long)d);
        }
        if (lv.variant == 1) {
            const bool sv = valid && kind == 1;
            const long long pgid = sv ? lv.gid[pt] : 0;
            const double steps = chunk_to_uniform(chunk16((unsigned long long)pgid, kT)) * Tt;
            const unsigned dmask = __ballot_sync(0xffffffffu, sv && !(sqrt(steps) > 0.0));
            if (dmask) direct_normals(dmask, uc / sqrt(steps + 1e-6), (unsigned long long)pgid * (unsigned long long)d);   // ScaSML_full_history.py:169
        }
    }
    // terminal fallback: sum first, scale afterwards: with delta0 == 0 (full history at t == T) the reference's mean(g N) / 0 is +-inf, not
    // NaN.  A row with T == t has Tt == 0, so its full-history step weights uc = Tt (...) are 0 (the reference's 0 * N): scaling them too is harmless.
    if (term_direct) {
        const double delta0 = (lv.variant == 0) ? (Tt + 1e-6) : Tt;
#pragma unroll
        for (int q = 0; q < JCH; ++q) zdir[q] *= 1.0 / delta0;         // the weights already carry 1 / MC_g
    }
#pragma unroll
    for (int q = 0; q < JCH; ++q) zdir_out[32 * q + lane] = zdir[q];
}

template <int JCH, int UNR, int MINB>
__global__ void __launch_bounds__(128, MINB) reduce_rows_kernel(const __grid_constant__ LevelDev lv) {
    __shared__ uint4 wbuf_all[4][4 * JCH + 3];
    __shared__ double zdir_all[4][32 * JCH];
    const long long R = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (R >= lv.NR) return;
    const int d = lv.d, D = lv.D, L = lv.L;
    const RowRec rr = lv.rows[R];
    const double* x = rr.x;
    const double t = rr.t;
    const bool top = (lv.out_uz != nullptr);
    const bool direct = rr.pad != 0 || !(sqrt(lv.T - t) > 0.0);       // rare: degenerate sample points (regenerated normals)
    const RowEntries E = row_entries(lv, R);

    double zs = 0.0, uacc = 0.0;
    // coordinates are processed in passes of 32*JCH so the per-lane accumulators stay in registers
    for (int jpass = 0; jpass < d; jpass += 32 * JCH) {
    double zacc[JCH], xr[JCH];                     // weighted row sums; the row itself
#pragma unroll
    for (int i = 0; i < JCH; ++i) { const int j = jpass + lane + 32 * i; zacc[i] = 0.0; xr[i] = (j < d) ? __ldg(x + j) : 0.0; }
    double Sc = 0.0, Sd = 0.0, ul = 0.0;           // per-lane parts of: sum of weights, sum of weight * drift, u

    // The row's sample points form one entry list; a chunk of 32 entries is resolved lane-parallel, so the weights of ALL segments
    // load in one round trip and the point rows stream without waiting for them (at level 1 a row owns ~10 points in 4 segments).
    for (long long base = 0; base < E.total; base += 32) {
        const long long e = base + lane;
        const bool valid = e < E.total;
        int kind, MC; long long sidx;
        const long long pt = entry_point(lv, R, E, valid ? e : 0, kind, sidx, MC);
        double cw = 0.0, uc = 0.0, tcol = t;
        if (valid) { cw = lv.ev0[pt]; uc = lv.ev1[pt]; tcol = __ldg(lv.P + pt * D + d); }
        const int n = (int)((E.total - base < 32) ? (E.total - base) : 32);
        // stream the rows of the chunk: weight cw (0 for inactive lanes) and point index pt live in the lanes
        int i = 0;
        for (; i + UNR <= n; i += UNR) {
            double ci[UNR]; const double* xp[UNR]; double v[UNR][JCH];
#pragma unroll
            for (int u = 0; u < UNR; ++u) xp[u] = lv.P + __shfl_sync(0xffffffffu, pt, i + u) * D + jpass + lane;
#pragma unroll
            for (int u = 0; u < UNR; ++u)
#pragma unroll
                for (int q = 0; q < JCH; ++q) { const int j = jpass + lane + 32 * q; v[u][q] = (j < d) ? __ldg(xp[u] + 32 * q) : 0.0; }
#pragma unroll
            for (int u = 0; u < UNR; ++u) ci[u] = __shfl_sync(0xffffffffu, cw, i + u);     // after the row loads: the weights arrive meanwhile
#pragma unroll
            for (int u = 0; u < UNR; ++u)
#pragma unroll
                for (int q = 0; q < JCH; ++q) zacc[q] = fma(ci[u], v[u][q], zacc[q]);
        }
        for (; i < n; ++i) {
            const double* xp = lv.P + __shfl_sync(0xffffffffu, pt, i) * D + jpass + lane;
            double v[JCH];
#pragma unroll
            for (int q = 0; q < JCH; ++q) { const int j = jpass + lane + 32 * q; v[q] = (j < d) ? __ldg(xp + 32 * q) : 0.0; }
            const double ci = __shfl_sync(0xffffffffu, cw, i);
#pragma unroll
            for (int q = 0; q < JCH; ++q) zacc[q] = fma(ci, v[q], zacc[q]);
        }
        ul += uc; Sc += cw; Sd = fma(cw, lv.mu * (tcol - t), Sd);
    }
    if (direct) { reduce_direct_terms<JCH>(lv, R, jpass, wbuf_all[warp], zdir_all[warp]); __syncwarp(); }

    // clip / cast / row sum (solvers/ScaSML.py:281-284)
    for (int o = 16; o >= 1; o >>= 1) {
        ul += __shfl_xor_sync(0xffffffffu, ul, o); Sc += __shfl_xor_sync(0xffffffffu, Sc, o); Sd += __shfl_xor_sync(0xffffffffu, Sd, o);
    }
    uacc = ul;
#pragma unroll
    for (int i = 0; i < JCH; ++i) {
        const int j = jpass + lane + 32 * i;
        if (j < d) {
            double z = zacc[i] - xr[i] * Sc - Sd;
            if (direct) z += zdir_all[warp][32 * i + lane];
            if (!lv.partial) {
                z = clip_keep_nan(z, lv.clip);
                if (lv.cast_levels && !top) z = round_f16(z);
            }
            zs += z;
            if (top) lv.out_uz[R * (d + 1) + 1 + j] = z;
        }
    }
    __syncwarp();
    }   // jpass
    if (!lv.partial) {
        uacc = clip_keep_nan(uacc, lv.clip);
        if (lv.cast_levels && !top) uacc = round_f16(uacc);
    }
    for (int o = 16; o >= 1; o >>= 1) zs += __shfl_xor_sync(0xffffffffu, zs, o);
    if (lane == 0) {
        lv.us[L][2 * R] = uacc;
        lv.us[L][2 * R + 1] = zs;
        if (top) lv.out_uz[R * (d + 1)] = uacc;
    }
}

template <int JCH>
int launch_reduce(const LevelDev& lv, cudaStream_t stream) {
    // a sharded rank may own no sample unit of a level (world > units): no points, and below the top level no rows either
    if (lv.npoints > 0) {
        point_weights_kernel<<<(unsigned)cdiv(lv.npoints, 256), 256, 0, stream>>>(lv);
        SC_LAUNCH_CHECK();
    }
    if (lv.NR == 0) return OK;
    const long long threads = lv.NR * 32;
    // UNR = 2 point rows in flight per lane, register budget cut for 8 resident CTAs per SM (measured at C3 against
    // <4, 5> and <2, 6>: 2.6 ms vs 2.9 / 2.8 ms, profiles/r1_sampler_reduce.md)
    reduce_rows_kernel<JCH, 2, 8><<<(unsigned)cdiv(threads, 128), 128, 0, stream>>>(lv);
    SC_LAUNCH_CHECK();
    return OK;
}

// records (REC) exist only where the resident-operand tcgen05 kernel does: D + 1 <= 128, i.e. JP <= 4
template <int JP>
int launch_sample_terminal_jp(const LevelDev& lv, int nslot, int cpts, unsigned grid, int nwarp, size_t smem, cudaStream_t stream) {
    if (JP <= 4 && lv.rec != nullptr) {
        constexpr int J = JP <= 4 ? JP : 4;
        SC_CUDA(cudaFuncSetAttribute(sample_terminal_kernel<J, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sample_terminal_kernel<J, true><<<grid, nwarp * 32, smem, stream>>>(lv, nslot, cpts);
    } else {
        SC_REQUIRE(lv.rec == nullptr, "sampler: operand records need d + 2 <= 128");
        SC_CUDA(cudaFuncSetAttribute(sample_terminal_kernel<JP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sample_terminal_kernel<JP, false><<<grid, nwarp * 32, smem, stream>>>(lv, nslot, cpts);
    }
    SC_LAUNCH_CHECK();
    return OK;
}
template <int JP>
int launch_sample_paths_jp(const LevelDev& lv, int l, int nslot, int cpts, unsigned grid, int nwarp, size_t smem, cudaStream_t stream) {
    if (JP <= 4 && lv.rec != nullptr) {
        constexpr int J = JP <= 4 ? JP : 4;
        SC_CUDA(cudaFuncSetAttribute(sample_paths_kernel<J, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sample_paths_kernel<J, true><<<grid, nwarp * 32, smem, stream>>>(lv, l, nslot, cpts);
    } else {
        SC_REQUIRE(lv.rec == nullptr, "sampler: operand records need d + 2 <= 128");
        SC_CUDA(cudaFuncSetAttribute(sample_paths_kernel<JP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sample_paths_kernel<JP, false><<<grid, nwarp * 32, smem, stream>>>(lv, l, nslot, cpts);
    }
    SC_LAUNCH_CHECK();
    return OK;
}
// JP = passes of 32 columns over the D = d + 1 columns of a point
int launch_sample_terminal(const LevelDev& lv, int nslot, int cpts, unsigned grid, int nwarp, size_t smem, cudaStream_t stream) {
    const int D = lv.D;
    if (D <= 32) return launch_sample_terminal_jp<1>(lv, nslot, cpts, grid, nwarp, smem, stream);
    if (D <= 64) return launch_sample_terminal_jp<2>(lv, nslot, cpts, grid, nwarp, smem, stream);
    if (D <= 128) return launch_sample_terminal_jp<4>(lv, nslot, cpts, grid, nwarp, smem, stream);
    if (D <= 256) return launch_sample_terminal_jp<8>(lv, nslot, cpts, grid, nwarp, smem, stream);
    if (D <= 1024) return launch_sample_terminal_jp<32>(lv, nslot, cpts, grid, nwarp, smem, stream);
    return launch_sample_terminal_jp<64>(lv, nslot, cpts, grid, nwarp, smem, stream);
}
int launch_sample_paths(const LevelDev& lv, int l, int nslot, int cpts, unsigned grid, int nwarp, size_t smem, cudaStream_t stream) {
    const int D = lv.D;
    if (D <= 32) return launch_sample_paths_jp<1>(lv, l, nslot, cpts, grid, nwarp, smem, stream);
    if (D <= 64) return launch_sample_paths_jp<2>(lv, l, nslot, cpts, grid, nwarp, smem, stream);
    if (D <= 128) return launch_sample_paths_jp<4>(lv, l, nslot, cpts, grid, nwarp, smem, stream);
    if (D <= 256) return launch_sample_paths_jp<8>(lv, l, nslot, cpts, grid, nwarp, smem, stream);
    if (D <= 1024) return launch_sample_paths_jp<32>(lv, l, nslot, cpts, grid, nwarp, smem, stream);
    return launch_sample_paths_jp<64>(lv, l, nslot, cpts, grid, nwarp, smem, stream);
}

}  // namespace

// ------------------------------------------------------------------ run -----------------------------------

int PicardPlan::run(const GpView* gp, int route, const double* x_t, double* out_uz, void* workspace, size_t ws_bytes,
                    const __half* normal_table, cudaStream_t stream, PicardStats* stats_out) {
    SC_REQUIRE(ws_bytes >= ws_bytes_, "picard: workspace too small");
    SC_REQUIRE(normal_table != nullptr, "picard: normal table not set (scasml_set_normal_table)");
    SC_REQUIRE(!p_.scasml || gp != nullptr, "picard: ScaSML variant needs a fitted GP");
    const int d = p_.d, D = d + 1, n = p_.n;
    long long launches = 0, eval_points = 0;
    if (B_ == 0) { if (stats_out) *stats_out = stats_; return OK; }
    if (n == 0) {
        SC_CUDA(cudaMemsetAsync(out_uz, 0, (size_t)B_ * D * sizeof(double), stream));
        if (stats_out) *stats_out = stats_;
        return OK;
    }
    char* ws = (char*)workspace;
    std::vector<LevelDev> lvs(n + 1);
    std::vector<std::vector<CallDev>> host_calls(n + 1);
    std::vector<std::vector<long long>> rowbases(MAX_LEVEL + 1);    // must outlive the async copies (synchronised below)
    for (int L = 1; L <= n; ++L) {
        const LevelRec& lr = levels_[L];
        LevelDev& lv = lvs[L];
        std::memset(&lv, 0, sizeof(lv));
        lv.L = L; lv.ncalls = (int)lr.calls.size();
        lv.calls = (const CallDev*)(ws + lr.off_calls);
        lv.rowbase = (const long long*)(ws + lr.off_rowbase);
        lv.NR = lr.NR;
        const bool strided = (L == n) && p_.world > 1;
        lv.world = strided ? p_.world : 1; lv.rank = strided ? p_.rank : 0;
        lv.MCg = mcg_of(L); lv.NT = lr.NT; lv.term_off = lr.term_off;
        for (int l = 0; l < L; ++l) {
            lv.q[l] = q_of(L, l); lv.MCf[l] = mcf_of(L, l); lv.NP[l] = lr.NP[l];
            for (int k = 0; k < lv.q[l]; ++k) {
                const int lk = l * MAX_Q + k;
                lv.set_off[lk] = lr.set_off[lk];
                if (p_.variant == 0) {
                    lv.cnode[lk] = p_.c[k * p_.qmax + (lv.q[l] - 1)];
                    lv.wnode[lk] = p_.w[k * p_.qmax + (lv.q[l] - 1)];
                }
            }
        }
        lv.P = (double*)(ws + lr.off_P);
        lv.gid = (long long*)(ws + lr.off_gid);
        // tcgen05 route, d + 2 <= 128: the samplers also write each point's operand record (f16 hi / lo images of a log2(e) x, |x|^2, sum x), so
        // the evaluation kernel stages a point tile with one bulk copy instead of converting FP64 rows at its tile boundary
        const bool recs = p_.scasml && route == 1 && gp != nullptr && tc_rec_nstep(D) > 0;
        lv.rec = recs ? (uint8_t*)(ws + lr.off_rec) : nullptr;
        lv.rec_nstep = recs ? tc_rec_nstep(D) : 0;
        lv.rec_ascale = recs ? tc_rec_ascale(gp->a) : 0.0f;
        lv.recf = recs ? (double*)(ws + lr.off_recf) : nullptr;
        if (recs) { const TcRecIdx fi = tc_rec_idx(*gp); for (int k = 0; k < TC_REC_NFEAT; ++k) lv.rec_col[k] = fi.col[k]; }
        lv.npoints = lr.npoints;
        lv.rows = (RowRec*)(ws + lr.off_rows);
        lv.ev0 = (double*)(ws + lr.off_ev0);
        lv.ev1 = (double*)(ws + lr.off_ev1);
        for (int q = 1; q <= n; ++q) lv.us[q] = (double*)(ws + levels_[q].off_us);
        lv.out_uz = (L == n) ? out_uz : nullptr;
        lv.d = d; lv.D = D; lv.variant = p_.variant; lv.scasml = p_.scasml;
        lv.stale_delta = p_.stale_delta; lv.cast_levels = p_.cast_levels;
        lv.partial = (L == n && p_.world > 1) ? 1 : 0;
        lv.T = p_.T; lv.mu = p_.mu; lv.sigma = p_.sigma; lv.clip = p_.clip;
        lv.seed = p_.seed; lv.gid0 = p_.gid0; lv.ntab = normal_table;
    }
    // per-call device records (sources resolve to slices of the parents' point buffers)
    for (int L = 1; L <= n; ++L) {
        const LevelRec& lr = levels_[L];
        std::vector<CallDev>& hc = host_calls[L];
        hc.resize(lr.calls.size());
        for (size_t ci = 0; ci < lr.calls.size(); ++ci) {
            const CallRec& r = calls_[lr.calls[ci]];
            CallDev& cd = hc[ci];
            std::memset(&cd, 0, sizeof(cd));
            cd.rowbase = r.rowbase; cd.nrows = r.nrows;
            if (r.parent < 0) {
                cd.xsrc = x_t; cd.gidsrc = nullptr;
            } else {
                const CallRec& pr = calls_[r.parent];
                const LevelDev& plv = lvs[pr.level];
                const long long poff = plv.set_off[r.pl * MAX_Q + r.pk] + pr.rowbase * plv.MCf[r.pl];
                cd.xsrc = plv.P + poff * D;
                cd.gidsrc = plv.gid + poff;
            }
            for (int lk = 0; lk < MAXLK; ++lk) {
                cd.key[lk] = r.key[lk];
                for (int role = 0; role < 2; ++role)
                    cd.childbase[lk][role] = r.child[lk][role] >= 0 ? calls_[r.child[lk][role]].rowbase : -1;
            }
        }
        SC_CUDA(cudaMemcpyAsync(ws + lr.off_calls, hc.data(), hc.size() * sizeof(CallDev), cudaMemcpyHostToDevice, stream));
        std::vector<long long>& rb = rowbases[L];
        rb.resize(hc.size());
        for (size_t i = 0; i < hc.size(); ++i) rb[i] = hc[i].rowbase;
        SC_CUDA(cudaMemcpyAsync(ws + lr.off_rowbase, rb.data(), rb.size() * sizeof(long long), cudaMemcpyHostToDevice, stream));
    }
    // sampler launch shape: persistent CTAs (two per SM when shared memory allows), 64 KB table + per-warp Philox / step buffers
    const int nslot = (d + 14) / 8;                    // Philox blocks covering d consecutive flat indices from any offset
    SC_REQUIRE(D <= 2048, "picard: d > 2047 is not supported by the sampler");
    int dev = 0, nsm = 0;
    SC_CUDA(cudaGetDevice(&dev));
    SC_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    auto sampler_shape = [&](int nq, long long npts, int* nwarp, size_t* smem, unsigned* grid, int* cpts) {
        const size_t per_warp = SMP_WARP_FIXED + (size_t)2 * nslot * 16 + (size_t)nq * 3 * 32 * 8;
        int w = SMP_MAX_WARPS;
        while (w > 1 && SMP_TABLE_BYTES + w * per_warp > (size_t)110 * 1024) w >>= 1;      // two CTAs per SM
        if (w < 8) { w = SMP_MAX_WARPS; while (w > 1 && SMP_TABLE_BYTES + w * per_warp > (size_t)220 * 1024) w >>= 1; }
        *nwarp = w; *smem = SMP_TABLE_BYTES + w * per_warp + 512;     // slack: predicated lanes read past the last warp's buffer
        // points per warp chunk: 32, fewer for small launches so that every resident warp of the machine gets work (a warp walks
        // its chunk two points at a time; small launches are latency-bound by the longest warp)
        int c = 32;
        while (c > 2 && cdiv(npts, c) < 2LL * nsm * w) c >>= 1;
        *cpts = c;
        const long long need = cdiv(cdiv(npts, c), w);
        *grid = (unsigned)std::min<long long>(need, 2LL * nsm);
    };
    // optional CUDA-event timing of the three kernel groups (sampler / evaluation / reduction)
    struct Span { cudaEvent_t a, b; int kind; };
    std::vector<Span> spans;
    const bool timing = p_.timing != 0;
    auto begin_span = [&](int kind) {
        if (!timing) return;
        Span sp; sp.kind = kind;
        cudaEventCreate(&sp.a); cudaEventCreate(&sp.b);
        cudaEventRecord(sp.a, stream);
        spans.push_back(sp);
    };
    auto end_span = [&]() { if (timing) cudaEventRecord(spans.back().b, stream); };
    long long eval_launches = 0, eval_flops = 0;
    // top-down: sample + evaluate
    for (int L = n; L >= 1; --L) {
        const LevelDev& lv = lvs[L];
        const LevelRec& lr = levels_[L];
        begin_span(0);
        if (lv.NR > 0) {                                 // a sharded rank may own nothing below the top level
            row_setup_kernel<<<(unsigned)cdiv(lv.NR, 256), 256, 0, stream>>>(lv);
            SC_LAUNCH_CHECK(); ++launches;
        }
        if (lv.NT > 0) {
            int nwarp, cpts; size_t smem; unsigned grid;
            sampler_shape(0, lv.NT, &nwarp, &smem, &grid, &cpts);
            const int rc = launch_sample_terminal(lv, nslot, cpts, grid, nwarp, smem, stream);
            if (rc != OK) return rc;
            ++launches;
        }
        for (int l = 0; l < L; ++l) {
            if (lv.NP[l] == 0) continue;
            int nwarp, cpts; size_t smem; unsigned grid;
            sampler_shape(p_.variant == 0 ? lv.q[l] : 1, lv.NP[l], &nwarp, &smem, &grid, &cpts);
            const int rc = launch_sample_paths(lv, l, nslot, cpts, grid, nwarp, smem, stream);
            if (rc != OK) return rc;
            ++launches;
        }
        end_span();
        if (p_.scasml) {
            struct Sec { long long off, cnt; int mode; } secs[3] = {
                {lr.term_off, lr.NT, EVAL_TERMINAL}, {lr.ug_off, lr.n_ug, EVAL_UG}, {lr.pde_off, lr.n_pde, EVAL_PDE}};
            const long long fu = 2LL * D * (2LL * gp->Nd + gp->Nb), fp = 2LL * D * ((long long)gp->Nd + gp->Nb);
            for (const Sec& s : secs) {
                if (s.cnt == 0) continue;
                int rc;
                begin_span(1);
                if (route == 1)
                    rc = launch_eval_tc(*gp, nullptr, lv.P + s.off * D, s.cnt, s.mode, lv.ev0 + s.off, lv.ev1 + s.off,
                                        nullptr, nullptr, stream, nullptr,
                                        lv.rec ? lv.rec + (size_t)s.off * tc_rec_bytes(lv.rec_nstep) : nullptr,
                                        lv.rec ? lv.recf + (size_t)s.off * TC_REC_NFEAT : nullptr);
                else
                    rc = launch_eval_f64(*gp, lv.P + s.off * D, s.cnt, s.mode, lv.ev0 + s.off, lv.ev1 + s.off,
                                         nullptr, nullptr, stream);
                if (rc != OK) return rc;
                end_span();
                ++launches; ++eval_launches; eval_points += s.cnt;
                eval_flops += s.cnt * (fu + (s.mode == EVAL_PDE ? fp : 0));
            }
        } else if (lv.NT > 0) {
            mlp_terminal_kernel<<<(unsigned)cdiv(lv.NT * 32, 256), 256, 0, stream>>>(lv);
            SC_LAUNCH_CHECK(); ++launches;
        }
    }
    // bottom-up: reduce
    for (int L = 1; L <= n; ++L) {
        const LevelDev& lv = lvs[L];
        int rc;
        begin_span(2);
        if (d <= 32) rc = launch_reduce<1>(lv, stream);
        else if (d <= 64) rc = launch_reduce<2>(lv, stream);
        else if (d <= 128) rc = launch_reduce<4>(lv, stream);
        else rc = launch_reduce<8>(lv, stream);
        if (rc != OK) return rc;
        end_span();
        launches += (lv.npoints > 0 ? 1 : 0) + (lv.NR > 0 ? 1 : 0);
    }
    stats_.eval_time_ns = stats_.sample_time_ns = stats_.reduce_time_ns = 0;
    if (timing) {
        SC_CUDA(cudaStreamSynchronize(stream));
        for (Span& sp : spans) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, sp.a, sp.b);
            const long long ns = (long long)(ms * 1e6);
            if (sp.kind == 0) stats_.sample_time_ns += ns; else if (sp.kind == 1) stats_.eval_time_ns += ns; else stats_.reduce_time_ns += ns;
            cudaEventDestroy(sp.a); cudaEventDestroy(sp.b);
        }
    }
    stats_.launches = launches;
    stats_.eval_points_total = eval_points;
    stats_.eval_launches = eval_launches;
    stats_.eval_flops = eval_flops;
    if (stats_out) *stats_out = stats_;
    return OK;
}

}  // namespace scasml
