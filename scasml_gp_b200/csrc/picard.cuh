// Flattened multilevel-Picard recursion: host plan + level-wise device batches.
// Reference: solvers/ScaSML.py:149-305, solvers/MLP.py:141-288 (quadrature variants),
//            solvers/ScaSML_full_history.py:75-221, solvers/MLP_full_history.py:64-196.
#pragma once
#include <vector>
#include "common.cuh"
#include "gp.cuh"

namespace scasml {

// plain-C parameter block of the C-ABI (mirrored in include/scasml_b200.h as scasml_picard_params)
struct PicardParams {
    int variant;              // 0 = quadrature (ScaSML.py / MLP.py), 1 = full history
    int scasml;               // 1 = defect form with the surrogate, 0 = plain MLP
    int n;                    // Picard level of the top call
    int d;                    // spatial dimension
    int M;                    // full-history sample base (u_solve(..., M=3))
    int qmax;                 // leading dimension of c / w
    int Qrow[MAX_LEVEL];      // Q [rho-1, 0..n)   (solvers/ScaSML.py:221)
    int Mfrow[MAX_LEVEL];     // Mf[rho-1, 0..n)   (solvers/ScaSML.py:223)
    int Mgrow[MAX_LEVEL + 1]; // Mg[rho-1, 0..n]   (solvers/ScaSML.py:187)
    double c[MAX_Q * MAX_Q];  // quadrature nodes  c[k*qmax + (q-1)]  (row-major copy of the reference's c)
    double w[MAX_Q * MAX_Q];  // quadrature weights
    double T, mu, sigma;      // terminal time, drift, diffusion (equations/equations.py:263-288)
    double clip;              // +-uncertainty (ScaSML) or +-norm_estimation (MLP)
    int stale_delta;          // 1 = MLP.py's stale delta_t (solvers/MLP.py:201,249,270)
    int cast_levels;          // 1 = round inner uz_solve returns to float16 (solvers/ScaSML.py:284)
    unsigned seed;
    unsigned key_counter;     // running random.split count at entry (solvers/ScaSML.py:27,228)
    int rank, world;          // top-level sample sharding: unit u owned iff u % world == rank
    long long gid0;           // global index of the first test point of this batch
    int timing;               // 1: CUDA-event timing of kernel groups (adds a stream sync)
    int reserved;             // must be 0
};

struct PicardStats {
    long long keys_used;          // random.split calls made by this solve
    long long eval_counter;       // increment of the reference's evaluation_counter
    long long sample_points;      // reference-equivalent sample points per test point
    long long executed_points;    // sample points actually generated+evaluated per test point (this rank, B-normalised)
    long long n_calls;            // uz_solve calls of level >= 1 in the tree
    long long launches;           // kernel launches issued by run()
    long long eval_points_total;  // surrogate evaluations launched (all rows, this rank)
    long long eval_launches;
    long long eval_time_ns, sample_time_ns, reduce_time_ns;
    long long eval_flops;         // 2 (d+1) x centres per distance contraction, summed over evaluation launches
};

constexpr int MAXLK = MAX_LEVEL * MAX_Q;

struct CallDev {
    long long rowbase, nrows;
    const double* xsrc;                 // [nrows][D] rows of this call
    const long long* gidsrc;            // [nrows] global row ids (null: gid0 + i)
    unsigned key[MAXLK];                // Philox stream id of step (l,k)
    long long childbase[MAXLK][2];      // row base of child call (level l / level l-1), -1 = none
};

// per-row record of a level (row_setup_kernel): what every sampler / reduction warp needs of its parent row in ONE load
// instead of the dependent chain  call search -> call record -> parent row pointer -> global id -> time column
struct RowRec {
    const double* x;                    // the row's point [D] (a row of the parent's point buffer, or of x_t at the top level)
    long long gid;                      // global row id (RNG addressing independent of batching / sharding)
    double t;                           // time column of the row
    int call, pad;                      // index of the owning call in LevelDev::calls
};

struct LevelDev {
    int L, ncalls;
    const CallDev* calls;
    const long long* rowbase;           // [ncalls] compact copy of calls[i].rowbase: the binary search touches 1-2 cache lines, not one per probe
    long long NR;                       // rows at this level
    long long npoints;                  // sample points of this level (this rank)
    int world, rank;                    // unit striding (top level only; else 1,0)
    int MCg; long long NT;              // terminal samples per row; owned terminal points
    long long term_off;                 // point offset of the terminal section
    int q[MAX_LEVEL], MCf[MAX_LEVEL];   // per l
    long long NP[MAX_LEVEL];            // owned points per step set of l
    long long set_off[MAXLK];           // point offset of set (l,k)
    double cnode[MAXLK], wnode[MAXLK];  // c[k][q-1], w[k][q-1]
    double* P;                          // point buffer [npoints][D]
    long long* gid;                     // [npoints]
    uint8_t* rec;                       // [npoints][tc_rec_bytes(rec_nstep)] operand records of the tcgen05 evaluation kernel (gp_tc.cuh), written by the
                                        // samplers while a point's coordinates are in registers; null: FP64 route, plain MLP, d + 2 > 128
    int rec_nstep; float rec_ascale;    // k-steps of the records; a log2(e)
    double* recf;                       // [npoints][12] feature blocks (x_0, t, x_I, x_{I+1}: gp_tc.cuh::TcRecIdx), written with the records
    int rec_col[12];                    // their columns in the FP64 row
    RowRec* rows;                       // [NR]
    double* ev0; double* ev1;           // evaluation outputs per point
    double* us[MAX_LEVEL + 1];          // finalized (u, zsum) per row, per level
    double* out_uz;                     // top level: [NR][1+d]
    int d, D, variant, scasml, stale_delta, cast_levels, partial;
    double T, mu, sigma, clip, sig_eq;
    unsigned seed;
    long long gid0;
    const __half* ntab;
};

class PicardPlan {
public:
    int build(const PicardParams& p, long B);
    size_t workspace_bytes() const { return ws_bytes_; }
    const PicardStats& stats() const { return stats_; }
    // gp may be null when p.scasml == 0.  x_t: [B][D] device, out_uz: [B][1+d] device.
    int run(const GpView* gp, int route, const double* x_t, double* out_uz, void* workspace, size_t ws_bytes,
            const __half* normal_table, cudaStream_t stream, PicardStats* stats_out);

    struct CallRec {
        int level; long long nrows, rowbase;
        int parent, pl, pk;
        unsigned key[MAXLK];
        int child[MAXLK][2];
    };
    struct LevelRec {
        long long NR = 0; std::vector<int> calls;
        long long NT = 0, npoints = 0, n_ug = 0, n_pde = 0;
        long long NP[MAX_LEVEL] = {0};
        long long set_off[MAXLK] = {0};
        long long term_off = 0, ug_off = 0, pde_off = 0;
        size_t off_P = 0, off_gid = 0, off_rec = 0, off_recf = 0, off_rows = 0, off_ev0 = 0, off_ev1 = 0, off_us = 0, off_calls = 0, off_rowbase = 0, off_lvdev = 0;
    };
    const std::vector<CallRec>& calls() const { return calls_; }
    const std::vector<LevelRec>& levels() const { return levels_; }

private:
    int q_of(int L, int l) const;
    int mcf_of(int L, int l) const;
    int mcg_of(int L) const;
    long long owned(long long units, bool strided) const;
    int add_call(int level, long long nrows, int parent, int pl, int pk);

    PicardParams p_{};
    long B_ = 0;
    std::vector<CallRec> calls_;
    std::vector<LevelRec> levels_;
    PicardStats stats_{};
    unsigned keyctr_ = 0;
    size_t ws_bytes_ = 0;
};

}  // namespace scasml
