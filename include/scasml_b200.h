/* scasml_b200.h -- C ABI of the B200-native ScaSML correction hot path.
 *
 * The reference (Francis-Fan-create/SCaSML_GP) has no FFI: its boundary is the Python class API
 * (equations/equations.py, models/GP.py, solvers/{MLP,ScaSML}{,_full_history}.py).  Every entry point
 * below is what a binding for one of those methods would call; the file:line it replaces is cited.
 * Conventions: plain pointers and sizes only; `*_dev` pointers are device memory owned by the caller
 * (row-major float64, contiguous); `stream` is a cudaStream_t passed as void*; every function returns
 * 0 on success, non-zero otherwise (scasml_last_error() has the message).  Status 3 (SCASML_ERR_NUMERIC)
 * is a numerical failure of the fit and maps to the reference's ValueError (models/GP.py:264-265);
 * NaN results are data, not errors.  One calling thread per process; the handle owns only the fitted
 * state (centres, alpha, feature tables).
 */
#ifndef SCASML_B200_H
#define SCASML_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SCASML_API __attribute__((visibility("default")))
#else
#define SCASML_API
#endif

#define SCASML_ABI_VERSION 1
#define SCASML_OK 0
#define SCASML_ERR_INVALID 1
#define SCASML_ERR_CUDA 2
#define SCASML_ERR_NUMERIC 3
#define SCASML_ERR_NOMEM 4

#define SCASML_MAX_LEVEL 8
#define SCASML_MAX_Q 8

/* evaluation modes of scasml_gp_eval */
#define SCASML_EVAL_U 0        /* out0 = u_hat(x)                               models/GP.py:653-671 */
#define SCASML_EVAL_TERMINAL 1 /* out0 = g(x) - u_hat(x)                        solvers/ScaSML.py:49-63 */
#define SCASML_EVAL_UG 2       /* out0 = u_hat, out1 = sum_i d_i u_hat          solvers/ScaSML.py:43-46 via GP.py:673-687 */
#define SCASML_EVAL_PDE 3      /* out0 = eps; out1..3 = div_x, lap_x, dt_x u_hat (nullable)   models/GP.py:746-769 */

/* arithmetic routes of the surrogate evaluation */
#define SCASML_ROUTE_F64 0     /* FP64 SIMT contraction (parity anchor) */
#define SCASML_ROUTE_TC 1      /* tcgen05 (kind::f16, split operands, FP32 TMEM accumulators) */

typedef struct scasml_gp scasml_gp; /* opaque: fitted GP_Grad_Dependent_Nonlinear state (models/GP.py:693-769) */

/* Parameter block of one uz_solve call (solvers/ScaSML.py:149, solvers/MLP.py:141,
 * solvers/ScaSML_full_history.py:75, solvers/MLP_full_history.py:64). */
typedef struct scasml_picard_params {
    int variant;      /* 0 quadrature, 1 full history */
    int scasml;       /* 1 ScaSML (defect form, needs a GP), 0 plain MLP */
    int n;            /* level of the top call */
    int d;            /* spatial dimension (n_input - 1) */
    int M;            /* full-history sample base (u_solve(..., M=3)) */
    int qmax;         /* leading dimension of c, w */
    int Qrow[SCASML_MAX_LEVEL];      /* Q [rho-1, :n]   solvers/ScaSML.py:136,221 */
    int Mfrow[SCASML_MAX_LEVEL];     /* Mf[rho-1, :n]   solvers/ScaSML.py:137,223 */
    int Mgrow[SCASML_MAX_LEVEL + 1]; /* Mg[rho-1, :n+1] solvers/ScaSML.py:138-139,187 */
    double c[SCASML_MAX_Q * SCASML_MAX_Q]; /* c[k*qmax + (q-1)]  solvers/ScaSML.py:141-146 */
    double w[SCASML_MAX_Q * SCASML_MAX_Q];
    double T, mu, sigma;             /* equations/equations.py:263-288, geometry() :344 */
    double clip;                     /* equation.uncertainty / norm_estimation, solvers/ScaSML.py:282, MLP.py:272 */
    int stale_delta;                 /* 1: solvers/MLP.py:201,249,270 delta_t behaviour */
    int cast_levels;                 /* 1: inner uz_solve returns rounded to float16 (solvers/ScaSML.py:284) */
    unsigned seed;
    unsigned key_counter;            /* random.split count at entry, solvers/ScaSML.py:27,228 */
    int rank, world;                 /* top-level sample sharding (unit u owned iff u % world == rank) */
    long long gid0;                  /* global index of the first row of x_t (RNG addressing) */
    int timing;                      /* 1: bracket kernel groups with CUDA events (adds a stream sync at the end) */
    int reserved;                    /* must be 0 */
} scasml_picard_params;

typedef struct scasml_picard_stats {
    long long keys_used;         /* random.split calls consumed */
    long long eval_counter;      /* increment of evaluation_counter (solvers/ScaSML.py:26,41,59,205,249,268) */
    long long sample_points;     /* reference-equivalent sample points per test point */
    long long executed_points;   /* points generated + evaluated on this rank (all rows) */
    long long n_calls;           /* uz_solve calls of level >= 1 in the tree */
    long long launches;          /* kernel launches issued */
    long long eval_points_total; /* surrogate evaluations launched on this rank */
    long long eval_launches;     /* surrogate-evaluation kernel launches */
    long long eval_time_ns;      /* device time of the evaluation launches (timing = 1) */
    long long sample_time_ns;    /* device time of the sampler launches (timing = 1) */
    long long reduce_time_ns;    /* device time of the reduction launches (timing = 1) */
    long long eval_flops;        /* algorithmic distance-contraction flops of the evaluation launches */
} scasml_picard_stats;

SCASML_API const char* scasml_last_error(void);
SCASML_API int scasml_abi_version(void);

/* Sampler table: 32768 float16 bit patterns, T[i] = ndtri(0.5 + (i + 0.5)/65536), built on the host.
 * Replaces jax.random.normal(..., dtype=float16) (solvers/ScaSML.py:190,229). Per current device. */
SCASML_API int scasml_set_normal_table(const uint16_t* half_bits_host);

/* equations/equations.py:387-417 (generate_data / generate_test_data -> DeepXDE GeometryXTime.random_points /
 * random_boundary_points on Hypercube x TimeDomain): n float16-valued points [n][d + 1] written on the device.  Philox stream
 * (stream_id, domain 2, seed) replaces NumPy's global generator (DeepXDE draws on the host); boundary != 0 puts one random coordinate of
 * every point on a face of the box. */
SCASML_API int scasml_geometry_points(unsigned seed, unsigned stream_id, long long n, int d, double xmin, double xmax, double t0, double t1,
                           int boundary, double* out_dev, void* stream);

/* equations/equations.py:248-261 (terminal g / exact solution) and :290-304 (generator f) */
SCASML_API int scasml_equation_g(const double* x_t_dev, long long R, int d, double* out_dev, void* stream);
SCASML_API int scasml_equation_f(const double* u_dev, const double* z_dev, long long R, int d, double sigma, double* out_dev,
                      void* stream);

/* models/GP.py:8-26 (__init__): kernel_a = 1 / (sigma_eq * sqrt(d))^2, idx_set = the 5 Hutchinson indices (:35) */
SCASML_API int scasml_gp_create(int d, int n_dom, int n_bdy, const int* idx_set5, double kernel_a, double sigma_eq,
                     double nugget, scasml_gp** out);
SCASML_API int scasml_gp_destroy(scasml_gp* gp);
SCASML_API int scasml_gp_clone(const scasml_gp* gp, scasml_gp** out); /* copy.deepcopy(solver), tests/ComputingBudget.py:138 */
/* models/GP.py:184-192: collocation sets x_t_domain [n_dom][d+1], x_t_boundary [n_bdy][d+1] */
SCASML_API int scasml_gp_set_centres(scasml_gp* gp, const double* x_dom_dev, const double* x_bdy_dev, void* stream);
SCASML_API int scasml_gp_set_alpha(scasml_gp* gp, const double* alpha_dev, void* stream);  /* right_vector, GP.py:599-600 */
SCASML_API int scasml_gp_get_alpha(const scasml_gp* gp, double* alpha_dev, void* stream);
/* models/GP.py:196-258: Gram matrix [phi][phi], phi = 4 n_dom + n_bdy (+ nugget on the diagonal if add_nugget) */
SCASML_API int scasml_gp_gram(const scasml_gp* gp, double* K_dev, int f16_entries, int add_nugget, void* stream);
/* models/GP.py:487-604 GPsolver: damped Newton, then alpha = (K + nugget I)^{-1} z.
 * loss_hist_host has gn_steps + 1 slots (unused slots are NaN); sol_out_dev [3 n_dom] may be NULL. */
SCASML_API size_t scasml_gp_fit_workspace_bytes(const scasml_gp* gp);
SCASML_API int scasml_gp_fit(scasml_gp* gp, const double* g_bdy_dev, const double* sol0_dev, int gn_steps, double damping,
                  double tol, int f16_gram, void* ws_dev, size_t ws_bytes, double* sol_out_dev,
                  double* loss_hist_host, int* steps_done, void* stream);
/* models/GP.py:653-671 predict, :746-769 compute_PDE_loss, and the scalars ScaSML.f needs of :673-687 */
SCASML_API int scasml_gp_eval(const scasml_gp* gp, const double* X_dev, long long R, int mode, int route, double* out0_dev,
                   double* out1_dev, double* out2_dev, double* out3_dev, void* stream);
/* models/GP.py:673-687 compute_gradient: full gradient [R][d+1] */
SCASML_API size_t scasml_gp_gradient_workspace_bytes(const scasml_gp* gp, long long R);
SCASML_API int scasml_gp_gradient(const scasml_gp* gp, const double* X_dev, long long R, double* grad_dev, void* ws_dev,
                       size_t ws_bytes, void* stream);

/* Host-only: enumerate the Picard tree for B rows; workspace size and counters (no GPU needed). */
SCASML_API int scasml_picard_plan(const scasml_picard_params* p, long long B, size_t* ws_bytes, scasml_picard_stats* stats);
/* uz_solve: x_t_dev [B][d+1] -> out_uz_dev [B][1+d] (float64; clipped, not yet rounded to float16).
 * With world > 1 the output holds this rank's un-clipped weighted partial sums: all-reduce (sum) them over
 * the ranks, then call scasml_clip. gp may be NULL when p->scasml == 0. */
SCASML_API int scasml_uz_solve(const scasml_gp* gp, const scasml_picard_params* p, int route, const double* x_t_dev,
                    long long B, double* out_uz_dev, void* ws_dev, size_t ws_bytes, scasml_picard_stats* stats,
                    void* stream);
/* jnp.clip(output_uz, -c, c) of solvers/ScaSML.py:284 (NaN preserved), in place */
SCASML_API int scasml_clip(double* x_dev, long long count, double c, void* stream);

/* The one collective of the path (the reference has no multi-GPU path -- north star: "correction samples are sharded across
 * the GPUs of one box, and a single NCCL allreduce over NVLink combines the per-level sums"): sum the per-rank partial (u, z) blocks that
 * scasml_uz_solve leaves with world > 1, in place, then scasml_clip.  One process per GPU; rank 0 creates the 128-byte id and the host
 * side hands it to the other ranks (file, pipe, MPI ...); scasml_comm_init binds the CURRENT CUDA device.  NCCL is loaded at run time
 * (libnccl.so.2).  The Python classes use torch.distributed for the same all-reduce; these entry points are for hosts without it. */
typedef struct scasml_comm scasml_comm;
SCASML_API int scasml_comm_unique_id(unsigned char* id128_host);
SCASML_API int scasml_comm_init(const unsigned char* id128_host, int rank, int world, scasml_comm** out);
SCASML_API int scasml_allreduce_partial(scasml_comm* comm, double* buf_dev, long long count, void* stream);
SCASML_API int scasml_comm_destroy(scasml_comm* comm);

/* 1 if this handle can use SCASML_ROUTE_TC (d <= 1022: resident-operand kernel up to d = 126, K-streamed kernel above) */
SCASML_API int scasml_gp_tc_supported(const scasml_gp* gp);

#ifdef __cplusplus
}
#endif
#endif /* SCASML_B200_H */
