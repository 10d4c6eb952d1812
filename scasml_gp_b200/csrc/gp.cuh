// Device-side view of the fitted PDE-constrained GP surrogate (reference: models/GP.py).
#pragma once
#include "common.cuh"

namespace scasml {

// Per-centre feature record (float64), built once per fit by build_centre_features().
// Centre order: [domain 0..Nd) | zero pad to NdPad | boundary 0..Nb) | zero pad to NbPad].
// Padded centres carry zero weights, so they contribute exactly 0.
enum CentreFeat : int {
    CF_NY = 0,      // |y|^2 over all d+1 coordinates
    CF_SY = 1,      // sum_{i<d} y_i
    CF_YT = 2,      // y_d (time)
    CF_Y0 = 3,      // y_0
    CF_SYROLL = 4,  // sum_{i<d} roll(y)_i = sum_{i=1..d} y_i
    CF_YI = 5,      // y_{I_m},   m = 0..4
    CF_YIR = 10,    // y_{I_m+1}, m = 0..4  (= roll(y)_{I_m})
    CF_A1 = 15,     // alpha on kappa (domain: block 1; boundary: block 2)
    CF_A3 = 16,     // alpha on lap_y kappa
    CF_A4 = 17,     // alpha on dt_y kappa
    CF_A5 = 18,     // alpha on div_y kappa
    CF_STRIDE = 20
};

constexpr int CENTRE_PAD = 64;   // centre counts are padded to a multiple of this

struct GpView {
    int d, D;                 // spatial dim, d+1
    int Nd, Nb;               // collocation counts
    int NdPad, NbPad;         // padded to CENTRE_PAD
    double a;                 // 1 / kernel_width^2   (models/GP.py:25,43)
    double sig2;              // equation sigma^2     (equations/equations.py:288)
    int I[MC_IDX];            // Hutchinson index set (models/GP.py:35)
    const double* C;          // [NdPad + NbPad][D] centres, float64 (float16-valued)
    const double* feat;       // [NdPad + NbPad][CF_STRIDE]
    const void* tc;           // host pointer to the TcState of the tcgen05 route (never dereferenced on the device)
};

// evaluation modes of the fused surrogate kernel
enum EvalMode : int {
    EVAL_U = 0,        // out0 = u_hat(x)                       (models/GP.py:653-671)
    EVAL_TERMINAL = 1, // out0 = g(x) - u_hat(x)                (solvers/ScaSML.py:49-63)
    EVAL_UG = 2,       // out0 = u_hat, out1 = div_x u_hat      (what ScaSML.f needs of compute_gradient)
    EVAL_PDE = 3       // out0 = eps (models/GP.py:746-769); optional out1..3 = div_x, lap_x, dt_x of u_hat
};

// FP64 SIMT route (gp_eval.cu).  X: [R][D] float64 device.  Null outputs are skipped.
// split_small: deal the centre tiles of a small batch out over several CTAs per point tile (public evaluation API: the top-level u_hat of a
// u_solve).  The partial sums are added in a fixed order, but the split depends on R -- the Picard plan keeps it off so that its results do
// not depend on how a batch is cut into workspace chunks.
int launch_eval_f64(const GpView& gp, const double* X, long R, int mode,
                    double* out0, double* out1, double* out2, double* out3, cudaStream_t stream, bool split_small = false);

// full gradient vector for the public compute_gradient API (models/GP.py:673-687)
int launch_gradient_f64(const GpView& gp, const double* X, long R, double* grad /*[R][D]*/,
                        double* scratch, size_t scratch_bytes, cudaStream_t stream);
size_t gradient_scratch_bytes(const GpView& gp, long R);

int build_centre_features(const GpView& gp, const double* alpha /*[4Nd+Nb] or null*/, double* feat_out,
                          cudaStream_t stream);

}  // namespace scasml
