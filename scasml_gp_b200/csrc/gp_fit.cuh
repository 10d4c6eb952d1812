// GP fit entry points (gp_fit.cu).
#pragma once
#include "gp.cuh"

namespace scasml {

int gram_assemble(const GpView& gp, double* K /*[phi][phi]*/, double nugget, int f16_entries, cudaStream_t st);
int cholesky_lower(double* A, long n, double* invdiag, int* d_fail, cudaStream_t st);
int tri_inverse_lower(const double* L, long n, const double* invdiag, double* X, double* scratch /* [n x n] */, cudaStream_t st);
int lu_solve_inplace(double* H, long n, double* rhs, int* d_fail, cudaStream_t st);
int dgemm(int M, int N, int Kd, double alpha, const double* A, long sai, long sak, const double* B, long sbk, long sbj,
          double beta, double* C, long ldc, int lower_only, cudaStream_t st);
size_t fit_workspace_bytes(int Nd, int Nb);
int gp_fit_device(const GpView& gp, const double* g_bdy, const double* sol0, int gn_steps, double damping, double tol,
                  double nugget, int f16_entries, void* workspace, size_t ws_bytes, double* alpha_out, double* sol_out,
                  double* loss_hist_host, int* steps_done, cudaStream_t st);

}  // namespace scasml
