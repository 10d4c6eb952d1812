// extern "C" entry points of libscasml_b200.so (declared in include/scasml_b200.h).
#include <cstring>
#include <mutex>
#include <new>
#include <vector>
#include "abi_handle.cuh"
#include "gp_fit.cuh"
#include "picard.cuh"

static_assert(sizeof(scasml_picard_params) == sizeof(scasml::PicardParams), "ABI struct mismatch");
static_assert(sizeof(scasml_picard_stats) == sizeof(scasml::PicardStats), "ABI struct mismatch");
static_assert(SCASML_MAX_LEVEL == scasml::MAX_LEVEL && SCASML_MAX_Q == scasml::MAX_Q, "ABI constants");
static_assert(sizeof(scasml::LevelDev) <= 4000, "LevelDev must fit the kernel parameter space");

namespace scasml {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error_cstr() { return g_err.c_str(); }

static __half* g_ntab[64] = {nullptr};

const __half* normal_table_for_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    return g_ntab[dev];
}

namespace {

// Collocation / test points on the device: DeepXDE's GeometryXTime(Hypercube, TimeDomain) samplers (the reference's
// equations/equations.py:387-417 calls random_points / random_boundary_points) with Philox uniforms instead of NumPy's global generator.
// Point i draws the flat indices i (d + 2) + j of stream (stream_id, domain 2, seed): j < d the coordinates, j = d the face dimension of a
// boundary point (that coordinate is rounded to a face: np.round, half to even), j = d + 1 the time.  float16-valued (DeepXDE float16).
// One warp per point, lane <-> coordinate.  NumPy statement: oracle/equation.py::EquationOracle.generate_data_philox.
__global__ void geometry_points_kernel(PhiloxKey key, long long n, int d, double xmin, double xmax, double t0, double t1, int boundary,
                                       double* __restrict__ out) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const unsigned long long base = (unsigned long long)i * (unsigned long long)(d + 2);
    int face = -1;
    if (boundary) {
        face = (int)(chunk_to_uniform(chunk16(base + (unsigned long long)d, key)) * (double)d);
        if (face > d - 1) face = d - 1;
    }
    double* row = out + i * (long long)(d + 1);
    for (int j = lane; j < d; j += 32) {
        double u = chunk_to_uniform(chunk16(base + (unsigned long long)j, key));
        if (j == face) u = rint(u);
        row[j] = round_f16(__dadd_rn(__dmul_rn(xmax - xmin, u), xmin));      // NumPy's two roundings, no FMA contraction
    }
    if (lane == 0) row[d] = round_f16(__dadd_rn(__dmul_rn(chunk_to_uniform(chunk16(base + (unsigned long long)d + 1ull, key)), t1 - t0), t0));
}

__global__ void equation_g_kernel(const double* __restrict__ x, long long R, int d, double* __restrict__ out) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const double* p = x + r * (d + 1);
    double acc = 0.0;
    for (int j = lane; j <= d; j += 32) acc += p[j];
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[r] = 1.0 - 1.0 / (1.0 + exp(acc));
}

__global__ void equation_f_kernel(const double* __restrict__ u, const double* __restrict__ z, long long R, int d,
                                  double sigma, double* __restrict__ out) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const double* p = z + r * d;
    double acc = 0.0;
    for (int j = lane; j < d; j += 32) acc += p[j];
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[r] = sigma * u[r] * acc;
}

__global__ void clip_kernel(double* __restrict__ x, long long n, double c) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    x[i] = (v < -c) ? -c : ((v > c) ? c : v);
}

// flag = 1 if some centre coordinate is not a float16 value (the tcgen05 route's stage-1 B operand would not be exact)
__global__ void centres_f16_check_kernel(const double* __restrict__ C, long n, int* __restrict__ flag) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = C[i];
    if ((double)__half2float(__double2half(v)) != v) *flag = 1;      // NaN centres fail too
}

}  // namespace
}  // namespace scasml

using namespace scasml;

static int gp_alloc(scasml_gp* g) {
    const long nc = g->ncpad();
    SC_CUDA(cudaMalloc(&g->C, (size_t)nc * g->v.D * sizeof(double)));
    SC_CUDA(cudaMalloc(&g->feat, (size_t)nc * CF_STRIDE * sizeof(double)));
    SC_CUDA(cudaMalloc(&g->alpha, (size_t)g->phi() * sizeof(double)));
    SC_CUDA(cudaMemset(g->C, 0, (size_t)nc * g->v.D * sizeof(double)));
    SC_CUDA(cudaMemset(g->feat, 0, (size_t)nc * CF_STRIDE * sizeof(double)));
    SC_CUDA(cudaMemset(g->alpha, 0, (size_t)g->phi() * sizeof(double)));
    g->v.C = g->C; g->v.feat = g->feat;
    g->tc = TcState();
    g->v.tc = nullptr;
    if (tc_supported(g->v)) {
        const size_t bytes = tc_image_bytes(g->v, &g->tc);
        SC_CUDA(cudaMalloc(&g->tc.images, bytes));
        SC_CUDA(cudaMemset(g->tc.images, 0, bytes));
        g->v.tc = &g->tc;
    }
    return OK;
}

namespace scasml {
int ensure_scratch_pool() {
    static bool done[64] = {};
    int dev = 0;
    SC_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || done[dev]) return OK;
    cudaMemPool_t pool;
    SC_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    unsigned long long keep = ~0ull;
    SC_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    done[dev] = true;
    return OK;
}
}  // namespace scasml

extern "C" {

const char* scasml_last_error(void) { return last_error_cstr(); }
int scasml_abi_version(void) { return SCASML_ABI_VERSION; }

int scasml_set_normal_table(const uint16_t* half_bits_host) {
    SC_REQUIRE(half_bits_host != nullptr, "normal table is null");
    int dev = 0;
    SC_CUDA(cudaGetDevice(&dev));
    SC_REQUIRE(dev >= 0 && dev < 64, "device index");
    if (!g_ntab[dev]) SC_CUDA(cudaMalloc(&g_ntab[dev], 32768 * sizeof(__half)));
    SC_CUDA(cudaMemcpy(g_ntab[dev], half_bits_host, 32768 * sizeof(__half), cudaMemcpyHostToDevice));
    return OK;
}

int scasml_geometry_points(unsigned seed, unsigned stream_id, long long n, int d, double xmin, double xmax, double t0, double t1,
                           int boundary, double* out_dev, void* stream) {
    SC_REQUIRE(d >= 1 && out_dev != nullptr, "geometry_points: invalid argument");
    if (n <= 0) return OK;
    geometry_points_kernel<<<(unsigned)cdiv(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(make_key(stream_id, 2u, seed), n, d, xmin, xmax, t0, t1,
                                                                                        boundary, out_dev);
    SC_LAUNCH_CHECK();
    return OK;
}

int scasml_equation_g(const double* x_t_dev, long long R, int d, double* out_dev, void* stream) {
    if (R <= 0) return OK;
    equation_g_kernel<<<(unsigned)cdiv(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(x_t_dev, R, d, out_dev);
    SC_LAUNCH_CHECK();
    return OK;
}

int scasml_equation_f(const double* u_dev, const double* z_dev, long long R, int d, double sigma, double* out_dev,
                      void* stream) {
    if (R <= 0) return OK;
    equation_f_kernel<<<(unsigned)cdiv(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(u_dev, z_dev, R, d, sigma, out_dev);
    SC_LAUNCH_CHECK();
    return OK;
}

int scasml_gp_create(int d, int n_dom, int n_bdy, const int* idx_set5, double kernel_a, double sigma_eq,
                     double nugget, scasml_gp** out) {
    SC_REQUIRE(out != nullptr, "gp_create: out is null");
    SC_REQUIRE(d >= 2 && n_dom >= 1 && n_bdy >= 0, "gp_create: sizes");
    SC_REQUIRE(idx_set5 != nullptr, "gp_create: idx_set is null");
    for (int m = 0; m < MC_IDX; ++m) SC_REQUIRE(idx_set5[m] >= 0 && idx_set5[m] < d, "gp_create: idx_set out of range");
    scasml_gp* g = new (std::nothrow) scasml_gp();
    SC_REQUIRE(g != nullptr, "gp_create: out of host memory");
    g->v.d = d; g->v.D = d + 1; g->v.Nd = n_dom; g->v.Nb = n_bdy;
    g->v.NdPad = (int)(cdiv(n_dom, CENTRE_PAD) * CENTRE_PAD);
    g->v.NbPad = (int)(cdiv(n_bdy, CENTRE_PAD) * CENTRE_PAD);
    g->v.a = kernel_a; g->v.sig2 = sigma_eq * sigma_eq;
    for (int m = 0; m < MC_IDX; ++m) g->v.I[m] = idx_set5[m];
    g->nugget = nugget; g->sigma_eq = sigma_eq;
    const int rc = gp_alloc(g);
    if (rc != OK) { scasml_gp_destroy(g); return rc; }
    *out = g;
    return OK;
}

int scasml_gp_destroy(scasml_gp* g) {
    if (!g) return OK;
    cudaFree(g->C); cudaFree(g->feat); cudaFree(g->alpha); cudaFree(g->tc.images);
    delete g;
    return OK;
}

int scasml_gp_clone(const scasml_gp* src, scasml_gp** out) {
    SC_REQUIRE(src && out, "gp_clone: null");
    scasml_gp* g = new (std::nothrow) scasml_gp();
    SC_REQUIRE(g != nullptr, "gp_clone: out of host memory");
    g->v = src->v; g->nugget = src->nugget; g->sigma_eq = src->sigma_eq;
    g->has_centres = src->has_centres; g->has_alpha = src->has_alpha; g->centres_f16 = src->centres_f16;
    g->C = nullptr; g->feat = nullptr; g->alpha = nullptr;
    int rc = gp_alloc(g);
    if (rc != OK) { scasml_gp_destroy(g); return rc; }
    const long nc = g->ncpad();
    SC_CUDA(cudaMemcpy(g->C, src->C, (size_t)nc * g->v.D * sizeof(double), cudaMemcpyDeviceToDevice));
    SC_CUDA(cudaMemcpy(g->feat, src->feat, (size_t)nc * CF_STRIDE * sizeof(double), cudaMemcpyDeviceToDevice));
    SC_CUDA(cudaMemcpy(g->alpha, src->alpha, (size_t)g->phi() * sizeof(double), cudaMemcpyDeviceToDevice));
    if (g->tc.images && src->tc.images)
        SC_CUDA(cudaMemcpy(g->tc.images, src->tc.images, g->tc.total_bytes, cudaMemcpyDeviceToDevice));
    *out = g;
    return OK;
}

int scasml_gp_set_centres(scasml_gp* g, const double* x_dom_dev, const double* x_bdy_dev, void* stream) {
    SC_REQUIRE(g && x_dom_dev && (x_bdy_dev || g->v.Nb == 0), "gp_set_centres: null");
    cudaStream_t st = (cudaStream_t)stream;
    const int D = g->v.D;
    SC_CUDA(cudaMemsetAsync(g->C, 0, (size_t)g->ncpad() * D * sizeof(double), st));
    SC_CUDA(cudaMemcpyAsync(g->C, x_dom_dev, (size_t)g->v.Nd * D * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (g->v.Nb > 0)
        SC_CUDA(cudaMemcpyAsync(g->C + (size_t)g->v.NdPad * D, x_bdy_dev, (size_t)g->v.Nb * D * sizeof(double),
                                cudaMemcpyDeviceToDevice, st));
    g->has_centres = true; g->has_alpha = false;
    // The tcgen05 route takes the centres as an EXACT f16 operand (DeepXDE float16 collocation points, the reference's
    // experiment_run.py:46): verify it instead of assuming it; other centres keep the FP64 route (scasml_gp_tc_supported = 0).
    g->centres_f16 = true;
    if (g->tc.images) {
        int* flag = nullptr;
        int h = 0;
        SC_CUDA(cudaMallocAsync((void**)&flag, sizeof(int), st));
        SC_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
        const long n = (long)g->ncpad() * D;
        centres_f16_check_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(g->C, n, flag);
        SC_LAUNCH_CHECK();
        SC_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        SC_CUDA(cudaStreamSynchronize(st));
        SC_CUDA(cudaFreeAsync(flag, st));
        g->centres_f16 = (h == 0);
    }
    return build_centre_features(g->v, nullptr, g->feat, st);
}

int scasml_gp_set_alpha(scasml_gp* g, const double* alpha_dev, void* stream) {
    SC_REQUIRE(g && alpha_dev, "gp_set_alpha: null");
    SC_REQUIRE(g->has_centres, "gp_set_alpha: centres not set");
    cudaStream_t st = (cudaStream_t)stream;
    if (alpha_dev != g->alpha)
        SC_CUDA(cudaMemcpyAsync(g->alpha, alpha_dev, (size_t)g->phi() * sizeof(double), cudaMemcpyDeviceToDevice, st));
    g->has_alpha = true;
    int rc = build_centre_features(g->v, g->alpha, g->feat, st);
    if (rc != OK) return rc;
    if (g->tc.images && g->centres_f16) rc = tc_build_images(g->v, g->tc, st);
    return rc;
}

int scasml_gp_get_alpha(const scasml_gp* g, double* alpha_dev, void* stream) {
    SC_REQUIRE(g && alpha_dev, "gp_get_alpha: null");
    SC_CUDA(cudaMemcpyAsync(alpha_dev, g->alpha, (size_t)g->phi() * sizeof(double), cudaMemcpyDeviceToDevice,
                            (cudaStream_t)stream));
    return OK;
}

int scasml_gp_gram(const scasml_gp* g, double* K_dev, int f16_entries, int add_nugget, void* stream) {
    SC_REQUIRE(g && K_dev, "gp_gram: null");
    SC_REQUIRE(g->has_centres, "gp_gram: centres not set");
    return gram_assemble(g->v, K_dev, add_nugget ? g->nugget : 0.0, f16_entries, (cudaStream_t)stream);
}

size_t scasml_gp_fit_workspace_bytes(const scasml_gp* g) { return g ? fit_workspace_bytes(g->v.Nd, g->v.Nb) : 0; }

int scasml_gp_fit(scasml_gp* g, const double* g_bdy_dev, const double* sol0_dev, int gn_steps, double damping,
                  double tol, int f16_gram, void* ws_dev, size_t ws_bytes, double* sol_out_dev,
                  double* loss_hist_host, int* steps_done, void* stream) {
    SC_REQUIRE(g && sol0_dev && ws_dev && loss_hist_host && steps_done, "gp_fit: null");
    SC_REQUIRE(g_bdy_dev || g->v.Nb == 0, "gp_fit: g_bdy is null");
    SC_REQUIRE(g->has_centres, "gp_fit: centres not set");
    SC_REQUIRE(gn_steps >= 0, "gp_fit: gn_steps");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = gp_fit_device(g->v, g_bdy_dev, sol0_dev, gn_steps, damping, tol, g->nugget, f16_gram, ws_dev, ws_bytes,
                           g->alpha, sol_out_dev, loss_hist_host, steps_done, st);
    if (rc != OK) return rc;
    return scasml_gp_set_alpha(g, g->alpha, stream);
}

int scasml_gp_eval(const scasml_gp* g, const double* X_dev, long long R, int mode, int route, double* out0_dev,
                   double* out1_dev, double* out2_dev, double* out3_dev, void* stream) {
    SC_REQUIRE(g != nullptr, "gp_eval: null handle");
    SC_REQUIRE(g->has_alpha, "gp_eval: GP is not fitted (call GPsolver first)");
    if (route == SCASML_ROUTE_TC) SC_REQUIRE(g->centres_f16, "tcgen05 route needs float16-valued collocation points (use SCASML_ROUTE_F64)");
    if (route == SCASML_ROUTE_TC)
        return launch_eval_tc(g->v, nullptr, X_dev, (long)R, mode, out0_dev, out1_dev, out2_dev, out3_dev, (cudaStream_t)stream);
    SC_REQUIRE(route == SCASML_ROUTE_F64, "gp_eval: unknown route");
    return launch_eval_f64(g->v, X_dev, (long)R, mode, out0_dev, out1_dev, out2_dev, out3_dev, (cudaStream_t)stream, true);
}

size_t scasml_gp_gradient_workspace_bytes(const scasml_gp* g, long long R) {
    return g ? gradient_scratch_bytes(g->v, (long)R) : 0;
}

int scasml_gp_gradient(const scasml_gp* g, const double* X_dev, long long R, double* grad_dev, void* ws_dev,
                       size_t ws_bytes, void* stream) {
    SC_REQUIRE(g && X_dev && grad_dev && ws_dev, "gp_gradient: null");
    SC_REQUIRE(g->has_alpha, "gp_gradient: GP is not fitted");
    return launch_gradient_f64(g->v, X_dev, (long)R, grad_dev, (double*)ws_dev, ws_bytes, (cudaStream_t)stream);
}

int scasml_picard_plan(const scasml_picard_params* p, long long B, size_t* ws_bytes, scasml_picard_stats* stats) {
    SC_REQUIRE(p != nullptr, "picard_plan: null params");
    PicardParams pp;
    std::memcpy(&pp, p, sizeof(pp));
    PicardPlan plan;
    const int rc = plan.build(pp, (long)B);
    if (rc != OK) return rc;
    if (ws_bytes) *ws_bytes = plan.workspace_bytes();
    if (stats) std::memcpy(stats, &plan.stats(), sizeof(PicardStats));
    return OK;
}

int scasml_uz_solve(const scasml_gp* g, const scasml_picard_params* p, int route, const double* x_t_dev, long long B,
                    double* out_uz_dev, void* ws_dev, size_t ws_bytes, scasml_picard_stats* stats, void* stream) {
    SC_REQUIRE(p && x_t_dev && out_uz_dev && ws_dev, "uz_solve: null");
    PicardParams pp;
    std::memcpy(&pp, p, sizeof(pp));
    SC_REQUIRE(!pp.scasml || (g && g->has_alpha), "uz_solve: ScaSML needs a fitted GP");
    SC_REQUIRE(!pp.scasml || g->v.d == pp.d, "uz_solve: dimension mismatch between GP and params");
    SC_REQUIRE(!(pp.scasml && route == SCASML_ROUTE_TC) || (g->tc.images && g->centres_f16),
               "uz_solve: tcgen05 route unavailable for this GP (d > 1022 or collocation points not float16-valued): use SCASML_ROUTE_F64");
    PicardPlan plan;
    int rc = plan.build(pp, (long)B);
    if (rc != OK) return rc;
    PicardStats st;
    rc = plan.run(g ? &g->v : nullptr, route, x_t_dev, out_uz_dev, ws_dev, ws_bytes, normal_table_for_current_device(),
                  (cudaStream_t)stream, &st);
    if (rc != OK) return rc;
    if (stats) std::memcpy(stats, &st, sizeof(st));
    return OK;
}

int scasml_clip(double* x_dev, long long count, double c, void* stream) {
    if (count <= 0) return OK;
    clip_kernel<<<(unsigned)cdiv(count, 256), 256, 0, (cudaStream_t)stream>>>(x_dev, count, c);
    SC_LAUNCH_CHECK();
    return OK;
}

int scasml_gp_tc_supported(const scasml_gp* g) { return (g && g->tc.images && g->centres_f16) ? 1 : 0; }

}  // extern "C"
