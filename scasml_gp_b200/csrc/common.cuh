// Shared device/host helpers for the ScaSML-on-B200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace scasml {

// ---- status codes of the C-ABI (include/scasml_b200.h) ----
enum : int { OK = 0, ERR_INVALID = 1, ERR_CUDA = 2, ERR_NUMERIC = 3, ERR_NOMEM = 4 };

void set_error(const std::string& msg);          // abi.cu
// Stream-ordered scratch (cudaMallocAsync): keeps freed blocks in the device's default memory pool across synchronisations.  With the default
// release threshold (0) every stream synchronisation hands the pool's memory back to the driver and the next cudaMallocAsync pays for a fresh
// allocation (tens of ms per u_solve through the public API, which synchronises for its device -> host copy).  abi.cu; once per device.
int ensure_scratch_pool();
const char* last_error_cstr();

#define SC_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            ::scasml::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));       \
            return ::scasml::ERR_CUDA;                                                     \
        }                                                                                  \
    } while (0)

#define SC_LAUNCH_CHECK()  SC_CUDA(cudaGetLastError())

#define SC_REQUIRE(cond, msg)                                                              \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            ::scasml::set_error(std::string("invalid argument: ") + (msg));                \
            return ::scasml::ERR_INVALID;                                                  \
        }                                                                                  \
    } while (0)

constexpr int MC_IDX = 5;          // Hutchinson index count (reference models/GP.py:30)
constexpr int MAX_LEVEL = 8;       // max Picard level n
constexpr int MAX_Q = 8;           // max quadrature nodes per level

// ---- Philox4x32-10, value = F(key, flat index)  (mirrors oracle/rng.py) ----
struct PhiloxKey { uint32_t k0, k1; };

__host__ __device__ inline PhiloxKey make_key(uint32_t stream, uint32_t domain, uint32_t seed) {
    PhiloxKey k;
    k.k0 = stream;
    k.k1 = ((domain & 1u) << 31) | (seed & 0x7FFFFFFFu);
    return k;
}

__device__ __forceinline__ uint4 philox4x32_10(uint64_t blk, PhiloxKey key) {
    uint32_t c0 = (uint32_t)blk, c1 = (uint32_t)(blk >> 32), c2 = 0u, c3 = 0u;
    uint32_t k0 = key.k0, k1 = key.k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// 16-bit chunk j (0..7) of a Philox block: word j/2, low half first.
__device__ __forceinline__ uint32_t chunk_of(const uint4& w, int j) {
    const uint32_t word = (j < 4) ? ((j < 2) ? w.x : w.y) : ((j < 6) ? w.z : w.w);
    return (j & 1) ? (word >> 16) : (word & 0xFFFFu);
}

__device__ __forceinline__ uint32_t chunk16(uint64_t flat, PhiloxKey key) {
    const uint4 w = philox4x32_10(flat >> 3, key);
    return chunk_of(w, (int)(flat & 7));
}

// float16-valued normal from the host-built half table (32768 entries, antisymmetric).
__device__ __forceinline__ double chunk_to_normal(const __half* __restrict__ tab, uint32_t c) {
    const bool neg = c < 32768u;
    const uint32_t idx = neg ? (32767u - c) : (c - 32768u);
    const float v = __half2float(__ldg(tab + idx));
    return (double)(neg ? -v : v);
}

__device__ __forceinline__ double chunk_to_uniform(uint32_t c) {
    return (double)(c >> 5) * (1.0 / 2048.0);
}

// Local quadrature time / weight, same operation order as the reference
// (solvers/ScaSML.py:174-175): ((T - t) * c) / T + t.  No FMA contraction, so the +-1 ulp
// steps of the duplicated lgwt nodes come out with the same sign as in NumPy.
__device__ __forceinline__ double cloc_of(double T, double t, double c) {
    return __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(T, t), c), T), t);
}
__device__ __forceinline__ double wloc_of(double T, double t, double w) {
    return __ddiv_rn(__dmul_rn(__dsub_rn(T, t), w), T);
}

__host__ __device__ inline long cdiv(long a, long b) { return (a + b - 1) / b; }

// round-to-nearest-even float16 value of a double (double -> half is a single rounding)
__device__ __forceinline__ double round_f16(double v) { return (double)__half2float(__double2half(v)); }

}  // namespace scasml
