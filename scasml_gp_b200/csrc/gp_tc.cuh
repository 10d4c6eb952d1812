// tcgen05 route: per-fit operand images and feature records (gp_eval_tc.cu).
#pragma once
#include "gp.cuh"

namespace scasml {

// FP32 centre feature record of the tcgen05 epilogue (weights pre-multiplied by K_j = exp(-a |y_j|^2 / 2))
enum TcFeat : int {
    TF_SY = 0, TF_YT = 1, TF_Y0 = 2, TF_SYROLL = 3,
    TF_YI = 4,       // 5 values
    TF_YIR = 9,      // 5 values
    TF_A1 = 16, TF_A3D = 17, TF_A4 = 18, TF_A5 = 19,
};
constexpr int TC_NF = 20;

struct TcState {
    uint8_t* images = nullptr;   // per centre tile: [C image][Croll image][feature records]
    size_t tile_bytes = 0;
    int KB = 0;                  // 64-wide K blocks of the contraction (d + 1 <= 64 KB)
    int ntile_dom = 0, ntile_bdy = 0;
};

int tc_supported(const GpView& gp);
size_t tc_image_bytes(const GpView& gp, TcState* st);              // fills KB / tile counts / tile_bytes
int tc_build_images(const GpView& gp, const TcState& st, cudaStream_t stream);
int tc_selftest(const void* A_dev, const void* B_dev, float* D_dev, int K, int N, unsigned lbo16, unsigned sbo16,
                unsigned layout, unsigned kstep_bytes, cudaStream_t stream);

}  // namespace scasml
