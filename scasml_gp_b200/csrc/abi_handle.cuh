// The opaque handle of the C ABI (include/scasml_b200.h: `scasml_gp`) -- shared by abi.cu and the debug hooks (abi_debug.cu).
#pragma once
#include "../../include/scasml_b200.h"
#include "common.cuh"
#include "gp.cuh"
#include "gp_tc.cuh"

struct scasml_gp {
    scasml::GpView v{};
    double nugget = 1e-2;
    double sigma_eq = 0.25;
    double* C = nullptr;      // [NdPad + NbPad][D]
    double* feat = nullptr;   // [NdPad + NbPad][CF_STRIDE]
    double* alpha = nullptr;  // [4 Nd + Nb]
    scasml::TcState tc{};     // tcgen05 route: operand images (rebuilt with every set_alpha)
    bool has_centres = false, has_alpha = false;
    bool centres_f16 = true;  // every centre coordinate survives double -> half -> double (the tcgen05 route's exact B operand)
    long phi() const { return 4L * v.Nd + v.Nb; }
    long ncpad() const { return (long)v.NdPad + v.NbPad; }
};

namespace scasml {
const __half* normal_table_for_current_device();   // abi.cu
}
