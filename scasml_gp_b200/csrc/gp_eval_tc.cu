// tcgen05 route of the surrogate evaluation -- placeholder until the TMEM kernel lands.
#include "picard.cuh"
namespace scasml {
int launch_eval_tc(const GpView&, const void*, const double*, long, int, double*, double*, double*, double*, cudaStream_t) {
    set_error("tcgen05 evaluation route is not available in this build");
    return ERR_INVALID;
}
}  // namespace scasml
