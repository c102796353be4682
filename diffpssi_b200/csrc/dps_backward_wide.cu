// dps_backward_wide.cu -- the adjoint for systems with 17..128 generators: the G = 0 (run-time generator count, block-wide
// exchange) instantiation of dps_adjoint_kernel (dps_backward.cuh), in its own translation unit so that the build parallelises.
#define DPS_BWD_WIDE_TU 1
#include "dps_backward.cuh"

namespace dps {
cudaError_t launch_bwd_wide(const BwdWideParams& p, int cmask, int ngp, cudaStream_t stream) {
    return launch_bwd_wide_impl(p, cmask, ngp, stream);
}
}  // namespace dps
