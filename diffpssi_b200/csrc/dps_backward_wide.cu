// dps_backward_wide.cu -- the adjoint for systems with 17..128 generators: the G = 0 (run-time generator count, block-wide
// exchange) instantiation of dps_adjoint_kernel (dps_backward.cuh), in its own translation unit so that the build parallelises.
#include "dps_backward.cuh"

namespace dps {
cudaError_t launch_bwd_wide(const BwdWideParams& p, bool ctrl, int ngp, cudaStream_t stream) {
    return launch_bwd_wide_impl(p, ctrl, ngp, stream);
}
}  // namespace dps
