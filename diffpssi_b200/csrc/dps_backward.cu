// dps_backward.cu -- discrete adjoint of the stepping kernel (placeholder until the adjoint kernel lands).
#include <cuda_runtime.h>
#include "dps_internal.h"

using namespace dps;

extern "C" {

size_t dps_workspace_bytes(const dps_plan* plan, int64_t P) {
    (void)plan; (void)P;
    return 0;
}

int dps_backward(const dps_plan* plan, int64_t P, const double* upar, const double* vpar, const double* zsub,
                 const double* ckpt, const double* grad_rec, const double* grad_x_final, const uint8_t* want,
                 double* g_x0, double* gslot, void* workspace, size_t workspace_bytes, void* stream) {
    (void)plan; (void)P; (void)upar; (void)vpar; (void)zsub; (void)ckpt; (void)grad_rec; (void)grad_x_final;
    (void)want; (void)g_x0; (void)gslot; (void)workspace; (void)workspace_bytes; (void)stream;
    set_error("dps_backward: not implemented in this build");
    return 4;
}

}  // extern "C"
