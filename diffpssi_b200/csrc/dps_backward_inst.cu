// dps_backward_inst.cu -- compiled once per supported generator count (-DDPS_G=<n>, see Makefile) so that the adjoint
// kernel specialisations of dps_backward.cuh build in parallel.
#include "dps_backward.cuh"

#ifndef DPS_G
#error "compile with -DDPS_G=<generator count>"
#endif

namespace dps {
template cudaError_t launch_bwd_g<DPS_G>(const BwdParams&, int, int, cudaStream_t);
}
