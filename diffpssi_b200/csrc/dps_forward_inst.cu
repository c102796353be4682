// dps_forward_inst.cu -- compiled once per supported generator count (-DDPS_G=<n>, see Makefile) so that the kernel
// specialisations of dps_forward.cuh build in parallel.
#include "dps_forward.cuh"

#ifndef DPS_G
#error "compile with -DDPS_G=<generator count>"
#endif

namespace dps {
template cudaError_t launch_forward_g<DPS_G>(const FwdParams&, int, int, cudaStream_t, int*, int*);
}
