// dps_forward.cu -- dispatch of the stepping kernel specialisations (kernel body: dps_forward.cuh).
#include <cuda_runtime.h>
#include "dps_internal.h"

namespace dps {

template <int G>
cudaError_t launch_forward_g(const FwdParams& p, int cmask, int pmode, cudaStream_t stream, int* regs, int* spb);

#define DPS_EXTERN(G) extern template cudaError_t launch_forward_g<G>(const FwdParams&, int, int, cudaStream_t, int*, int*);
DPS_EXTERN(1) DPS_EXTERN(2) DPS_EXTERN(3) DPS_EXTERN(4) DPS_EXTERN(5) DPS_EXTERN(6) DPS_EXTERN(8) DPS_EXTERN(10) DPS_EXTERN(16)
#undef DPS_EXTERN

int fwd_template_g(int n_gen) {
    static const int sizes[] = {1, 2, 3, 4, 5, 6, 8, 10, 16};
    for (int v : sizes)
        if (n_gen <= v) return v;
    return -1;
}

void fwd_geometry(int g_template, int* threads, int* sims_per_block) {
    *threads = 32 * DPS_FWD_WARPS;
    *sims_per_block = DPS_FWD_WARPS * (32 / g_template);
}

cudaError_t launch_forward(const FwdParams& p, int g_template, int cmask, int pmode, cudaStream_t stream, int* regs, int* spb) {
#ifdef DPS_ONLY_G10      // kernel A/B experiment builds (tools/build_variant.sh): only the bench specialisation is linked
    return g_template == 10 ? launch_forward_g<10>(p, cmask, pmode, stream, regs, spb) : cudaErrorInvalidValue;
#else
    switch (g_template) {
        case 1: return launch_forward_g<1>(p, cmask, pmode, stream, regs, spb);
        case 2: return launch_forward_g<2>(p, cmask, pmode, stream, regs, spb);
        case 3: return launch_forward_g<3>(p, cmask, pmode, stream, regs, spb);
        case 4: return launch_forward_g<4>(p, cmask, pmode, stream, regs, spb);
        case 5: return launch_forward_g<5>(p, cmask, pmode, stream, regs, spb);
        case 6: return launch_forward_g<6>(p, cmask, pmode, stream, regs, spb);
        case 8: return launch_forward_g<8>(p, cmask, pmode, stream, regs, spb);
        case 10: return launch_forward_g<10>(p, cmask, pmode, stream, regs, spb);
        case 16: return launch_forward_g<16>(p, cmask, pmode, stream, regs, spb);
        default: return cudaErrorInvalidValue;
    }
#endif
}

}  // namespace dps
