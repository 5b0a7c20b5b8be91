// rto_render.cu — SPP dispatch of the render kernel (kernel: rto_render_kernel.cuh, one translation unit per SPP value:
// rto_render_spp.cu).  The SPP list is the reference's instantiation list (renderer/src/cuda/volrend.cu:266-278); anything
// else is an error there too.
#include "rto_internal.h"

namespace rto {

#define RTO_DECL_SPP(n) cudaError_t launch_render_spp##n(const RenderArgs& a, int trace, cudaStream_t stream);
RTO_DECL_SPP(1) RTO_DECL_SPP(2) RTO_DECL_SPP(3) RTO_DECL_SPP(4) RTO_DECL_SPP(6) RTO_DECL_SPP(8) RTO_DECL_SPP(16) RTO_DECL_SPP(32)
#undef RTO_DECL_SPP

cudaError_t launch_render(const RenderArgs& a, int spp, int trace, cudaStream_t stream, bool* bad_spp) {
    *bad_spp = false;
    switch (spp) {
        case 1: return launch_render_spp1(a, trace, stream);
        case 2: return launch_render_spp2(a, trace, stream);
        case 3: return launch_render_spp3(a, trace, stream);
        case 4: return launch_render_spp4(a, trace, stream);
        case 6: return launch_render_spp6(a, trace, stream);
        case 8: return launch_render_spp8(a, trace, stream);
        case 16: return launch_render_spp16(a, trace, stream);
        case 32: return launch_render_spp32(a, trace, stream);
        default: *bad_spp = true; return cudaSuccess;
    }
}

}  // namespace rto
