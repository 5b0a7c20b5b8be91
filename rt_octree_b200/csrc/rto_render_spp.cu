// rto_render_spp.cu — one translation unit per SPP value of the render kernel (compiled with -DRTO_SPP=n, see the
// Makefile): every TU instantiates render_kernel<SPP, TRACE, GRID> for its SPP and exports its launcher, so the 24 kernel
// variants x 8 SPP values (the reference's instantiation list, renderer/src/cuda/volrend.cu:266-278) build in parallel.
#include "rto_render_kernel.cuh"

#ifndef RTO_SPP
#error "compile with -DRTO_SPP=<1|2|3|4|6|8|16|32>"
#endif

#define RTO_CAT_(a, b) a##b
#define RTO_CAT(a, b) RTO_CAT_(a, b)

namespace rto {

cudaError_t RTO_CAT(launch_render_spp, RTO_SPP)(const RenderArgs& a, int trace, cudaStream_t stream) {
    return launch_spp<RTO_SPP>(a, trace, stream);
}

}  // namespace rto

#if defined(RTO_TILE_LOG) && RTO_SPP == 6
// diagnostic build (tools/tile_log.py): the log pointer of the SPP 6 kernels
extern "C" int rto_debug_set_tile_log(void* dev_ptr) {
    unsigned long long* p = static_cast<unsigned long long*>(dev_ptr);
    return cudaMemcpyToSymbol(rto::rto_tile_log, &p, sizeof p) == cudaSuccess ? 0 : -3;
}
#endif
