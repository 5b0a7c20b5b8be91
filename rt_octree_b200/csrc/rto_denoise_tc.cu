// rto_denoise_tc.cu — fused GuidanceNet (tcgen05 implicit GEMM) + kernel filter.  Placeholder until the
// tensor-core kernel lands: reports "not available" so rto_denoise uses the CUDA-core path.
#include <cuda_runtime.h>

#include "rto_internal.h"

namespace rto {
size_t denoise_tc_packed_bytes() { return 0; }
cudaError_t denoise_tc_pack_weights(const NetDev&, void*, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_denoise_tc(const NetDev&, const void*, const DenoiseArgs&, cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace rto
