// Probe (GPU box): 3-D tensor-map TMA load of an fp32 [8][H][W] tile with negative / out-of-range coordinates and zero fill,
// as rto_denoise_tc.cu uses it.  nvcc -gencode arch=compute_100a,code=sm_100a tools/probe/tma3d_probe.cu -o tools/probe/tma3d_probe -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

constexpr int BW = 68, BH = 14, BC = 8;

__global__ void probe(const __grid_constant__ CUtensorMap tm, int x0, int y0, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t bb = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"(BW * BH * BC * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(sb), "l"(&tm), "r"(bb), "r"(x0), "r"(y0), "r"(0) : "memory");
    }
    __syncthreads();
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
        "@p bra D;\n\tbra W;\nD:\n\t}\n" ::"r"(bb) : "memory");
    const float* s = reinterpret_cast<const float*>(smem);
    for (int i = threadIdx.x; i < BW * BH * BC; i += blockDim.x) out[i] = s[i];
}

int main() {
    const int W = 800, H = 800;
    std::vector<float> h((size_t)8 * H * W);
    for (int c = 0; c < 8; ++c)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) h[((size_t)c * H + y) * W + x] = c * 1000000.f + y * 1000.f + x;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&o, BW * BH * BC * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    alignas(64) CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, 8};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {BW, BH, BC};
    const cuuint32_t es[3] = {1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, BW * BH * BC * 4);
    std::vector<float> res(BW * BH * BC);
    const int cases[5][2] = {{100, 200}, {-4, -2}, {776, 790}, {-4, 398}, {56, -2}};
    for (auto& cs : cases) {
        probe<<<1, 256, BW * BH * BC * 4>>>(tm, cs[0], cs[1], o);
        cudaError_t e = cudaDeviceSynchronize();
        printf("case (%d,%d): %s\n", cs[0], cs[1], cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int c = 0; c < BC; ++c)
            for (int y = 0; y < BH; ++y)
                for (int x = 0; x < BW; ++x) {
                    const int gx = cs[0] + x, gy = cs[1] + y;
                    const float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? c * 1000000.f + gy * 1000.f + gx : 0.f;
                    if (res[(c * BH + y) * BW + x] != want) ++bad;
                }
        printf("  mismatches: %d\n", bad);
    }
    return 0;
}
