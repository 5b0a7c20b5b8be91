// rto_internal.h — device-side argument blocks shared by the kernels and the C-ABI layer (not installed).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "rto_ray.cuh"

// Cache hints for the per-frame streaming buffers (aux, maps, image): written once and read once per frame, they should
// not displace the tree's brick grid from L2.  st.global.cs / ld.global.cs mark the lines evict-first.
// (B200, bench workload, 4 frames in flight: +0.8 % frames/s; with the L2 persistence window over the bricks +3 %.)
#ifndef RTO_NO_STREAM_HINTS
#define RTO_ST(p, v) __stcs((p), (v))
#define RTO_LD_LAST(p) __ldcs(p)
#else
#define RTO_ST(p, v) (*(p) = (v))
#define RTO_LD_LAST(p) __ldg(p)
#endif

namespace rto {

constexpr int kMaxDevices = 64;   // per-device launch state (function attributes are per device)

// float4 image -> RGBA8 with the reference CLI's conversion `(uint8_t)(v * 255)` (main_headless.cpp:534-537): truncation,
// no clamp, low byte of the integer — so the bytes equal what the reference hands to its PNG writer.
#if defined(__CUDACC__)
__device__ __forceinline__ uchar4 rgba8_of(float r, float g, float b, float a) {
    return make_uchar4((unsigned char)(__float2int_rz(__fmul_rn(r, 255.f)) & 0xff), (unsigned char)(__float2int_rz(__fmul_rn(g, 255.f)) & 0xff),
                       (unsigned char)(__float2int_rz(__fmul_rn(b, 255.f)) & 0xff), (unsigned char)(__float2int_rz(__fmul_rn(a, 255.f)) & 0xff));
}
#endif


// HBM layout of a loaded tree (structure of arrays)
struct TreeDev {
    const uint32_t* nodes;  // [cap*8] node words: internal = ABSOLUTE child node id; leaf = 0x80000000 | sigma fp16 bits
    const __half* sh;       // [cap*8][sh_stride] per-leaf colour payload (SH coefficients [3][basis] or rgb), fp16
    int sh_stride;          // halfs per leaf, multiple of 8 (27 -> 32 = 64 B for SH9)
    int basis_dim;          // >0: SH with that many coefficients per channel ; <=0: RGBA leaves
    int max_depth;          // max child look-ups to reach a leaf (sizes the per-ray ancestor stack)
    GridDev grid;           // sparse brick grid (grid.K == 0: none)
    size_t grid_brick_bytes;
};

struct TraceOut {           // all optional (nullptr); indexed by the FULL-FRAME pixel index
    uint32_t* steps;
    int32_t* term;
    uint32_t* src_bits;
    uint32_t* t_bits;
    uint64_t* leaf_hash;
    uint32_t* depth_sum;
    uint32_t* n_hits;
    uint32_t* n_loads;
    int32_t* hit_leaf;      // [n][spp]
    uint32_t* hit_cnt;      // [n][spp]
    int32_t* leaf_seq;      // [n][max_seq]
    float* thresh;          // [n][spp] sorted thresholds dst[] as used by the kernel
    int max_seq;
};

struct RenderArgs {
    FrameParams fp;
    TreeDev tree;
    uint64_t rng_state, rng_inc;   // frame-level pcg32 state (ctx.rng passed by value: volrend.cu:90,157)
    int x0, y0, x1, y1;            // pixel rectangle to render (full frame, or a band for the tile split)
    float* aux;                    // [8][H][W] fp32
    float4* img;                   // [H][W] float4, may be nullptr
    uchar4* img8;                  // [H][W] RGBA8 copy of img written by the same store (nullptr: not wanted)
    int* tile_counter;             // [2] device ints owned by the context: next tile, finished warps (self re-arming)
    const AdvanceMap* adv_rows;    // [H] pcg32 jump-ahead maps for iy*W*spp
    const AdvanceMap* adv_cols;    // [W] ... for ix*spp
    TraceOut tr;
};

// trace: 0 = off, 1 = tree walker with the traversal record, 2 = production brick-grid marcher with the record
cudaError_t launch_render(const RenderArgs& a, int spp, int trace, cudaStream_t stream, bool* bad_spp);

// GuidanceNet (deployed form) weights on the device, fp16
struct NetDev {
    const __half* w1;  // [mid][in][3][3]
    const __half* b1;  // [mid]
    const __half* w2;  // [2L][mid][3][3]
    const __half* b2;  // [2L]
    int in_ch, mid_ch, levels;
    int fused_bias;    // 0: half(half(acc)+b) (ATen cuDNN path, default) ; 1: half(acc+b) (PyTorch CPU fp16 conv)
};

struct DenoiseArgs {
    const float* aux;      // [8][H][W]
    float4* img_out;       // [H][W]
    float* weight_map;     // [L][H][W] optional debug/inspection output (nullptr = not written)
    float* guidance_map;   // [L][H][W] optional
    int W, H;
    int y0, y1;            // rows to produce (band for the tile split); input rows outside [0,H) are padding
};

}  // namespace rto

namespace rto {
cudaError_t launch_guidance_net_simt(const NetDev& net, const DenoiseArgs& d, cudaStream_t stream);
// rgb[c*chan_stride + p*pix_stride]: (HW,1) for the planar aux buffer, (1,4) for an interleaved [H][W][4] image
cudaError_t launch_filter_simt(const float* rgb, size_t chan_stride, int pix_stride, const float* weight,
                               const float* guidance, int L, int W, int H, int y0, int y1, float4* out,
                               cudaStream_t stream, float4* save_rgb = nullptr, float* save_max = nullptr,
                               float* save_inv = nullptr);
cudaError_t launch_rgba8(const float4* img, uchar4* out, size_t n, cudaStream_t stream);
cudaError_t launch_filter_backward(const float* dout, const float* img_in, const float* weight, const float* guidance,
                                   const float* rgb_f, const float* max_map, const float* inv_sum, int L, int W, int H,
                                   float* grad_weight, float* grad_guidance, cudaStream_t stream);
}  // namespace rto
