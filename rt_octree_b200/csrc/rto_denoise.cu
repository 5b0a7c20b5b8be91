// rto_denoise.cu — GuidanceNet forward + multi-level kernel filter, plain CUDA-core version ("simt" path).
//
// Replaces (a) the TorchScript forward the reference runs through libtorch/cuDNN
// (renderer/src/denoiser/denoiser.cpp:42-48; graph = denoiser/network.py:152-168,195-201:
//  half(aux) -> [conv3x3 'same' + bias -> relu6] x2 in fp16 -> float -> softmax(ch 0..L-1) || ch L..2L-1)
// and (b) denoiser::filtering -> applying<surf,16,32,S> for S=1..L (denoiser/extension/filtering.cu:108-228,
// 701-717).  This file is the straightforward, generic (any mid width <= 64, L <= 6) implementation: it is the
// bring-up path and the on-device cross-check for the tensor-core kernel in rto_denoise_tc.cu.
//
// fp16 rounding points follow ATen's cuDNN path (cudnn_convolution then output.add_(bias) on half tensors):
//   y = half(sum_fp32) ; y = half(float(y) + float(b)) ; relu6.
// The fp32 accumulation order here is (ci, ky, kx) ascending — the same as oracle/rt_oracle.c, so this path is
// bit-identical to the oracle; cuDNN's internal order is unpinned.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cfloat>

#include "rto_internal.h"

namespace rto {

__device__ __forceinline__ float round_h(float x) { return __half2float(__float2half_rn(x)); }

constexpr int kNT = 16;  // net output tile (kNT x kNT pixels per block, one thread per pixel)

// smem layout (floats unless noted):
//   in_t  [8][kNT+4][kNT+4]      fp16-rounded aux, zero outside the image
//   mid_t [mid][kNT+2][kNT+2]    relu6(conv1), zero outside the image ('same' padding of conv2 sees zeros)
//   w1    [8][3][3][mid] , w2 [mid][3][3][2L]  (output channel fastest => broadcast vector reads)
template <int MAXC>
__global__ void __launch_bounds__(kNT * kNT) guidance_net_kernel(const NetDev net, const DenoiseArgs d) {
    extern __shared__ float smem[];
    const int mid = net.mid_ch, L = net.levels, out_ch = 2 * L;
    constexpr int IT = kNT + 4, MT = kNT + 2;
    float* in_t = smem;
    float* mid_t = in_t + 8 * IT * IT;
    float* w1 = mid_t + mid * MT * MT;
    float* w2 = w1 + 8 * 9 * mid;
    const int tid = threadIdx.y * kNT + threadIdx.x;
    const int W = d.W, H = d.H;
    const int bx = blockIdx.x * kNT, by = d.y0 + blockIdx.y * kNT;
    const size_t HW = (size_t)W * H;

    for (int i = tid; i < 8 * 9 * mid; i += kNT * kNT) {  // w1 src [co][ci][ky][kx] -> [ci][k][co]
        const int co = i % mid, r = i / mid, k = r % 9, ci = r / 9;
        w1[i] = __half2float(net.w1[(co * 8 + ci) * 9 + k]);
    }
    for (int i = tid; i < mid * 9 * out_ch; i += kNT * kNT) {
        const int co = i % out_ch, r = i / out_ch, k = r % 9, ci = r / 9;
        w2[i] = __half2float(net.w2[(co * mid + ci) * 9 + k]);
    }
    for (int i = tid; i < 8 * IT * IT; i += kNT * kNT) {
        const int x = i % IT, r = i / IT, y = r % IT, c = r / IT;
        const int gx = bx + x - 2, gy = by + y - 2;
        float v = 0.f;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) v = round_h(d.aux[c * HW + (size_t)gy * W + gx]);
        in_t[i] = v;
    }
    __syncthreads();

    // conv1 + bias + relu6 on the (kNT+2)^2 region
    for (int pos = tid; pos < MT * MT; pos += kNT * kNT) {
        const int x = pos % MT, y = pos / MT;
        const int gx = bx + x - 1, gy = by + y - 1;
        const bool inside = gx >= 0 && gx < W && gy >= 0 && gy < H;
        float acc[MAXC];
#pragma unroll
        for (int c = 0; c < MAXC; ++c) acc[c] = 0.f;
        for (int ci = 0; ci < 8; ++ci)
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const float v = in_t[(ci * IT + y + k / 3) * IT + x + k % 3];
                const float* wr = w1 + (ci * 9 + k) * mid;
#pragma unroll
                for (int c = 0; c < MAXC; ++c)
                    if (c < mid) acc[c] = __fmaf_rn(v, wr[c], acc[c]);
            }
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
            if (c < mid) {
                const float bb = __half2float(net.b1[c]);
                float v = net.fused_bias ? round_h(acc[c] + bb) : round_h(round_h(acc[c]) + bb);
                v = fminf(fmaxf(v, 0.f), 6.f);
                mid_t[(c * MT + y) * MT + x] = inside ? v : 0.f;
            }
    }
    __syncthreads();

    // conv2 + bias + relu6 -> float -> softmax / guidance
    const int x = threadIdx.x, y = threadIdx.y;
    const int gx = bx + x, gy = by + y;
    float acc[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) acc[c] = 0.f;
    for (int ci = 0; ci < mid; ++ci)
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const float v = mid_t[(ci * MT + y + k / 3) * MT + x + k % 3];
            const float* wr = w2 + (ci * 9 + k) * out_ch;
#pragma unroll
            for (int c = 0; c < 12; ++c)
                if (c < out_ch) acc[c] = __fmaf_rn(v, wr[c], acc[c]);
        }
    if (gx >= W || gy >= H || gy >= d.y1) return;
    float o[12];
#pragma unroll
    for (int c = 0; c < 12; ++c)
        if (c < out_ch) {
            const float bb = __half2float(net.b2[c]);
            const float v = net.fused_bias ? round_h(acc[c] + bb) : round_h(round_h(acc[c]) + bb);
            o[c] = fminf(fmaxf(v, 0.f), 6.f);
        }
    float mx = -FLT_MAX, sum = 0.f, e[6];
#pragma unroll
    for (int l = 0; l < 6; ++l)
        if (l < L) mx = fmaxf(mx, o[l]);
#pragma unroll
    for (int l = 0; l < 6; ++l)
        if (l < L) { e[l] = expf(o[l] - mx); sum += e[l]; }
    const size_t p = (size_t)gy * W + gx;
#pragma unroll
    for (int l = 0; l < 6; ++l)
        if (l < L) {
            d.weight_map[l * HW + p] = e[l] / sum;
            d.guidance_map[l * HW + p] = o[L + l];
        }
}

// All L levels of the reference's `applying` in one launch (the reference launches one kernel per level and
// read-modify-writes the output surface: filtering.cu:223-227).  Same per-level arithmetic: window max, weights
// exp(g - max), normalise, scale by the level weight; out-of-image taps weigh 0; alpha = 1.
constexpr int kFW = 32, kFH = 16;
__global__ void __launch_bounds__(kFW * kFH) filter_kernel(const float* __restrict__ rgb /* aux ch0..2 planes, or [H][W][4] */,
                                                           size_t chan_stride, int pix_stride, const float* __restrict__ weight,
                                                           const float* __restrict__ guidance, int L, int W, int H,
                                                           int y0, int y1, float4* __restrict__ out,
                                                           float4* __restrict__ save_rgb /* [L][H][W] or null */,
                                                           float* __restrict__ save_max, float* __restrict__ save_inv) {
    extern __shared__ float smem[];
    const int R = L;  // max support
    const int TW = kFW + 2 * R, TH = kFH + 2 * R;
    float* rt = smem;                 // [3][TH][TW]
    float* gt = rt + 3 * TH * TW;     // [L][TH][TW]
    const int tid = threadIdx.y * kFW + threadIdx.x;
    const int bx = blockIdx.x * kFW, by = y0 + blockIdx.y * kFH;
    const size_t HW = (size_t)W * H;
    for (int i = tid; i < TH * TW; i += kFW * kFH) {
        const int x = i % TW, y = i / TW;
        const int gx = bx + x - R, gy = by + y - R;
        const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        const size_t p = (size_t)gy * W + gx;
#pragma unroll
        for (int c = 0; c < 3; ++c) rt[c * TH * TW + i] = in ? rgb[c * chan_stride + p * pix_stride] : 0.f;
        for (int l = 0; l < L; ++l) gt[l * TH * TW + i] = in ? guidance[l * HW + p] : -FLT_MAX;
    }
    __syncthreads();
    const int gx = bx + threadIdx.x, gy = by + threadIdx.y;
    if (gx >= W || gy >= H || gy >= y1) return;
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
    for (int l = 0; l < L; ++l) {
        const int S = l + 1;
        const float* g = gt + l * TH * TW;
        float mx = -FLT_MAX;
        for (int dy = -S; dy <= S; ++dy)
            for (int dx = -S; dx <= S; ++dx)
                mx = fmaxf(mx, g[(threadIdx.y + R + dy) * TW + threadIdx.x + R + dx]);
        float r = 0.f, gg = 0.f, b = 0.f, ksum = 0.f;
        for (int dy = -S; dy <= S; ++dy)
            for (int dx = -S; dx <= S; ++dx) {
                const int q = (threadIdx.y + R + dy) * TW + threadIdx.x + R + dx;
                const float k = __expf(g[q] - mx);
                ksum += k;
                r = __fmaf_rn(rt[q], k, r);
                gg = __fmaf_rn(rt[TH * TW + q], k, gg);
                b = __fmaf_rn(rt[2 * TH * TW + q], k, b);
            }
        const float inv = 1.0f / ksum;
        if (save_rgb) {   // saved for backward (filtering.cu:207-217)
            const size_t sp = l * HW + (size_t)gy * W + gx;
            save_rgb[sp] = make_float4(r * inv, gg * inv, b * inv, 0.f);
            save_max[sp] = mx;
            save_inv[sp] = inv;
        }
        const float w = weight[l * HW + (size_t)gy * W + gx] * inv;
        // level 0 overwrites, levels >= 1 accumulate (filtering.cu:218-227): mul, then add
        o0 = __fadd_rn(o0, __fmul_rn(r, w)); o1 = __fadd_rn(o1, __fmul_rn(gg, w)); o2 = __fadd_rn(o2, __fmul_rn(b, w));
    }
    out[(size_t)gy * W + gx] = make_float4(o0, o1, o2, 1.0f);
}

// Backward of the kernel filter (training side; denoiser/extension/filtering.cu:230-301, 472-576).
//   grad_weight_l(p)   = <dout(p), F_l(p)>                                     (grad_weight_accumulate)
//   grad_guidance_l(q) = sum_{p : q in N_l(p)} w_l(p) k_l(p,q) <dout(p), rgb(q) - F_l(p)>,  k = exp(g_l(q) - max_l(p)) / Z_l(p)
// The reference scatters the second sum with one thread per (p, tap) and atomicAdd; the window is symmetric, so here
// each thread GATHERS its own q over the same taps: no atomics, deterministic summation order.
__global__ void __launch_bounds__(256) filter_backward_kernel(const float4* __restrict__ dout, const float4* __restrict__ img_in,
                                                              const float* __restrict__ weight, const float* __restrict__ guidance,
                                                              const float4* __restrict__ rgb_f, const float* __restrict__ max_map,
                                                              const float* __restrict__ inv_sum, int L, int W, int H,
                                                              float* __restrict__ grad_weight, float* __restrict__ grad_guidance) {
    const size_t HW = (size_t)W * H;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW * L) return;
    const int l = (int)(i / HW);
    const size_t q = i - (size_t)l * HW;
    const int qx = (int)(q % W), qy = (int)(q / W);
    const int S = l + 1;
    {
        const float4 d = dout[q], f = rgb_f[i];
        grad_weight[i] = d.x * f.x + d.y * f.y + d.z * f.z;
    }
    const float g = guidance[i];
    const float4 c = img_in[q];
    float acc = 0.f;
    for (int dy = -S; dy <= S; ++dy) {
        const int py = qy + dy;
        if (py < 0 || py >= H) continue;
        for (int dx = -S; dx <= S; ++dx) {
            const int px = qx + dx;
            if (px < 0 || px >= W) continue;
            const size_t p = (size_t)l * HW + (size_t)py * W + px;
            const float k = __expf(g - max_map[p]) * inv_sum[p];
            const float4 d = dout[(size_t)py * W + px], f = rgb_f[p];
            float res = d.x * (c.x - f.x);
            res += d.y * (c.y - f.y);
            res += d.z * (c.z - f.z);
            acc += res * (weight[p] * k);
        }
    }
    grad_guidance[i] = acc;
}

cudaError_t launch_filter_backward(const float* dout, const float* img_in, const float* weight, const float* guidance,
                                   const float* rgb_f, const float* max_map, const float* inv_sum, int L, int W, int H,
                                   float* grad_weight, float* grad_guidance, cudaStream_t stream) {
    const size_t n = (size_t)W * H * L;
    filter_backward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(dout), reinterpret_cast<const float4*>(img_in), weight, guidance,
        reinterpret_cast<const float4*>(rgb_f), max_map, inv_sum, L, W, H, grad_weight, grad_guidance);
    return cudaGetLastError();
}

// float4 image -> RGBA8 (rgba8_of, rto_internal.h): the stand-alone conversion, used when the image was not produced by a
// kernel that writes the RGBA8 copy itself (render with denoise off and the separable filter both do).
__global__ void __launch_bounds__(256) rgba8_kernel(const float4* __restrict__ img, uchar4* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = img[i];
    out[i] = rgba8_of(v.x, v.y, v.z, v.w);
}
cudaError_t launch_rgba8(const float4* img, uchar4* out, size_t n, cudaStream_t stream) {
    rgba8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(img, out, n);
    return cudaGetLastError();
}

cudaError_t launch_guidance_net_simt(const NetDev& net, const DenoiseArgs& d, cudaStream_t stream) {
    const int rows = d.y1 - d.y0;
    if (rows <= 0) return cudaSuccess;
    dim3 grid((d.W + kNT - 1) / kNT, (rows + kNT - 1) / kNT), block(kNT, kNT);
    const size_t smem = sizeof(float) * (8 * (kNT + 4) * (kNT + 4) + (size_t)net.mid_ch * (kNT + 2) * (kNT + 2) +
                                         8 * 9 * net.mid_ch + (size_t)net.mid_ch * 9 * 2 * net.levels);
    cudaError_t e;
    if (net.mid_ch <= 32) {
        e = cudaFuncSetAttribute(guidance_net_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        guidance_net_kernel<32><<<grid, block, smem, stream>>>(net, d);
    } else {
        e = cudaFuncSetAttribute(guidance_net_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        guidance_net_kernel<64><<<grid, block, smem, stream>>>(net, d);
    }
    return cudaGetLastError();
}

cudaError_t launch_filter_simt(const float* rgb, size_t chan_stride, int pix_stride, const float* weight,
                               const float* guidance, int L, int W, int H, int y0, int y1, float4* out,
                               cudaStream_t stream, float4* save_rgb, float* save_max, float* save_inv) {
    const int rows = y1 - y0;
    if (rows <= 0) return cudaSuccess;
    dim3 grid((W + kFW - 1) / kFW, (rows + kFH - 1) / kFH), block(kFW, kFH);
    const size_t smem = sizeof(float) * (size_t)(3 + L) * (kFW + 2 * L) * (kFH + 2 * L);
    cudaError_t e = cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    filter_kernel<<<grid, block, smem, stream>>>(rgb, chan_stride, pix_stride, weight, guidance, L, W, H, y0, y1, out, save_rgb,
                                                 save_max, save_inv);
    return cudaGetLastError();
}

}  // namespace rto
