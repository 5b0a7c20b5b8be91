// TEST INFRASTRUCTURE — not part of the product.
//
// Host compile of the REFERENCE's own device headers, included from /root/reference where they lie:
//   renderer/include/volrend/cuda/rt_core.cuh          trace_ray<float,SPP>, _dda_world, _dda_unit,
//                                                      _get_delta_scale, sample_dst<SPP>   (:19-332)
//   renderer/include/volrend/internal/n3tree_query.hpp query_single_from_root             (:13-48)
//   renderer/include/volrend/internal/lumisphere.hpp   maybe_precalc_basis                (:8-87)
//   renderer/3rdparty/pcg32.h                          pcg32
// through macro shims for the few CUDA-only intrinsics.  The per-pixel wrapper below (ray generation, rng
// advance, background composite, aux layout) follows renderer/src/cuda/volrend.cu:24-34,136-202, which
// cannot be included under g++ (kernel launch syntax).  Used (a) to pin oracle/rt_oracle.c on CPU and
// (b) as the `cpu_baseline.kind = "reference"` arm of bench.py.  Tolerance-level, not bit-exact, against
// the GPU: host FMA contraction and logf/expf differ from nvcc's contraction and MUFU lg2/ex2.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

// --- shims for device-only intrinsics used by rt_core.cuh -------------------------------------------------
static inline float min(float a, float b) { return std::fmin(a, b); }
static inline float max(float a, float b) { return std::fmax(a, b); }
// __logf(x) on the GPU is lg2.approx(x) * ln2 ; keep the same factorisation so that the C oracle's CPU
// fallback thresholds (log2f(x) * -0.6931472f) are bit-identical to this shim's.
static inline float ref_shim_logf(float x) { return log2f(x) * 0.6931472f; }
static inline float ref_shim_expf(float x) { return exp2f(x * 1.442695f); }
#define __logf ref_shim_logf
#define __expf ref_shim_expf

#include "volrend/cuda/rt_core.cuh"

using namespace volrend;

extern "C" {

// Renders pixels [pix_begin, pix_end) of a W x H frame with the reference's trace_ray.
//   aux    : [8][H][W] fp32 (volrend.cu:187-202), written only for the pixel range
//   thresh : optional [W*H][SPP] sorted thresholds to inject instead of sample_dst (NULL = reference RNG)
// returns 0, or -1 for an unsupported SPP (volrend.cu:266-278).
int ref_cpu_render(const int32_t* child, const void* data_f16, int capacity, int data_dim, int basis_dim,
                   const float* offset, const float* scale,
                   const float* c2w12, int W, int H, float fx, float fy,
                   int spp, float step_size, float sigma_thresh, float background,
                   uint64_t rng_state, uint64_t rng_inc,
                   int pix_begin, int pix_end, float* aux, int nthreads);
}

namespace {

template <int SPP>
void render_range(const internal::TreeSpec& tree, const RenderOptions& opt, const float* m, int W, int H,
                  float fx, float fy, pcg32 rng0, int pix_begin, int pix_end, float* aux, int nthreads) {
    const int SIZE = W * H;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads)
    for (int idx = pix_begin; idx < pix_end; ++idx) {
        const int ix = idx % W, iy = idx / W;
        float dir[3], cen[3], out[4] = {0.f, 0.f, 0.f, 0.f};
        // screen2worlddir, volrend.cu:24-34
        float xyz[3] = {(ix - 0.5f * W) / fx, -(iy - 0.5f * H) / fy, -1.0f};
        _mv3(m, xyz, dir);
        _normalize(dir);
        _copy3(m + 9, cen);
        float vdir[3] = {dir[0], dir[1], dir[2]};
        for (int i = 0; i < 3; ++i) cen[i] = tree.offset[i] + tree.scale[i] * cen[i];  // :142-144
        pcg32 rng = rng0;
        rng.advance(idx * SPP);  // :157
        device::trace_ray<float, SPP>(tree, dir, vdir, cen, opt, 1e9f, out, rng);
        const float nalpha = 1.f - out[3];  // :174-179 (offscreen)
        const float remain = opt.background_brightness * nalpha;
        out[0] += remain; out[1] += remain; out[2] += remain;
        aux[idx] = out[0];                 aux[idx + SIZE] = out[1];
        aux[idx + 2 * SIZE] = out[2];      aux[idx + 3 * SIZE] = out[3];
        aux[idx + 4 * SIZE] = out[0] * out[0]; aux[idx + 5 * SIZE] = out[1] * out[1];
        aux[idx + 6 * SIZE] = out[2] * out[2]; aux[idx + 7 * SIZE] = out[3] * out[3];
    }
}

}  // namespace

extern "C" int ref_cpu_render(const int32_t* child, const void* data_f16, int capacity, int data_dim,
                              int basis_dim, const float* offset, const float* scale, const float* c2w12,
                              int W, int H, float fx, float fy, int spp, float step_size,
                              float sigma_thresh, float background, uint64_t rng_state, uint64_t rng_inc,
                              int pix_begin, int pix_end, float* aux, int nthreads) {
    N3Tree t;
    t.N = 2;
    t.data_dim = data_dim;
    t.data_format.format = basis_dim > 0 ? DataFormat::SH : DataFormat::RGBA;
    t.data_format.basis_dim = basis_dim > 0 ? basis_dim : -1;
    t.capacity = capacity;
    for (int i = 0; i < 3; ++i) { t.scale[i] = scale[i]; t.offset[i] = offset[i]; }
    t.data_.ptr = data_f16;
    t.child_.ptr = child;
    t.extra_.ptr = nullptr;
    internal::TreeSpec spec(t, /*cpu=*/true);
    RenderOptions opt;
    opt.step_size = step_size;
    opt.sigma_thresh = sigma_thresh;
    opt.background_brightness = background;
    opt.spp = spp;
    pcg32 rng;
    rng.state = rng_state;
    rng.inc = rng_inc;
    if (nthreads <= 0) nthreads = 1;
#define CASE(S) case S: render_range<S>(spec, opt, c2w12, W, H, fx, fy, rng, pix_begin, pix_end, aux, nthreads); break;
    switch (spp) {
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(6) CASE(8) CASE(16) CASE(32)
        default: return -1;
    }
#undef CASE
    return 0;
}
