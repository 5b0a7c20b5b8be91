// rto_api.cu — the extern "C" layer declared in include/rtoctree_b200.h.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/rtoctree_b200.h"
#include "rto_internal.h"
#include "rto_tree.h"

namespace rto {
cudaError_t launch_guidance_net_tc(const NetDev& net, const void* packed, const DenoiseArgs& d, bool exp_guidance, cudaStream_t stream);
cudaError_t launch_filter_fast(const float* aux, const float* weight, const float* guidance, int W, int H, int y0, int y1,
                               float4* out, uchar4* out8, cudaStream_t stream);
size_t denoise_tc_packed_bytes();
cudaError_t denoise_tc_pack_weights(const NetDev& net, void* packed_dev, cudaStream_t stream);
}  // namespace rto

struct rto_tree {
    uint32_t* nodes = nullptr;
    __half* payload = nullptr;
    uint32_t* grid_top = nullptr;      // sparse brick grid (rto_ray.cuh GridDev), optional
    uint32_t* grid_bricks = nullptr;
    uint8_t* grid_bricks8 = nullptr;   // byte plane of the bricks (depth | dense flag)
    uint32_t* grid_leaf_top = nullptr;     // leaf-id planes: flat leaf index per level-K cell / per brick cell
    uint32_t* grid_leaf_bricks = nullptr;
    uint32_t* grid_top_m = nullptr;        // march table of the fused-index marcher (rto_ray.cuh FusedIdx)
    int grid_K = 0;
    int64_t n_bricks = 0;
    rto_tree_info info{};
};

struct rto_context {
    int W = 0, H = 0;
    float* aux = nullptr;
    float4* img = nullptr;
    uchar4* img8 = nullptr;         // RGBA8 copy of img, allocated on first use; once it exists the kernels that produce img
                                    // (render with denoise off, separable filter) write it in the same epilogue
    uint64_t img_gen = 0, img8_gen = 0;   // host-side generation of img / of the RGBA8 copy (equal: img8 is current)
    // tile split: where the kernels that produce the final image store it instead of img / img8 (rto_context_set_image_target):
    // typically the image of ANOTHER context, possibly in a peer GPU's memory (peer-direct stores over NVLink)
    float4* img_target = nullptr;
    uchar4* img8_target = nullptr;
    float4* out_img() const { return img_target ? img_target : img; }
    uchar4* out_img8() const { return img_target ? img8_target : img8; }
    float* weight_map = nullptr;    // [6][H][W] scratch for the two-kernel denoise path
    float* guidance_map = nullptr;
    int* tile_counter = nullptr;    // [2] work counter of the persistent render kernel
    rto::AdvanceMap* adv = nullptr;  // [H + W] pcg32 jump-ahead tables for (adv_spp, adv_inc)
    rto::AdvanceMap* adv_host = nullptr;   // pinned staging copy: the upload is an async copy on the render stream
    cudaEvent_t adv_copied = nullptr;      // completion of the last table upload (the staging buffer is reused)
    int adv_spp = 0;
    uint64_t adv_inc = 0;
    bool capturing = false;                // inside rto_frame_create's stream capture: no timer events
    rto::Pcg32 rng{};
    // Timer
    bool timing = false;
    cudaEvent_t ev_start[3]{}, ev_stop[3]{};
    bool ev_used[3]{};
    cudaStream_t timer_stream = nullptr;
    float sum_ms[3]{};
    int frames = 0;
};

struct rto_net {
    __half *w1 = nullptr, *b1 = nullptr, *w2 = nullptr, *b2 = nullptr;
    void* packed = nullptr;   // tensor-core operand images (rto_denoise_tc.cu)
    int in_ch = 0, mid_ch = 0, levels = 0;
    int impl = 0;
    int fused_bias = 0;
    rto::NetDev dev() const { return rto::NetDev{w1, b1, w2, b2, in_ch, mid_ch, levels, fused_bias}; }
    bool tc_capable() const { return in_ch == 8 && mid_ch == 32 && levels == 4 && packed != nullptr; }
};

namespace {
thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define RTO_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(RTO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

void pcg32_seed(rto::Pcg32& r, uint64_t initstate, uint64_t initseq = 1) {  // pcg32.h:53-59
    r.state = 0u;
    r.inc = (initseq << 1u) | 1u;
    rto::pcg32_next(r);
    r.state += initstate;
    rto::pcg32_next(r);
}

int check_opt(const rto_render_options* opt) {
    if (!opt) return fail(RTO_ERR_INVALID, "options is NULL");
    if (opt->enable_probe) return fail(RTO_ERR_UNSUPPORTED, "enable_probe is not supported (lumisphere probe is out of scope)");
    switch (opt->spp) {
        case 1: case 2: case 3: case 4: case 6: case 8: case 16: case 32: return RTO_OK;
        default: return fail(RTO_ERR_UNSUPPORTED, "spp == %d not supported.", opt->spp);  // volrend.cu:275-277
    }
}

void timer_start(rto_context* c, int i, cudaStream_t s) {
    if (!c->timing || c->capturing) return;
    c->timer_stream = s;
    cudaEventRecord(c->ev_start[i], s);
}
void timer_stop(rto_context* c, int i, cudaStream_t s) {
    if (!c->timing || c->capturing) return;
    cudaEventRecord(c->ev_stop[i], s);
    c->ev_used[i] = true;
}
}  // namespace

extern "C" {

const char* rto_last_error(void) { return g_err.c_str(); }
int rto_abi_version(void) { return RTO_ABI_VERSION; }
int64_t rto_launch_count(void) { return g_launches.load(); }

int rto_set_device(int device) {
    RTO_CUDA(cudaSetDevice(device));
    return RTO_OK;
}
int rto_device_count(int* count) {
    if (!count) return fail(RTO_ERR_INVALID, "count is NULL");
    RTO_CUDA(cudaGetDeviceCount(count));
    return RTO_OK;
}

int rto_synchronize(void* stream) {
    if (stream) RTO_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    else RTO_CUDA(cudaDeviceSynchronize());
    return RTO_OK;
}

void rto_render_options_default(rto_render_options* o) {
    if (!o) return;
    o->step_size = 1e-4f;
    o->sigma_thresh = 1e-2f;
    o->stop_thresh = 1e-2f;
    o->background_brightness = 1.f;
    o->denoise = 1;
    o->spp = 1;
    o->enable_probe = 0;
}

// ---------------------------------------------------------------------------------------------------- tree
// argument checks shared by the dense and the quantised entry point; normalises basis_dim for RGBA
static int check_tree_args(int64_t capacity, int N, int data_dim, int format, int* basis_dim) {
    if (N != 2) return fail(RTO_ERR_UNSUPPORTED, "N == %d: only N = 2 octrees are supported (n3tree.cpp:273-275 warns the same)", N);
    if (capacity <= 0 || capacity >= (int64_t)1 << 28) return fail(RTO_ERR_INVALID, "capacity %lld out of range", (long long)capacity);
    if (format == RTO_FORMAT_SG || format == RTO_FORMAT_ASG)
        return fail(RTO_ERR_UNSUPPORTED, "SG/ASG data formats are not supported (SH and RGBA only)");
    if (format == RTO_FORMAT_SH) {
        const int b = *basis_dim;
        if (!(b == 1 || b == 4 || b == 9 || b == 16 || b == 25))
            return fail(RTO_ERR_INVALID, "SH basis_dim %d not in {1,4,9,16,25}", b);
        if (data_dim != 3 * b + 1) return fail(RTO_ERR_INVALID, "data_dim %d != 3*%d+1", data_dim, b);
    } else if (format == RTO_FORMAT_RGBA) {
        if (data_dim != 4) return fail(RTO_ERR_INVALID, "RGBA format needs data_dim 4, got %d", data_dim);
        *basis_dim = -1;
    } else {
        return fail(RTO_ERR_INVALID, "unknown data format %d", format);
    }
    return RTO_OK;
}

static int tree_from_source(rto_tree** out, const rto::TreeSource& src, int N, int format, int basis_dim,
                            const float offset[3], const float scale[3]) {
    rto_tree* t = new (std::nothrow) rto_tree;
    if (!t) return fail(RTO_ERR_NOMEM, "out of host memory");
    rto::TreeBuilt b;
    std::string err;
    int64_t launches = 0;
    const int rc = rto::build_tree_device(src, b, err, &launches);
    g_launches += launches;
    if (rc != RTO_OK) {
        delete t;
        return fail(rc, "%s", err.c_str());
    }
    t->nodes = b.nodes; t->payload = b.payload;
    t->grid_top = b.grid_top; t->grid_bricks = b.grid_bricks; t->grid_bricks8 = b.grid_bricks8;
    t->grid_leaf_top = b.grid_leaf_top; t->grid_leaf_bricks = b.grid_leaf_bricks; t->grid_top_m = b.grid_top_m;
    t->grid_K = b.grid_K; t->n_bricks = b.n_bricks;
    const int64_t n_entries = src.capacity * 8;
    rto_tree_info& I = t->info;
    I.capacity = src.capacity; I.N = N; I.data_dim = src.data_dim; I.format = format; I.basis_dim = basis_dim;
    I.max_depth = b.max_depth; I.n_leaves = b.n_leaves;
    I.node_bytes = n_entries * (int64_t)sizeof(uint32_t);
    I.payload_bytes = n_entries * (int64_t)b.stride * (int64_t)sizeof(__half);
    I.payload_stride_halfs = b.stride;
    I.grid_level = t->grid_K;
    I.n_bricks = t->n_bricks;
    I.grid_bytes = t->grid_K ? (int64_t)((((size_t)1 << (3 * t->grid_K)) + (size_t)t->n_bricks * 512) * sizeof(uint32_t) * (t->grid_leaf_top ? 2 : 1) +
                                         (size_t)t->n_bricks * 512 + (t->grid_top_m ? ((size_t)1 << (3 * t->grid_K)) * sizeof(uint32_t) : 0)) : 0;
    for (int i = 0; i < 3; ++i) { I.offset[i] = offset[i]; I.scale[i] = scale[i]; }
    I.ndc_width = -1.f; I.ndc_height = 0.f; I.ndc_focal = 0.f;
    *out = t;
    return RTO_OK;
}

int rto_tree_create(rto_tree** out, const int32_t* child, const void* data_f16, int64_t capacity, int N,
                    int data_dim, int format, int basis_dim, const float offset[3], const float scale[3]) {
    if (!out || !child || !data_f16 || !offset || !scale) return fail(RTO_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (int rc = check_tree_args(capacity, N, data_dim, format, &basis_dim)) return rc;
    rto::TreeSource src;
    src.child = child; src.capacity = capacity; src.data_dim = data_dim; src.data_f16 = data_f16;
    return tree_from_source(out, src, N, format, basis_dim, offset, scale);
}

int rto_tree_create_quantized(rto_tree** out, const int32_t* child, int64_t capacity, int N, int data_dim, int format,
                              int basis_dim, const float offset[3], const float scale[3], const void* quant_colors_f16,
                              const uint16_t* quant_map, int n_quant, const void* sigma_f16, const void* data_retained_f16,
                              int n_retained) {
    if (!out || !child || !offset || !scale || !sigma_f16) return fail(RTO_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (int rc = check_tree_args(capacity, N, data_dim, format, &basis_dim)) return rc;
    if (n_quant < 0 || n_retained < 0 || (n_quant > 0 && (!quant_colors_f16 || !quant_map)) || (n_retained > 0 && !data_retained_f16))
        return fail(RTO_ERR_INVALID, "NULL codebook / map / retained array");
    // n3tree.cpp:301-316 writes data[j + n_ret + k*n_basis], k < 3: the colour part must hold 3 * n_basis values
    if (n_quant + n_retained == 0 || 3 * (n_quant + n_retained) > data_dim - 1)
        return fail(RTO_ERR_INVALID, "quantised tree: 3*(%d codebooks + %d retained) does not fit data_dim-1 = %d", n_quant, n_retained, data_dim - 1);
    rto::TreeSource src;
    src.child = child; src.capacity = capacity; src.data_dim = data_dim;
    src.quant_colors = quant_colors_f16; src.quant_map = quant_map; src.sigma_f16 = sigma_f16;
    src.retained_f16 = data_retained_f16; src.n_q = n_quant; src.n_ret = n_retained;
    return tree_from_source(out, src, N, format, basis_dim, offset, scale);
}

int rto_tree_read_plane(const rto_tree* t, int plane, void* host_dst, size_t bytes) {
    if (!t || !host_dst) return fail(RTO_ERR_INVALID, "NULL argument");
    const void* src = nullptr;
    size_t have = 0;
    switch (plane) {
        case RTO_PLANE_NODES: src = t->nodes; have = (size_t)t->info.node_bytes; break;
        case RTO_PLANE_PAYLOAD: src = t->payload; have = (size_t)t->info.payload_bytes; break;
        case RTO_PLANE_GRID_TOP: src = t->grid_top; have = t->grid_K ? ((size_t)1 << (3 * t->grid_K)) * sizeof(uint32_t) : 0; break;
        case RTO_PLANE_GRID_BRICKS: src = t->grid_bricks; have = t->grid_K ? (size_t)t->n_bricks * 512 * sizeof(uint32_t) : 0; break;
        case RTO_PLANE_GRID_BRICKS8: src = t->grid_bricks8; have = t->grid_K ? (size_t)t->n_bricks * 512 : 0; break;
        case RTO_PLANE_GRID_LEAF_TOP: src = t->grid_leaf_top; have = t->grid_leaf_top ? ((size_t)1 << (3 * t->grid_K)) * sizeof(uint32_t) : 0; break;
        case RTO_PLANE_GRID_LEAF_BRICKS: src = t->grid_leaf_bricks; have = t->grid_leaf_top ? (size_t)t->n_bricks * 512 * sizeof(uint32_t) : 0; break;
        case RTO_PLANE_GRID_MARCH_TOP: src = t->grid_top_m; have = t->grid_top_m ? ((size_t)1 << (3 * t->grid_K)) * sizeof(uint32_t) : 0; break;
        default: return fail(RTO_ERR_INVALID, "unknown plane %d", plane);
    }
    if (bytes != have) return fail(RTO_ERR_INVALID, "plane %d holds %zu bytes, caller asked for %zu", plane, have, bytes);
    if (bytes) RTO_CUDA(cudaMemcpy(host_dst, src, bytes, cudaMemcpyDeviceToHost));
    return RTO_OK;
}

int rto_tree_set_ndc(rto_tree* t, float w, float h, float f) {
    if (!t) return fail(RTO_ERR_INVALID, "tree is NULL");
    t->info.ndc_width = w; t->info.ndc_height = h; t->info.ndc_focal = f;
    return RTO_OK;
}
int rto_tree_get_info(const rto_tree* t, rto_tree_info* info) {
    if (!t || !info) return fail(RTO_ERR_INVALID, "NULL argument");
    *info = t->info;
    return RTO_OK;
}
void rto_tree_destroy(rto_tree* t) {
    if (!t) return;
    cudaFree(t->nodes);
    cudaFree(t->payload);
    cudaFree(t->grid_top);
    cudaFree(t->grid_bricks);
    cudaFree(t->grid_bricks8);
    cudaFree(t->grid_leaf_top);
    cudaFree(t->grid_leaf_bricks);
    cudaFree(t->grid_top_m);
    delete t;
}

// ------------------------------------------------------------------------------------------------- context
int rto_context_create(rto_context** out, int W, int H) {
    if (!out) return fail(RTO_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (W <= 0 || H <= 0 || (int64_t)W * H > (int64_t)1 << 28) return fail(RTO_ERR_INVALID, "bad image size %dx%d", W, H);
    rto_context* c = new (std::nothrow) rto_context;
    if (!c) return fail(RTO_ERR_NOMEM, "out of host memory");
    c->W = W; c->H = H;
    const size_t px = (size_t)W * H;
    cudaError_t e = cudaMalloc(&c->aux, px * 8 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&c->img, px * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&c->weight_map, px * 6 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&c->guidance_map, px * 6 * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&c->adv, (size_t)(W + H) * sizeof(rto::AdvanceMap));
    if (e == cudaSuccess) e = cudaHostAlloc(&c->adv_host, (size_t)(W + H) * sizeof(rto::AdvanceMap), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->adv_copied, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&c->tile_counter, 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(c->tile_counter, 0, 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(c->aux, 0, px * 8 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(c->img, 0, px * sizeof(float4));
    for (int i = 0; i < 3 && e == cudaSuccess; ++i) {
        e = cudaEventCreate(&c->ev_start[i]);
        if (e == cudaSuccess) e = cudaEventCreate(&c->ev_stop[i]);
    }
    if (e != cudaSuccess) {
        rto_context_destroy(c);
        return fail(e == cudaErrorMemoryAllocation ? RTO_ERR_NOMEM : RTO_ERR_CUDA, "context allocation: %s", cudaGetErrorString(e));
    }
    pcg32_seed(c->rng, 20230418ull);  // render_context.hpp:16
    *out = c;
    return RTO_OK;
}
void rto_context_destroy(rto_context* c) {
    if (!c) return;
    cudaFree(c->aux); cudaFree(c->img); cudaFree(c->img8); cudaFree(c->weight_map); cudaFree(c->guidance_map); cudaFree(c->tile_counter); cudaFree(c->adv);
    if (c->adv_host) cudaFreeHost(c->adv_host);
    if (c->adv_copied) cudaEventDestroy(c->adv_copied);
    for (int i = 0; i < 3; ++i) {
        if (c->ev_start[i]) cudaEventDestroy(c->ev_start[i]);
        if (c->ev_stop[i]) cudaEventDestroy(c->ev_stop[i]);
    }
    delete c;
}
float* rto_context_aux(rto_context* c) { return c ? c->aux : nullptr; }
float* rto_context_image(rto_context* c) { return c ? reinterpret_cast<float*>(c->img) : nullptr; }
int rto_context_rng_seed(rto_context* c, uint64_t seed) {
    if (!c) return fail(RTO_ERR_INVALID, "ctx is NULL");
    pcg32_seed(c->rng, seed);
    return RTO_OK;
}
int rto_context_rng_advance(rto_context* c, int64_t delta) {
    if (!c) return fail(RTO_ERR_INVALID, "ctx is NULL");
    rto::pcg32_advance(c->rng, (uint64_t)delta);
    return RTO_OK;
}
int rto_context_rng_set_frame(rto_context* c, int64_t warmup, int64_t frame) {
    if (!c) return fail(RTO_ERR_INVALID, "ctx is NULL");
    pcg32_seed(c->rng, 20230418ull);
    // (warmup+frame) calls of advance(2^32) compose to one advance((warmup+frame)*2^32) (mod 2^64)
    rto::pcg32_advance(c->rng, (uint64_t)(warmup + frame) << 32);
    return RTO_OK;
}
int rto_context_rng_get(const rto_context* c, uint64_t* state, uint64_t* inc) {
    if (!c || !state || !inc) return fail(RTO_ERR_INVALID, "NULL argument");
    *state = c->rng.state; *inc = c->rng.inc;
    return RTO_OK;
}
int rto_context_read_aux(rto_context* c, float* dst, void* stream) {
    if (!c || !dst) return fail(RTO_ERR_INVALID, "NULL argument");
    RTO_CUDA(cudaMemcpyAsync(dst, c->aux, (size_t)c->W * c->H * 8 * sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return RTO_OK;
}
int rto_context_write_aux(rto_context* c, const float* src, void* stream) {
    if (!c || !src) return fail(RTO_ERR_INVALID, "NULL argument");
    RTO_CUDA(cudaMemcpyAsync(c->aux, src, (size_t)c->W * c->H * 8 * sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return RTO_OK;
}
int rto_context_read_image(rto_context* c, float* dst, void* stream) {
    if (!c || !dst) return fail(RTO_ERR_INVALID, "NULL argument");
    RTO_CUDA(cudaMemcpyAsync(dst, c->img, (size_t)c->W * c->H * sizeof(float4), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return RTO_OK;
}
static int ensure_img8(rto_context* c) {
    if (c->img8) return RTO_OK;
    RTO_CUDA(cudaMalloc(&c->img8, (size_t)c->W * c->H * sizeof(uchar4)));
    return RTO_OK;
}
// make the RGBA8 copy current on `stream` (no-op when the kernel that produced the image wrote it already)
static int refresh_img8(rto_context* c, cudaStream_t s) {
    if (int rc = ensure_img8(c)) return rc;
    if (c->img8_gen == c->img_gen && c->img_gen != 0) return RTO_OK;
    cudaError_t e = rto::launch_rgba8(c->img, c->img8, (size_t)c->W * c->H, s);
    if (e != cudaSuccess) return fail(RTO_ERR_CUDA, "rgba8 launch: %s", cudaGetErrorString(e));
    ++g_launches;
    c->img8_gen = c->img_gen;
    return RTO_OK;
}
int rto_context_read_image_rgba8(rto_context* c, unsigned char* dst, void* stream) {
    if (!c || !dst) return fail(RTO_ERR_INVALID, "NULL argument");
    if (int rc = refresh_img8(c, (cudaStream_t)stream)) return rc;
    RTO_CUDA(cudaMemcpyAsync(dst, c->img8, (size_t)c->W * c->H * sizeof(uchar4), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return RTO_OK;
}
int rto_context_read_rows_rgba8(rto_context* c, unsigned char* dst, int y0, int y1, void* stream) {
    if (!c || !dst) return fail(RTO_ERR_INVALID, "NULL argument");
    if (y0 < 0 || y1 > c->H || y0 >= y1) return fail(RTO_ERR_INVALID, "bad row range [%d, %d) for height %d", y0, y1, c->H);
    if (c->img_target) return fail(RTO_ERR_INVALID, "an image target is set: the rows were stored there, not on this context");
    if (!c->img8) return fail(RTO_ERR_INVALID, "no RGBA8 copy on this context: call rto_context_image_rgba8 before producing the rows");
    const size_t off = (size_t)y0 * c->W;
    RTO_CUDA(cudaMemcpyAsync(dst + off * sizeof(uchar4), c->img8 + off, (size_t)(y1 - y0) * c->W * sizeof(uchar4), cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
    return RTO_OK;
}
int rto_context_read_image_rows(rto_context* c, float* dst, int y0, int y1, void* stream) {
    if (!c || !dst) return fail(RTO_ERR_INVALID, "NULL argument");
    if (y0 < 0 || y1 > c->H || y0 >= y1) return fail(RTO_ERR_INVALID, "bad row range [%d, %d) for height %d", y0, y1, c->H);
    if (c->img_target) return fail(RTO_ERR_INVALID, "an image target is set: the rows were stored there, not on this context");
    const size_t off = (size_t)y0 * c->W;
    RTO_CUDA(cudaMemcpyAsync(dst + off * 4, c->img + off, (size_t)(y1 - y0) * c->W * sizeof(float4), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return RTO_OK;
}
unsigned char* rto_context_image_rgba8(rto_context* c) {
    if (!c || ensure_img8(c) != RTO_OK) return nullptr;
    return reinterpret_cast<unsigned char*>(c->img8);
}

// --------------------------------------------------------------------------------------------------- render
// (re)build the pcg32 jump-ahead tables when (W, spp, inc) changed: host arithmetic into the context's pinned staging
// buffer, then an ASYNC copy on the render stream (stream-ordered before the launch that needs it; no device-wide sync).
static int ensure_adv_tables(rto_context* c, int spp, cudaStream_t s) {
    if (c->adv_spp == spp && c->adv_inc == c->rng.inc) return RTO_OK;
    RTO_CUDA(cudaEventSynchronize(c->adv_copied));   // the previous upload (if any) has finished reading the staging buffer
    for (int y = 0; y < c->H; ++y) c->adv_host[y] = rto::pcg32_advance_map(c->rng.inc, (uint64_t)y * (uint64_t)c->W * (uint64_t)spp);
    for (int x = 0; x < c->W; ++x) c->adv_host[(size_t)c->H + x] = rto::pcg32_advance_map(c->rng.inc, (uint64_t)x * (uint64_t)spp);
    RTO_CUDA(cudaMemcpyAsync(c->adv, c->adv_host, ((size_t)c->H + c->W) * sizeof(rto::AdvanceMap), cudaMemcpyHostToDevice, s));
    RTO_CUDA(cudaEventRecord(c->adv_copied, s));
    c->adv_spp = spp;
    c->adv_inc = c->rng.inc;
    return RTO_OK;
}

// everything render_kernel needs for one frame, from the handles and PODs of the C ABI
static int fill_render_args(rto_context* c, const rto_tree* t, const rto_camera* cam, const rto_render_options* opt,
                            int x0, int y0, int x1, int y1, const rto_trace* trace, rto::RenderArgs& a) {
    if (!c || !t || !cam) return fail(RTO_ERR_INVALID, "NULL argument");
    int rc = check_opt(opt);
    if (rc != RTO_OK) return rc;
    if (cam->width != c->W || cam->height != c->H)
        return fail(RTO_ERR_INVALID, "camera %dx%d does not match context %dx%d", cam->width, cam->height, c->W, c->H);
    x0 = x0 < 0 ? 0 : x0; y0 = y0 < 0 ? 0 : y0;
    x1 = x1 > c->W ? c->W : x1; y1 = y1 > c->H ? c->H : y1;
    a = rto::RenderArgs{};
    rto::FrameParams& fp = a.fp;
    memcpy(fp.c2w, cam->c2w, sizeof fp.c2w);
    for (int i = 0; i < 3; ++i) { fp.offset[i] = t->info.offset[i]; fp.scale[i] = t->info.scale[i]; }
    fp.fx = cam->fx; fp.fy = cam->fy;
    fp.ndc_width = t->info.ndc_width; fp.ndc_height = t->info.ndc_height; fp.ndc_focal = t->info.ndc_focal;
    fp.step_size = opt->step_size; fp.sigma_thresh = opt->sigma_thresh; fp.background = opt->background_brightness;
    fp.W = c->W; fp.H = c->H;
    a.tree = rto::TreeDev{t->nodes, t->payload, t->info.payload_stride_halfs, t->info.basis_dim, t->info.max_depth,
                          rto::make_grid_dev(t->grid_top, t->grid_bricks, t->grid_K, t->grid_bricks8, t->grid_leaf_top, t->grid_leaf_bricks,
                                             rto::bias_march_table(t->grid_top_m, t->grid_K)),
                          (size_t)t->n_bricks * 512 * sizeof(uint32_t)};
    a.rng_state = c->rng.state; a.rng_inc = c->rng.inc;
    a.x0 = x0; a.y0 = y0; a.x1 = x1; a.y1 = y1;
    a.aux = c->aux;
    a.tile_counter = c->tile_counter;
    a.adv_rows = c->adv;
    a.adv_cols = c->adv + c->H;
    // with the denoiser on, the final image comes from rto_denoise; the reference then renders into a separate
    // noisy surface whose rgb equals aux channels 0..2 (volrend.cu:188-192 vs :205-212), so nothing is lost here
    a.img = opt->denoise ? nullptr : c->out_img();
    a.img8 = opt->denoise ? nullptr : c->out_img8();   // RGBA8 copy in the same store once a caller has asked for it
    if (trace) {
        a.tr = rto::TraceOut{trace->steps, trace->term, trace->src_bits, trace->t_bits, trace->leaf_hash,
                             trace->depth_sum, trace->n_hits, trace->n_loads, trace->hit_leaf, trace->hit_cnt,
                             trace->leaf_seq, trace->thresh, trace->max_seq};
    }
    return RTO_OK;
}

static int render_impl(rto_context* c, const rto_tree* t, const rto_camera* cam, const rto_render_options* opt,
                       int x0, int y0, int x1, int y1, const rto_trace* trace, void* stream) {
    rto::RenderArgs a;
    int rc = fill_render_args(c, t, cam, opt, x0, y0, x1, y1, trace, a);
    if (rc != RTO_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = ensure_adv_tables(c, opt->spp, s)) != RTO_OK) return rc;
    timer_start(c, 0, s);
    bool bad_spp = false;
    cudaError_t e = rto::launch_render(a, opt->spp, trace ? (trace->marcher == 1 ? 2 : 1) : 0, s, &bad_spp);
    timer_stop(c, 0, s);
    if (bad_spp) return fail(RTO_ERR_UNSUPPORTED, "spp == %d not supported.", opt->spp);
    if (e != cudaSuccess) return fail(RTO_ERR_CUDA, "render kernel launch: %s", cudaGetErrorString(e));
    ++g_launches;
    if (a.img) {   // a full-frame or band render produced (part of) a new image
        ++c->img_gen;
        if (a.img8 && !c->img_target && x0 <= 0 && y0 <= 0 && x1 >= c->W && y1 >= c->H) c->img8_gen = c->img_gen;
    }
    return RTO_OK;
}
int rto_render(rto_context* c, const rto_tree* t, const rto_camera* cam, const rto_render_options* opt, void* stream) {
    return render_impl(c, t, cam, opt, 0, 0, c ? c->W : 0, c ? c->H : 0, nullptr, stream);
}
int rto_render_rect(rto_context* c, const rto_tree* t, const rto_camera* cam, const rto_render_options* opt, int x0,
                    int y0, int x1, int y1, void* stream) {
    return render_impl(c, t, cam, opt, x0, y0, x1, y1, nullptr, stream);
}
int rto_render_trace(rto_context* c, const rto_tree* t, const rto_camera* cam, const rto_render_options* opt,
                     const rto_trace* trace, void* stream) {
    if (!trace) return fail(RTO_ERR_INVALID, "trace is NULL");
    return render_impl(c, t, cam, opt, 0, 0, c ? c->W : 0, c ? c->H : 0, trace, stream);
}

// ------------------------------------------------------------------------------------------------- denoiser
int rto_net_create(rto_net** out, const void* w1, const void* b1, const void* w2, const void* b2, int in_ch,
                   int mid_ch, int levels) {
    if (!out || !w1 || !b1 || !w2 || !b2) return fail(RTO_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (in_ch != 8) return fail(RTO_ERR_UNSUPPORTED, "in_channels %d: the aux buffer has 8 channels (RenderContext::CHANNELS)", in_ch);
    if (mid_ch < 1 || mid_ch > 64) return fail(RTO_ERR_UNSUPPORTED, "mid_channels %d not in 1..64", mid_ch);
    if (levels < 1 || levels > 6) return fail(RTO_ERR_UNSUPPORTED, "Kernel size == %d not supported.", levels * 2 + 1);  // filtering.cu:362-366
    rto_net* n = new (std::nothrow) rto_net;
    if (!n) return fail(RTO_ERR_NOMEM, "out of host memory");
    n->in_ch = in_ch; n->mid_ch = mid_ch; n->levels = levels;
    const size_t n1 = (size_t)mid_ch * in_ch * 9, n2 = (size_t)2 * levels * mid_ch * 9;
    cudaError_t e = cudaMalloc(&n->w1, n1 * 2);
    if (e == cudaSuccess) e = cudaMalloc(&n->b1, (size_t)mid_ch * 2);
    if (e == cudaSuccess) e = cudaMalloc(&n->w2, n2 * 2);
    if (e == cudaSuccess) e = cudaMalloc(&n->b2, (size_t)2 * levels * 2);
    if (e == cudaSuccess) e = cudaMemcpy(n->w1, w1, n1 * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(n->b1, b1, (size_t)mid_ch * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(n->w2, w2, n2 * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(n->b2, b2, (size_t)2 * levels * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && in_ch == 8 && mid_ch == 32 && levels == 4 && rto::denoise_tc_packed_bytes() > 0) {
        e = cudaMalloc(&n->packed, rto::denoise_tc_packed_bytes());
        if (e == cudaSuccess) e = rto::denoise_tc_pack_weights(n->dev(), n->packed, nullptr);
        if (e == cudaSuccess) { ++g_launches; e = cudaDeviceSynchronize(); }
    }
    if (e != cudaSuccess) {
        rto_net_destroy(n);
        return fail(e == cudaErrorMemoryAllocation ? RTO_ERR_NOMEM : RTO_ERR_CUDA, "net upload: %s", cudaGetErrorString(e));
    }
    *out = n;
    return RTO_OK;
}
void rto_net_destroy(rto_net* n) {
    if (!n) return;
    cudaFree(n->w1); cudaFree(n->b1); cudaFree(n->w2); cudaFree(n->b2); cudaFree(n->packed);
    delete n;
}
int rto_net_set_impl(rto_net* n, int impl) {
    if (!n) return fail(RTO_ERR_INVALID, "net is NULL");
    if (impl != 0 && impl != 1) return fail(RTO_ERR_INVALID, "impl must be 0 (auto) or 1 (simt)");
    n->impl = impl;
    return RTO_OK;
}

int rto_net_set_bias_mode(rto_net* n, int fused) {
    if (!n) return fail(RTO_ERR_INVALID, "net is NULL");
    n->fused_bias = fused != 0;
    return RTO_OK;
}

int rto_denoise_rows(rto_context* c, const rto_net* n, int y0, int y1, void* stream) {
    if (!c || !n) return fail(RTO_ERR_INVALID, "NULL argument");
    y0 = y0 < 0 ? 0 : y0; y1 = y1 > c->H ? c->H : y1;
    cudaStream_t s = (cudaStream_t)stream;
    // GuidanceNet rows must cover the filter's halo: the filter at row y reads guidance rows y-L..y+L
    const int L = n->levels;
    rto::DenoiseArgs dn{c->aux, c->out_img(), c->weight_map, c->guidance_map, c->W, c->H, y0 - L < 0 ? 0 : y0 - L,
                        y1 + L > c->H ? c->H : y1 + L};
    const bool tc = n->impl == 0 && n->tc_capable();
    timer_start(c, 1, s);
    // tensor-core path: the map written for the separable filter holds e^{guidance} (exponentiated once, in the net epilogue)
    cudaError_t e = tc ? rto::launch_guidance_net_tc(n->dev(), n->packed, dn, true, s) : rto::launch_guidance_net_simt(n->dev(), dn, s);
    timer_stop(c, 1, s);
    if (e != cudaSuccess) return fail(RTO_ERR_CUDA, "guidance net launch (%s): %s", tc ? "tcgen05" : "simt", cudaGetErrorString(e));
    timer_start(c, 2, s);
    if (tc)   // guidance is relu6-bounded here, so the precomputed-exp filter applies
        e = rto::launch_filter_fast(c->aux, c->weight_map, c->guidance_map, c->W, c->H, y0, y1, c->out_img(), c->out_img8(), s);
    else
        e = rto::launch_filter_simt(c->aux, (size_t)c->W * c->H, 1, c->weight_map, c->guidance_map, L, c->W, c->H, y0, y1,
                                    c->out_img(), s);
    timer_stop(c, 2, s);
    if (e != cudaSuccess) return fail(RTO_ERR_CUDA, "filter launch: %s", cudaGetErrorString(e));
    g_launches += 2;
    ++c->img_gen;
    if (tc && c->img8 && !c->img_target && y0 <= 0 && y1 >= c->H) c->img8_gen = c->img_gen;   // the filter wrote the RGBA8 copy as well
    return RTO_OK;
}
int rto_denoise(rto_context* c, const rto_net* n, void* stream) {
    return rto_denoise_rows(c, n, 0, c ? c->H : 0, stream);
}

int rto_net_forward(const rto_net* n, const float* aux_dev, int W, int H, float* weight_dev, float* guidance_dev,
                    void* stream) {
    if (!n || !aux_dev || !weight_dev || !guidance_dev) return fail(RTO_ERR_INVALID, "NULL argument");
    if (W <= 0 || H <= 0) return fail(RTO_ERR_INVALID, "bad size");
    rto::DenoiseArgs d{aux_dev, nullptr, weight_dev, guidance_dev, W, H, 0, H};
    cudaError_t e = (n->impl == 0 && n->tc_capable()) ? rto::launch_guidance_net_tc(n->dev(), n->packed, d, false, (cudaStream_t)stream)
                                                       : rto::launch_guidance_net_simt(n->dev(), d, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(RTO_ERR_CUDA, "guidance net launch: %s", cudaGetErrorString(e));
    ++g_launches;
    return RTO_OK;
}

int rto_filter(const float* weight_dev, const float* guidance_dev, const float* img_in_dev, int L, int W, int H,
               float* img_out_dev, void* stream) {
    if (!weight_dev || !guidance_dev || !img_in_dev || !img_out_dev) return fail(RTO_ERR_INVALID, "NULL argument");
    if (L < 1 || L > 6) return fail(RTO_ERR_UNSUPPORTED, "Kernel size == %d not supported.", L * 2 + 1);
    if (W <= 0 || H <= 0) return fail(RTO_ERR_INVALID, "bad size");
    cudaError_t e = rto::launch_filter_simt(img_in_dev, 1, 4, weight_dev, guidance_dev, L, W, H, 0, H,
                                            reinterpret_cast<float4*>(img_out_dev), (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(RTO_ERR_CUDA, "filter launch: %s", cudaGetErrorString(e));
    ++g_launches;
    return RTO_OK;
}

// ---- training side of the filter: forward that saves (rgb_filtered, max, 1/sum) per level, and the backward ----
int rto_filter_forward_save(const float* weight_dev, const float* guidance_dev, const float* img_in_dev, int L, int W, int H,
                            float* img_out_dev, float* rgb_filtered_dev, float* max_map_dev, float* inv_kernel_sum_dev,
                            void* stream) {
    if (!weight_dev || !guidance_dev || !img_in_dev || !img_out_dev || !rgb_filtered_dev || !max_map_dev || !inv_kernel_sum_dev)
        return fail(RTO_ERR_INVALID, "NULL argument");
    if (L < 1 || L > 6) return fail(RTO_ERR_UNSUPPORTED, "Kernel size == %d not supported.", L * 2 + 1);
    if (W <= 0 || H <= 0) return fail(RTO_ERR_INVALID, "bad size");
    cudaError_t e = rto::launch_filter_simt(img_in_dev, 1, 4, weight_dev, guidance_dev, L, W, H, 0, H,
                                            reinterpret_cast<float4*>(img_out_dev), (cudaStream_t)stream,
                                            reinterpret_cast<float4*>(rgb_filtered_dev), max_map_dev, inv_kernel_sum_dev);
    if (e != cudaSuccess) return fail(RTO_ERR_CUDA, "filter launch: %s", cudaGetErrorString(e));
    ++g_launches;
    return RTO_OK;
}

int rto_filter_backward(const float* grad_output_dev, const float* img_in_dev, const float* weight_dev, const float* guidance_dev,
                        const float* rgb_filtered_dev, const float* max_map_dev, const float* inv_kernel_sum_dev, int L, int W,
                        int H, float* grad_weight_dev, float* grad_guidance_dev, void* stream) {
    if (!grad_output_dev || !img_in_dev || !weight_dev || !guidance_dev || !rgb_filtered_dev || !max_map_dev ||
        !inv_kernel_sum_dev || !grad_weight_dev || !grad_guidance_dev)
        return fail(RTO_ERR_INVALID, "NULL argument");
    if (L < 1 || L > 6) return fail(RTO_ERR_UNSUPPORTED, "Kernel size == %d not supported.", L * 2 + 1);
    if (W <= 0 || H <= 0) return fail(RTO_ERR_INVALID, "bad size");
    cudaError_t e = rto::launch_filter_backward(grad_output_dev, img_in_dev, weight_dev, guidance_dev, rgb_filtered_dev, max_map_dev,
                                                inv_kernel_sum_dev, L, W, H, grad_weight_dev, grad_guidance_dev, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(RTO_ERR_CUDA, "filter backward launch: %s", cudaGetErrorString(e));
    ++g_launches;
    return RTO_OK;
}


// --------------------------------------------------------------------------- tile split: peer-direct image stores
int rto_context_set_image_target(rto_context* c, float* image_dev, unsigned char* rgba8_dev) {
    if (!c) return fail(RTO_ERR_INVALID, "ctx is NULL");
    if (!image_dev && rgba8_dev) return fail(RTO_ERR_INVALID, "an RGBA8 target needs an image target");
    c->img_target = reinterpret_cast<float4*>(image_dev);
    c->img8_target = reinterpret_cast<uchar4*>(rgba8_dev);
    return RTO_OK;
}
int rto_context_mark_image_written(rto_context* c, int rgba8_too) {
    if (!c) return fail(RTO_ERR_INVALID, "ctx is NULL");
    ++c->img_gen;
    if (rgba8_too && c->img8) c->img8_gen = c->img_gen;
    return RTO_OK;
}
int rto_peer_enable(int peer_device) {
    int dev = 0;
    RTO_CUDA(cudaGetDevice(&dev));
    if (dev == peer_device) return RTO_OK;
    int can = 0;
    RTO_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
    if (!can) return fail(RTO_ERR_UNSUPPORTED, "device %d cannot access device %d directly", dev, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return RTO_OK; }
    if (e != cudaSuccess) return fail(RTO_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
    return RTO_OK;
}
int rto_ipc_export(const void* dev_ptr, unsigned char handle[64]) {
    if (!dev_ptr || !handle) return fail(RTO_ERR_INVALID, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    RTO_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    memcpy(handle, &h, 64);
    return RTO_OK;
}
int rto_ipc_open(const unsigned char handle[64], void** dev_ptr) {
    if (!dev_ptr || !handle) return fail(RTO_ERR_INVALID, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    RTO_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RTO_OK;
}
int rto_ipc_close(void* dev_ptr) {
    if (dev_ptr) RTO_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return RTO_OK;
}
int rto_event_create(void** event) {
    if (!event) return fail(RTO_ERR_INVALID, "event is NULL");
    cudaEvent_t e = nullptr;
    RTO_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *event = (void*)e;
    return RTO_OK;
}
int rto_event_create_timed(void** event) {
    if (!event) return fail(RTO_ERR_INVALID, "event is NULL");
    cudaEvent_t e = nullptr;
    RTO_CUDA(cudaEventCreate(&e));
    *event = (void*)e;
    return RTO_OK;
}
int rto_event_elapsed_ms(void* start, void* end, float* ms) {
    if (!start || !end || !ms) return fail(RTO_ERR_INVALID, "NULL argument");
    RTO_CUDA(cudaEventSynchronize((cudaEvent_t)end));
    RTO_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)end));
    return RTO_OK;
}
int rto_event_record(void* event, void* stream) {
    RTO_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return RTO_OK;
}
int rto_stream_wait_event(void* stream, void* event) {
    RTO_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
    return RTO_OK;
}
int rto_event_destroy(void* event) {
    if (event) RTO_CUDA(cudaEventDestroy((cudaEvent_t)event));
    return RTO_OK;
}

// ------------------------------------------------------------------------------- streams, pinned memory, frame graph
int rto_stream_create(void** stream) {
    if (!stream) return fail(RTO_ERR_INVALID, "stream is NULL");
    cudaStream_t s = nullptr;
    RTO_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void*)s;
    return RTO_OK;
}
int rto_stream_destroy(void* stream) {
    if (stream) RTO_CUDA(cudaStreamDestroy((cudaStream_t)stream));
    return RTO_OK;
}
int rto_host_alloc(void** ptr, size_t bytes) {
    if (!ptr) return fail(RTO_ERR_INVALID, "ptr is NULL");
    *ptr = nullptr;
    cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? RTO_ERR_NOMEM : RTO_ERR_CUDA, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    return RTO_OK;
}
int rto_host_free(void* ptr) {
    if (ptr) RTO_CUDA(cudaFreeHost(ptr));
    return RTO_OK;
}

}  // extern "C"

// One frame of the hot path as an instantiated CUDA graph: render -> GuidanceNet -> filter (+RGBA8) -> device->host copies.
// The graph is captured once from the very launches rto_render / rto_denoise / rto_context_read_* make; per frame only the
// render kernel's by-value argument block (camera transform, rng state) is replaced (cudaGraphExecKernelNodeSetParams) and
// the whole frame goes to the GPU with ONE cudaGraphLaunch.
struct rto_frame {
    rto_context* ctx = nullptr;
    const rto_tree* tree = nullptr;
    const rto_net* net = nullptr;
    rto_render_options opt{};
    rto_camera cam{};
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaGraphNode_t render_node = nullptr;
    cudaKernelNodeParams kp{};
    rto::RenderArgs args{};
    void* kparams[1] = {nullptr};
    int launches_per_frame = 0;
    int64_t pending = -1;   // rto_frame_sequence: index of the frame in flight on this slot (-1: none)
};

extern "C" {

int rto_frame_create(rto_frame** out, rto_context* c, const rto_frame_desc* d) {
    if (!out || !c || !d || !d->tree) return fail(RTO_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (int rc = check_opt(&d->opt)) return rc;
    if (d->opt.denoise && !d->net) return fail(RTO_ERR_INVALID, "options.denoise is set but no net was given");
    rto_frame* f = new (std::nothrow) rto_frame;
    if (!f) return fail(RTO_ERR_NOMEM, "out of host memory");
    f->ctx = c; f->tree = d->tree; f->net = d->net; f->opt = d->opt;
    f->cam.width = c->W; f->cam.height = c->H; f->cam.fx = d->fx; f->cam.fy = d->fy;
    for (int i = 0; i < 12; ++i) f->cam.c2w[i] = 0.f;   // identity rotation at the origin; replaced at every launch
    f->cam.c2w[0] = f->cam.c2w[4] = f->cam.c2w[8] = 1.f;
    cudaStream_t cap = nullptr;
    int rc = RTO_OK;
    auto body = [&](cudaStream_t s) -> int {
        int r = render_impl(c, d->tree, &f->cam, &d->opt, 0, 0, c->W, c->H, nullptr, s);
        if (r == RTO_OK && d->opt.denoise) r = rto_denoise(c, d->net, s);
        if (r == RTO_OK && d->host_rgba8) r = rto_context_read_image_rgba8(c, d->host_rgba8, s);
        if (r == RTO_OK && d->host_image) r = rto_context_read_image(c, d->host_image, s);
        if (r == RTO_OK && d->host_aux) r = rto_context_read_aux(c, d->host_aux, s);
        return r;
    };
    cudaError_t e = cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete f; return fail(RTO_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    // 1. one uncaptured frame: runs every lazy initialisation (function attributes, L2 set-aside, jump-ahead tables, the
    //    RGBA8 buffer) that is not allowed inside a capture, and validates the arguments with ordinary error reporting
    if (d->host_rgba8) rc = ensure_img8(c);
    if (rc == RTO_OK) rc = body(cap);
    if (rc == RTO_OK && cudaStreamSynchronize(cap) != cudaSuccess) rc = fail(RTO_ERR_CUDA, "frame warm-up failed: %s", cudaGetErrorString(cudaGetLastError()));
    // 2. the same calls again, captured
    if (rc == RTO_OK) {
        const int64_t l0 = g_launches.load();
        e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) rc = fail(RTO_ERR_CUDA, "cudaStreamBeginCapture: %s", cudaGetErrorString(e));
        if (rc == RTO_OK) {
            c->capturing = true;
            rc = body(cap);
            c->capturing = false;
            e = cudaStreamEndCapture(cap, &f->graph);
            if (rc == RTO_OK && e != cudaSuccess) rc = fail(RTO_ERR_CUDA, "cudaStreamEndCapture: %s (host destinations must be pinned: rto_host_alloc)", cudaGetErrorString(e));
        }
        f->launches_per_frame = (int)(g_launches.load() - l0);
        g_launches -= f->launches_per_frame;   // captured, not launched
    }
    if (rc == RTO_OK) {
        e = cudaGraphInstantiate(&f->exec, f->graph, 0);
        if (e != cudaSuccess) rc = fail(RTO_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    if (rc == RTO_OK) {   // the render kernel is the only root of the (linear) graph
        cudaGraphNode_t roots[4];
        size_t n = 4;
        e = cudaGraphGetRootNodes(f->graph, roots, &n);
        cudaGraphNodeType ty = cudaGraphNodeTypeEmpty;
        if (e == cudaSuccess && n == 1) e = cudaGraphNodeGetType(roots[0], &ty);
        if (e != cudaSuccess || n != 1 || ty != cudaGraphNodeTypeKernel) rc = fail(RTO_ERR_CUDA, "frame graph: unexpected root (%zu roots, type %d): %s", n, (int)ty, cudaGetErrorString(e));
        else {
            f->render_node = roots[0];
            e = cudaGraphKernelNodeGetParams(f->render_node, &f->kp);
            if (e != cudaSuccess) rc = fail(RTO_ERR_CUDA, "cudaGraphKernelNodeGetParams: %s", cudaGetErrorString(e));
        }
    }
    cudaStreamDestroy(cap);
    if (rc != RTO_OK) {
        std::string keep = g_err;
        rto_frame_destroy(f);
        g_err = keep;
        return rc;
    }
    f->kparams[0] = &f->args;
    f->kp.kernelParams = f->kparams;
    f->kp.extra = nullptr;
    *out = f;
    return RTO_OK;
}

int rto_frame_launch(rto_frame* f, const float c2w[12], void* stream) {
    if (!f || !c2w) return fail(RTO_ERR_INVALID, "NULL argument");
    rto_context* c = f->ctx;
    memcpy(f->cam.c2w, c2w, sizeof f->cam.c2w);
    if (c->adv_spp != f->opt.spp || c->adv_inc != c->rng.inc)
        return fail(RTO_ERR_INVALID, "the context's spp / rng stream changed since rto_frame_create; create a new frame");
    int rc = fill_render_args(c, f->tree, &f->cam, &f->opt, 0, 0, c->W, c->H, nullptr, f->args);
    if (rc != RTO_OK) return rc;
    RTO_CUDA(cudaGraphExecKernelNodeSetParams(f->exec, f->render_node, &f->kp));
    RTO_CUDA(cudaGraphLaunch(f->exec, (cudaStream_t)stream));
    g_launches += f->launches_per_frame;
    ++c->img_gen;
    if (c->img8) c->img8_gen = c->img_gen;
    return RTO_OK;
}

int rto_frame_launch_indexed(rto_frame* f, const float c2w[12], int64_t warmup, int64_t frame, void* stream) {
    if (!f) return fail(RTO_ERR_INVALID, "NULL argument");
    if (int rc = rto_context_rng_set_frame(f->ctx, warmup, frame)) return rc;
    return rto_frame_launch(f, c2w, stream);
}

int rto_frame_sequence(rto_frame* const* frames, void* const* streams, int n_slots, const float* c2w, int64_t n_poses,
                       int64_t warmup, int64_t first, int64_t count, int drain, rto_frame_retired_fn retired, void* user) {
    if (!frames || !streams || !c2w || n_slots <= 0 || n_poses <= 0 || first < 0 || count < 0) return fail(RTO_ERR_INVALID, "bad argument");
    for (int k = 0; k < n_slots; ++k)
        if (!frames[k]) return fail(RTO_ERR_INVALID, "frames[%d] is NULL", k);
    auto retire = [&](int k) -> int {   // slot k's previous frame is complete on the host
        RTO_CUDA(cudaStreamSynchronize((cudaStream_t)streams[k]));
        rto_frame* f = frames[k];
        if (f->pending >= 0) {
            const int64_t done = f->pending;
            f->pending = -1;
            if (retired) retired(user, done, k);
        }
        return RTO_OK;
    };
    for (int64_t i = first; i < first + count; ++i) {
        const int k = (int)(i % n_slots);
        if (int rc = retire(k)) return rc;
        if (int rc = rto_frame_launch_indexed(frames[k], c2w + 12 * (i % n_poses), warmup, i, streams[k])) return rc;
        frames[k]->pending = i;
    }
    if (drain) {
        // oldest first: the slots after the last one issued
        for (int j = 1; j <= n_slots; ++j)
            if (int rc = retire((int)((first + count - 1 + j) % n_slots))) return rc;
    }
    return RTO_OK;
}

void rto_frame_destroy(rto_frame* f) {
    if (!f) return;
    if (f->exec) cudaGraphExecDestroy(f->exec);
    if (f->graph) cudaGraphDestroy(f->graph);
    delete f;
}

// ---------------------------------------------------------------------------------------------------- timer
int rto_timer_enable(rto_context* c, int enable) {
    if (!c) return fail(RTO_ERR_INVALID, "ctx is NULL");
    c->timing = enable != 0;
    return RTO_OK;
}
int rto_timer_reset(rto_context* c) {
    if (!c) return fail(RTO_ERR_INVALID, "ctx is NULL");
    c->frames = 0;
    for (int i = 0; i < 3; ++i) { c->sum_ms[i] = 0.f; c->ev_used[i] = false; }
    return RTO_OK;
}
int rto_timer_record(rto_context* c, int denoise) {  // Timer::record, render_context.hpp:179-188
    if (!c) return fail(RTO_ERR_INVALID, "ctx is NULL");
    if (!c->timing) return fail(RTO_ERR_INVALID, "timer not enabled");
    const int last = denoise ? 2 : 0;
    if (!c->ev_used[last]) return fail(RTO_ERR_INVALID, "nothing recorded for this frame");
    RTO_CUDA(cudaEventSynchronize(c->ev_stop[last]));
    c->frames++;
    for (int i = 0; i < 3; ++i) {
        if (!c->ev_used[i] || (!denoise && i > 0)) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev_start[i], c->ev_stop[i]) == cudaSuccess) c->sum_ms[i] += ms;
    }
    return RTO_OK;
}
int rto_timer_report(const rto_context* c, float ms[3], int* frames) {
    if (!c || !ms) return fail(RTO_ERR_INVALID, "NULL argument");
    for (int i = 0; i < 3; ++i) ms[i] = c->frames ? c->sum_ms[i] / c->frames : 0.f;
    if (frames) *frames = c->frames;
    return RTO_OK;
}

}  // extern "C"
