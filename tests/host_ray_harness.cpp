// TEST-ONLY host instantiation of rt_octree_b200/csrc/rto_ray.cuh (the header the CUDA kernels are built from).
// It lets the CPU test-suite check the integer-coordinate / ancestor-resume traversal, the exact op sequence
// and the SoA node-word encoding against oracle/rt_oracle.c without a GPU.  It is NOT linked into the product
// library and the product has no CPU path.  Build: g++ -O2 -ffp-contract=off -mfma -mf16c (tests/conftest.py).
#include <cstdint>
#include <vector>

#include "../rt_octree_b200/csrc/rto_grid_host.h"
#include "../rt_octree_b200/csrc/rto_ray.cuh"

using namespace rto;

namespace {
template <int SPP>
struct HostRay {   // per-ray scratch (shared memory on the GPU: SmemRay in rto_render.cu)
    std::vector<uint32_t> stk;
    float d[SPP + 1];
    uint32_t hl[SPP];
    float hc[SPP];
    uint32_t& stack(int l) { return stk[l]; }
    float& dst(int i) { return d[i]; }
    uint32_t& hit_leaf(int i) { return hl[i]; }
    float& hit_cnt(int i) { return hc[i]; }
    float sc[3];
    float& scratch(int i) { return sc[i]; }
};

template <int SPP>
void run(const uint32_t* nodes, const GridDev* grid, int max_depth, const FrameParams& fp, uint64_t rng_state, uint64_t rng_inc,
         int pix_begin, int pix_end, const float* thresh, uint32_t* steps, int32_t* term, uint32_t* src_bits,
         uint32_t* t_bits, uint64_t* leaf_hash, uint32_t* depth_sum, uint32_t* n_hits, uint32_t* n_loads,
         int32_t* hit_leaf, uint32_t* hit_cnt, int32_t* leaf_seq, int max_seq) {
    HostRay<SPP> mem;
    mem.stk.assign(max_depth + 1, 0u);
    for (int idx = pix_begin; idx < pix_end; ++idx) {
        const size_t r = (size_t)(idx - pix_begin);
        RaySetup rs;
        setup_ray(fp, idx % fp.W, idx / fp.W, rs);
        if (thresh) {
            for (int i = 0; i < SPP; ++i) mem.d[i] = thresh[r * SPP + i];
            mem.d[SPP] = FLT_MAX;
        } else {
            sorted_thresholds<SPP>(rng_state, rng_inc, idx, mem);
        }
        WalkOut wo;
        auto sink = [&](uint32_t step, uint32_t leaf) {
            if (leaf_seq && (int)step < max_seq) leaf_seq[r * max_seq + step] = (int32_t)leaf;
        };
        if (grid && grid->bricks8 && grid->leaf_top && grid->top_m) {   // PRODUCTION: fused-index marcher (byte plane, deferred hits)
            switch (grid->K) {
#define FK(KK) case KK: walk_grid_fused<SPP, true, KK>(nodes, *grid, mem, rs, fp.step_size, fp.sigma_thresh, wo, sink); \
                        resolve_hits_fused<SPP, KK>(*grid, mem, wo.n_hits); break;
                FK(1) FK(2) FK(3) FK(4) FK(5) FK(6) FK(7) FK(8)
#undef FK
            }
        } else if (grid && grid->bricks8 && grid->leaf_top) {   // v9: byte plane + deferred leaf look-up, shift-built indices
            walk_grid<SPP, true, true, true>(nodes, *grid, mem, rs, fp.step_size, fp.sigma_thresh, wo, sink);
            resolve_hits<SPP>(*grid, mem, wo.n_hits);
        } else if (grid && grid->bricks8)
            walk_grid<SPP, true, true>(nodes, *grid, mem, rs, fp.step_size, fp.sigma_thresh, wo, sink);
        else if (grid)
            walk_grid<SPP, true, false>(nodes, *grid, mem, rs, fp.step_size, fp.sigma_thresh, wo, sink);
        else
            walk<SPP, true>(nodes, mem, rs, fp.step_size, fp.sigma_thresh, wo, sink);
        steps[r] = wo.steps; term[r] = wo.term; src_bits[r] = u_bits(wo.src); t_bits[r] = u_bits(wo.t);
        leaf_hash[r] = wo.hash; depth_sum[r] = wo.depth_sum; n_hits[r] = wo.n_hits; n_loads[r] = wo.n_loads;
        for (int i = 0; i < SPP; ++i) {
            const bool live = i < (int)wo.n_hits;
            hit_leaf[r * SPP + i] = live ? (int32_t)mem.hl[i] : -1;
            hit_cnt[r * SPP + i] = live ? (uint32_t)mem.hc[i] : 0u;
        }
        if (leaf_seq)
            for (int s = (int)wo.steps; s < max_seq; ++s) leaf_seq[r * max_seq + s] = -1;
    }
}
}  // namespace

static std::vector<uint32_t> g_top, g_bricks, g_leaf_top, g_leaf_bricks, g_top_m;
static std::vector<uint8_t> g_bricks8;
static GridDev g_grid{nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr};
static bool g_fused = true;
static bool g_grid_on = false;
static bool g_byte_bricks = true;
static bool g_defer = true;

// 1 (default): walk_grid reads the byte plane like the production kernel; 0: the 4-byte leaf words (RTO_GRID8=0 variant)
extern "C" void host_ray_use_byte_bricks(int on) {
    g_byte_bricks = on != 0;
    g_grid.bricks8 = g_byte_bricks && !g_bricks8.empty() ? g_bricks8.data() : nullptr;
}

// 1 (default): collisions record a cell reference, resolved through the leaf-id planes after the march (production);
// 0: root descent at every collision
extern "C" void host_ray_use_deferred_hits(int on) {
    g_defer = on != 0;
    g_grid.leaf_top = g_defer && !g_leaf_top.empty() ? g_leaf_top.data() : nullptr;
    g_grid.leaf_bricks = g_defer && !g_leaf_bricks.empty() ? g_leaf_bricks.data() : nullptr;
}

// 1 (default): the fused-index marcher (production, needs byte bricks + deferred hits); 0: the v9 shift-built indices
extern "C" void host_ray_use_fused_index(int on) {
    g_fused = on != 0;
    g_grid.top_m = g_fused && !g_top_m.empty() ? bias_march_table(g_top_m.data(), g_grid.K) : nullptr;
}

// Build (or drop, child == NULL) the sparse brick grid used by subsequent host_ray_walk calls.
// Returns K (0 = not built for this depth), n_bricks through *n_bricks.
extern "C" int host_ray_set_grid(const int32_t* child, const uint16_t* data, int data_dim, int64_t capacity, int max_depth,
                                 int64_t* n_bricks) {
    g_grid_on = false;
    if (!child) return 0;
    int K = 0;
    if (!build_grid_host(child, data, data_dim, capacity, max_depth, g_top, g_bricks, K)) return 0;
    if (g_bricks.empty()) g_bricks.assign(512, 0u);
    grid_bytes_host(g_bricks, g_bricks8);
    build_grid_leaf_host(child, K, g_top, g_bricks.size() / 512, g_leaf_top, g_leaf_bricks);
    build_march_top_host(g_top, K, g_top_m);
    g_grid = make_grid_dev(g_top.data(), g_bricks.data(), K, g_byte_bricks ? g_bricks8.data() : nullptr,
                           g_defer ? g_leaf_top.data() : nullptr, g_defer ? g_leaf_bricks.data() : nullptr,
                           g_fused ? bias_march_table(g_top_m.data(), K) : nullptr);
    g_grid_on = true;
    if (n_bricks) *n_bricks = (int64_t)(g_bricks.size() / 512);
    return K;
}

extern "C" int host_ray_walk(const uint32_t* nodes, int max_depth, const float* c2w12, const float* offset,
                             const float* scale, float fx, float fy, float ndc_w, float ndc_h, float ndc_f,
                             float step_size, float sigma_thresh, int W, int H, int spp, uint64_t rng_state,
                             uint64_t rng_inc, int pix_begin, int pix_end, const float* thresh, uint32_t* steps,
                             int32_t* term, uint32_t* src_bits, uint32_t* t_bits, uint64_t* leaf_hash,
                             uint32_t* depth_sum, uint32_t* n_hits, uint32_t* n_loads, int32_t* hit_leaf,
                             uint32_t* hit_cnt, int32_t* leaf_seq, int max_seq) {
    FrameParams fp{};
    for (int i = 0; i < 12; ++i) fp.c2w[i] = c2w12[i];
    for (int i = 0; i < 3; ++i) { fp.offset[i] = offset[i]; fp.scale[i] = scale[i]; }
    fp.fx = fx; fp.fy = fy; fp.ndc_width = ndc_w; fp.ndc_height = ndc_h; fp.ndc_focal = ndc_f;
    fp.step_size = step_size; fp.sigma_thresh = sigma_thresh; fp.background = 1.f; fp.W = W; fp.H = H;
#define CASE(S) case S: run<S>(nodes, g_grid_on ? &g_grid : nullptr, max_depth, fp, rng_state, rng_inc, pix_begin, pix_end, thresh, steps, term, src_bits, t_bits, leaf_hash, depth_sum, n_hits, n_loads, hit_leaf, hit_cnt, leaf_seq, max_seq); return 0;
    switch (spp) { CASE(1) CASE(2) CASE(3) CASE(4) CASE(6) CASE(8) CASE(16) CASE(32) default: return -1; }
#undef CASE
}
