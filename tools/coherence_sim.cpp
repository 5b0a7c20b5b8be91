// tools/coherence_sim.cpp — ANALYSIS TOOL (not linked into the product): a CPU model of the marching loop's memory
// coherence.  It marches the 32 rays of every 8x4 warp tile with the kernels' own per-ray code (rto_ray.cuh, host build)
// and, for every loop iteration k, counts the distinct 32-byte sectors and 128-byte lines that the warp's table load and
// brick load touch under alternative memory layouts of the same grid.  Calibration target: ncu's per-instruction
// `L1 Tag Requests Global` / `L2 Theoretical Sectors Global` of the shipped kernel.
// Build: g++ -O2 -std=c++17 -ffp-contract=off -mfma -mf16c -fPIC -shared tools/coherence_sim.cpp -o build/libcoherence_sim.so
#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../rt_octree_b200/csrc/rto_grid_host.h"
#include "../rt_octree_b200/csrc/rto_ray.cuh"

using namespace rto;

namespace {
constexpr int SPP = 6;
struct Mem {
    float d[SPP + 1];
    float& dst(int i) { return d[i]; }
};
struct Step { uint32_t x, y, z; uint32_t brick; };   // finest-level cell coordinates (K+3 bits each); brick = 0xffffffff: table leaf

// spread the low 10 bits of v so that there are two zero bits between each
inline uint32_t part3(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x30000ffu;
    v = (v | (v << 8)) & 0x300f00fu;
    v = (v | (v << 4)) & 0x30c30c3u;
    v = (v | (v << 2)) & 0x9249249u;
    return v;
}
inline uint32_t morton3(uint32_t x, uint32_t y, uint32_t z) { return (part3(x) << 2) | (part3(y) << 1) | part3(z); }

struct Layout {
    const char* name;
    int table_morton;   // table index order: 0 = x-major (shipped), 1 = Morton
    int cell_bits;      // bits per brick cell: 32 (v7), 8 (v8), 4
    int cell_order;     // 0 = x-major inside the brick (shipped), 1 = Morton inside the brick
};
const Layout kLayouts[] = {
    {"u32 cells, x-major (v7)", 0, 32, 0}, {"u8 cells, x-major (v8, shipped)", 0, 8, 0}, {"u8 cells, Morton in brick", 0, 8, 1},
    {"u4 cells, x-major", 0, 4, 0},        {"u4 cells, Morton in brick", 0, 4, 1},       {"u8 cells + Morton table", 1, 8, 0},
};
constexpr int kNL = sizeof(kLayouts) / sizeof(kLayouts[0]);

inline size_t count_distinct(std::vector<uint64_t>& v) {
    std::sort(v.begin(), v.end());
    return (size_t)(std::unique(v.begin(), v.end()) - v.begin());
}
}  // namespace

// out[layout][0..5] = warp-level table loads, table lines, table sectors, warp-level brick loads, brick lines, brick sectors
extern "C" int coherence_sim(const int32_t* child, const uint16_t* data, int data_dim, int64_t capacity, int max_depth,
                             const float* c2w12, const float* offset, const float* scale, float fx, float fy, int W, int H,
                             uint64_t rng_state, uint64_t rng_inc, int tile_w, int tile_h, int tile_stride, double* out,
                             int* n_layouts, double* steps_total) {
    std::vector<uint32_t> top, bricks;
    int K = 0;
    if (!build_grid_host(child, data, data_dim, capacity, max_depth, top, bricks, K)) return -1;
    if (bricks.empty()) bricks.assign(512, 0u);
    const GridDev g = make_grid_dev(top.data(), bricks.data(), K);
    FrameParams fp{};
    for (int i = 0; i < 12; ++i) fp.c2w[i] = c2w12[i];
    for (int i = 0; i < 3; ++i) { fp.offset[i] = offset[i]; fp.scale[i] = scale[i]; }
    fp.fx = fx; fp.fy = fy; fp.ndc_width = -1.f; fp.step_size = 1e-4f; fp.sigma_thresh = 1e-2f; fp.background = 1.f;
    fp.W = W; fp.H = H;
    const SigmaThresh sth = sigma_thresh_half(fp.sigma_thresh);
    *n_layouts = kNL;
    std::fill(out, out + kNL * 6, 0.0);
    double steps = 0;
    const int lanes = tile_w * tile_h;
    std::vector<std::vector<Step>> ray(lanes);
    std::vector<uint64_t> tl, ts, bl, bs;
    int tile_no = 0;
    for (int ty = 0; ty + tile_h <= H; ty += tile_h)
        for (int tx = 0; tx + tile_w <= W; tx += tile_w) {
            if (tile_no++ % tile_stride) continue;   // sample every tile_stride-th tile
            size_t longest = 0;
            for (int l = 0; l < lanes; ++l) {
                ray[l].clear();
                const int ix = tx + l % tile_w, iy = ty + l / tile_w, idx = iy * W + ix;
                RaySetup rs;
                setup_ray(fp, ix, iy, rs);
                if (!rs.hit) continue;
                Mem mem;
                sorted_thresholds<SPP>(rng_state, rng_inc, idx, mem);
                float t = rs.tmin, src = 0.f;
                int nspp = 0;
                while (t < rs.tmax) {   // walk_grid (rto_ray.cuh), minus the bookkeeping
                    float p[3];
                    for (int k = 0; k < 3; ++k) p[k] = f_fma_clamp01(t, rs.dir[k], rs.cen[k]);
                    const uint32_t bx = coord_bits(p[0]), by = coord_bits(p[1]), bz = coord_bits(p[2]);
                    uint32_t nl = 0;
                    const uint32_t sh = 23 - (K + 3);
                    const uint32_t cx = (bx & 0x7fffffu) >> sh, cy = (by & 0x7fffffu) >> sh, cz = (bz & 0x7fffffu) >> sh;
                    const uint32_t e = g.top[(((size_t)(cx >> 3) << K) | (cy >> 3)) << K | (cz >> 3)];
                    const uint32_t word = grid_lookup<false>(g, bx, by, bz, nl);
                    ray[l].push_back(Step{cx, cy, cz, (e & RTO_LEAF_FLAG) ? 0xffffffffu : e});
                    const uint32_t cube_bits = word & 0x7f800000u;
                    const float dt = step_length_cs(p, rs.invdir, rs.addk, f_bits(cube_bits), f_bits(0x7f000000u - cube_bits), fp.step_size);
                    if (sigma_above(word, sth)) {
                        const float s_new = f_fma(f_mul(rs.delta_scale, dt), f_half_bits_to_float(word & 0xffffu), src);
                        src = s_new;
                        if (s_new >= mem.dst(nspp)) {
                            do { ++nspp; } while (s_new >= mem.dst(nspp));
                            if (nspp == SPP) break;
                        }
                    }
                    t = f_add(t, dt);
                }
                longest = std::max(longest, ray[l].size());
                steps += (double)ray[l].size();
            }
            for (size_t k = 0; k < longest; ++k)
                for (int L = 0; L < kNL; ++L) {
                    const Layout& lay = kLayouts[L];
                    tl.clear(); ts.clear(); bl.clear(); bs.clear();
                    for (int l = 0; l < lanes; ++l) {
                        if (k >= ray[l].size()) continue;
                        const Step& s = ray[l][k];
                        const uint64_t tidx = lay.table_morton ? morton3(s.x >> 3, s.y >> 3, s.z >> 3)
                                                               : ((((uint64_t)(s.x >> 3) << K) | (s.y >> 3)) << K | (s.z >> 3));
                        tl.push_back(tidx * 4 / 128);
                        ts.push_back(tidx * 4 / 32);
                        if (s.brick != 0xffffffffu) {
                            const uint32_t lx = s.x & 7, ly = s.y & 7, lz = s.z & 7;
                            const uint64_t c = lay.cell_order ? morton3(lx, ly, lz) : ((lx << 6) | (ly << 3) | lz);
                            const uint64_t bit = ((uint64_t)s.brick * 512 + c) * (uint64_t)lay.cell_bits;
                            bl.push_back(bit / (128 * 8));
                            bs.push_back(bit / (32 * 8));
                        }
                    }
                    double* o = out + L * 6;
                    if (!tl.empty()) { o[0] += 1; o[1] += (double)count_distinct(tl); o[2] += (double)count_distinct(ts); }
                    if (!bl.empty()) { o[3] += 1; o[4] += (double)count_distinct(bl); o[5] += (double)count_distinct(bs); }
                }
        }
    *steps_total = steps;
    return K;
}

extern "C" const char* coherence_layout_name(int i) { return i >= 0 && i < kNL ? kLayouts[i].name : ""; }
