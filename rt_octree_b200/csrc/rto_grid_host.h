// rto_grid_host.h — host-side builder of the sparse brick grid (plain C++, shared by rto_api.cu and the CPU test harness).
#pragma once
#include <cstdint>
#include <vector>

#include "rto_ray.cuh"

namespace rto {

// Sparse brick grid (rto_ray.cuh GridDev), built on the host from the original tree.npz arrays at load time.
// Returns false (no grid) when the depth is outside the supported range.
inline bool build_grid_host(const int32_t* child, const uint16_t* data, int data_dim, int64_t capacity, int max_depth,
                     std::vector<uint32_t>& top, std::vector<uint32_t>& bricks, int& K) {
    (void)capacity;
    const int D = max_depth;
    if (D < 4 || D > 11) return false;   // K = D-3 in [1, 8]
    K = D - 3;
    const size_t S = (size_t)1 << K;
    top.assign(S * S * S, 0u);
    bricks.clear();
    struct Item { int64_t node; int l; uint32_t x, y, z; int64_t brick; };
    std::vector<Item> stack;
    stack.push_back({0, 0, 0u, 0u, 0u, -1});
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        for (int o = 0; o < 8; ++o) {
            const uint32_t cx = it.x * 2 + ((o >> 2) & 1), cy = it.y * 2 + ((o >> 1) & 1), cz = it.z * 2 + (o & 1);
            const int d = it.l + 1;   // depth (look-ups) of a leaf child / level of an internal child
            const int64_t e = it.node * 8 + o;
            const int32_t skip = child[e];
            if (skip == 0) {
                const uint32_t word = RTO_LEAF_FLAG | ((uint32_t)(127 + d) << 23) | data[e * data_dim + data_dim - 1];
                if (d <= K) {
                    const uint32_t s = 1u << (K - d);
                    for (uint32_t a = 0; a < s; ++a)
                        for (uint32_t b = 0; b < s; ++b)
                            for (uint32_t c = 0; c < s; ++c)
                                top[(((size_t)(cx * s + a) << K) | (cy * s + b)) << K | (cz * s + c)] = word;
                } else {
                    const uint32_t s = 1u << (D - d);
                    const uint32_t lx = (cx * s) & 7u, ly = (cy * s) & 7u, lz = (cz * s) & 7u;
                    uint32_t* br = bricks.data() + (size_t)it.brick * 512;
                    for (uint32_t a = 0; a < s; ++a)
                        for (uint32_t b = 0; b < s; ++b)
                            for (uint32_t c = 0; c < s; ++c) br[brick_cell_index(lx + a, ly + b, lz + c)] = word;
                }
            } else {
                int64_t brick = it.brick;
                if (d == K) {
                    brick = (int64_t)(bricks.size() / 512);
                    bricks.resize(bricks.size() + 512, 0u);
                    top[(((size_t)cx << K) | cy) << K | cz] = (uint32_t)brick;
                }
                stack.push_back({it.node + skip, d, cx, cy, cz, brick});
            }
        }
    }
    return true;
}

// leaf-id planes (rto_ray.cuh GridDev::leaf_top / leaf_bricks) from the tree itself: every level-K cell and every finest-level
// cell of a brick is located by a plain descent from the root, independently of how `top` / `bricks` were filled.
inline void build_grid_leaf_host(const int32_t* child, int K, const std::vector<uint32_t>& top, size_t n_bricks,
                                 std::vector<uint32_t>& leaf_top, std::vector<uint32_t>& leaf_bricks) {
    const size_t S = (size_t)1 << K;
    leaf_top.assign(S * S * S, 0u);
    leaf_bricks.assign((n_bricks ? n_bricks : 1) * 512, 0u);
    for (uint32_t x = 0; x < S; ++x)
        for (uint32_t y = 0; y < S; ++y)
            for (uint32_t z = 0; z < S; ++z) {
                const size_t t = (((size_t)x << K) | y) << K | z;
                int64_t node = 0;
                bool leaf = false;
                for (int d = 1; d <= K && !leaf; ++d) {
                    const int sh = K - d;
                    const int64_t e = node * 8 + ((((x >> sh) & 1u) << 2) | (((y >> sh) & 1u) << 1) | ((z >> sh) & 1u));
                    if (child[e] == 0) { leaf_top[t] = (uint32_t)e; leaf = true; }
                    else node += child[e];
                }
                if (leaf) continue;
                uint32_t* out = leaf_bricks.data() + (size_t)top[t] * 512;
                for (uint32_t c = 0; c < 512; ++c) {
                    const uint32_t lx = c >> 6, ly = (c >> 3) & 7u, lz = c & 7u;
                    int64_t n2 = node;
                    for (int j = 1; j <= 3; ++j) {
                        const int sh = 3 - j;
                        const int64_t e = n2 * 8 + ((((lx >> sh) & 1u) << 2) | (((ly >> sh) & 1u) << 1) | ((lz >> sh) & 1u));
                        if (child[e] == 0) { out[c] = (uint32_t)e; break; }
                        n2 += child[e];
                    }
                }
            }
}

// march table of the fused-index marcher (rto_ray.cuh FusedIdx / march_top_entry) from the finished top table
inline void build_march_top_host(const std::vector<uint32_t>& top, int K, std::vector<uint32_t>& top_m) {
    const uint32_t S = 1u << K;
    top_m.resize(top.size());
    for (uint32_t x = 0; x < S; ++x)
        for (uint32_t y = 0; y < S; ++y)
            for (uint32_t z = 0; z < S; ++z) {
                const size_t t = (((size_t)x << K) | y) << K | z;
                top_m[t] = march_top_entry(top[t], x, y, z);
            }
}

// byte plane of the bricks (rto_ray.cuh brick_byte): depth | 0x80 where sigma is non-zero
inline void grid_bytes_host(const std::vector<uint32_t>& bricks, std::vector<uint8_t>& bricks8) {
    bricks8.resize(bricks.size());
    for (size_t i = 0; i < bricks.size(); ++i) bricks8[i] = brick_byte(bricks[i]);
}

}  // namespace rto
