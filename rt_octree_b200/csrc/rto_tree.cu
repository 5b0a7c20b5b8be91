// rto_tree.cu — device octree loader: upload tree.npz arrays and re-lay them out on the GPU as SoA.
//
// Replaces N3Tree::load_cuda (renderer/src/cuda/n3tree.cu:9-41), which uploads the npz arrays verbatim (AoS:
// `data` fp16 [cap][8][data_dim] with sigma last, `child` int32 relative offsets).  HBM layout here:
//   nodes   u32 [cap*8]            internal: ABSOLUTE child node id ; leaf: 0x80000000 | sigma fp16 bits
//   payload fp16 [cap*8][stride]   the data_dim-1 colour coefficients of each entry, zero padded to 16 B multiples
// so a traversal step reads one 4-byte word from a 32-byte node record and never touches colour data; colour is
// read (64 B aligned for SH9) only for the <= SPP collided leaves of a ray.  Node numbering is unchanged, so the flat
// leaf index node*8+octant is the reference's `sub_ptr` — the identity the bit-exact trace is compared on.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "rto_grid_host.h"
#include "rto_internal.h"

namespace rto {

__global__ void build_nodes_kernel(const int32_t* __restrict__ child, const __half* __restrict__ data, int data_dim,
                                   int64_t n_entries, int64_t capacity, uint32_t* __restrict__ nodes,
                                   unsigned long long* __restrict__ n_leaves, int* __restrict__ bad) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned leaf = 0;
    if (e < n_entries) {
        const int32_t skip = child[e];
        if (skip == 0) {
            const uint16_t sig = __half_as_ushort(data[e * data_dim + data_dim - 1]);
            nodes[e] = RTO_LEAF_FLAG | sig;
            leaf = 1;
        } else {
            const int64_t tgt = (e >> 3) + skip;   // ptr += skip * N3 (n3tree_query.hpp:46), in nodes
            if (tgt <= 0 || tgt >= capacity) { *bad = 1; nodes[e] = RTO_LEAF_FLAG; }
            else nodes[e] = (uint32_t)tgt;
        }
    }
    const unsigned cnt = __popc(__ballot_sync(0xffffffffu, leaf));
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_leaves, (unsigned long long)cnt);
}

// one thread per (entry, 8-half chunk): 16-byte stores, coalesced along the payload
__global__ void build_payload_kernel(const __half* __restrict__ data, int data_dim, int stride, int64_t n_entries,
                                     __half* __restrict__ payload) {
    const int chunks = stride / 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_entries * chunks) return;
    const int64_t e = i / chunks;
    const int c = (int)(i - e * chunks);
    const __half* src = data + e * data_dim;
    __align__(16) __half v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int j = c * 8 + k;
        v[k] = j < data_dim - 1 ? src[j] : __float2half(0.f);
    }
    *reinterpret_cast<uint4*>(payload + e * stride + c * 8) = *reinterpret_cast<const uint4*>(v);
}

cudaError_t launch_build_nodes(const int32_t* child, const __half* data, int data_dim, int64_t n_entries,
                               int64_t capacity, uint32_t* nodes, unsigned long long* n_leaves, int* bad,
                               cudaStream_t s) {
    const int B = 256;
    build_nodes_kernel<<<(unsigned)((n_entries + B - 1) / B), B, 0, s>>>(child, data, data_dim, n_entries, capacity,
                                                                         nodes, n_leaves, bad);
    return cudaGetLastError();
}
cudaError_t launch_build_payload(const __half* data, int data_dim, int stride, int64_t n_entries, __half* payload,
                                 cudaStream_t s) {
    const int B = 256;
    const int64_t total = n_entries * (stride / 8);
    build_payload_kernel<<<(unsigned)((total + B - 1) / B), B, 0, s>>>(data, data_dim, stride, n_entries, payload);
    return cudaGetLastError();
}

// Host-side structural check + depth: breadth-first walk from the root over the ORIGINAL child array.
// Returns max look-ups to a leaf, or -1 if a pointer leaves the array or the walk visits more than `capacity`
// nodes (cycle / shared subtree blow-up).
int tree_max_depth_host(const int32_t* child, int64_t capacity) {
    std::vector<int64_t> cur{0}, nxt;
    int depth = 0;
    int64_t visited = 0;
    while (!cur.empty()) {
        ++depth;
        nxt.clear();
        for (int64_t n : cur) {
            if (++visited > capacity) return -1;
            for (int i = 0; i < 8; ++i) {
                const int32_t skip = child[n * 8 + i];
                if (skip != 0) {
                    const int64_t t = n + skip;
                    if (t <= 0 || t >= capacity) return -1;
                    nxt.push_back(t);
                }
            }
        }
        cur.swap(nxt);
    }
    return depth;
}


}  // namespace rto
