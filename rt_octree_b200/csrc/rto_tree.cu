// rto_tree.cu — device octree loader: upload the tree.npz arrays and build every HBM plane ON THE GPU.
//
// Replaces N3Tree::load_cuda (renderer/src/cuda/n3tree.cu:9-41), which uploads the npz arrays verbatim (AoS:
// `data` fp16 [cap][8][data_dim] with sigma last, `child` int32 relative offsets), and the host-side codebook decode of
// quantised files (renderer/src/n3tree.cpp:279-340).  HBM layout here:
//   nodes   u32 [cap*8]            internal: ABSOLUTE child node id ; leaf: 0x80000000 | sigma fp16 bits
//   payload fp16 [cap*8][stride]   the data_dim-1 colour coefficients of each entry, zero padded to 16 B multiples
//   grid    level-K table + 8^3 bricks of leaf words (rto_ray.cuh GridDev) for the marching loop
// so a traversal step reads one 4-byte word and never touches colour data; colour is read (64 B aligned for SH9) only
// for the <= SPP collided leaves of a ray.  Node numbering is unchanged, so the flat leaf index node*8+octant is the
// reference's `sub_ptr` — the identity the bit-exact trace is compared on.
//
// Pipeline (all kernels below, default stream, load time only):
//   build_nodes_kernel            child + sigma           -> nodes (range check of every offset)
//   build_payload_kernel          dense data              -> payload            } one of the two
//   build_payload_quant_kernel    codebook + map (+kept)  -> payload            }
//   bfs_level_kernel x depth      nodes                   -> max depth, cycle check (visited > capacity)
//   grid_top_kernel               nodes                   -> top table leaf words / level-K node ids
//   grid_scan_kernel              per-block brick counts  -> exclusive offsets
//   grid_assign_kernel            ids in the host builder's order -> top table brick ids, brick -> node map
//   grid_brick_kernel             nodes                   -> bricks (leaf words) + byte bricks (depth | dense flag)
//   grid_leaf_top_kernel          nodes + top             -> leaf-id plane of the level-K cells, level-K node of every brick
//   grid_leaf_brick_kernel        nodes                   -> leaf-id plane of the brick cells
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/rtoctree_b200.h"
#include "rto_grid_host.h"
#include "rto_internal.h"
#include "rto_tree.h"

namespace rto {

// sigma of entry e = sig[e * sig_stride + sig_off]: (data_dim, data_dim-1) for the dense `data` array, (1, 0) for the
// separate `sigma` array of a quantised file
__global__ void build_nodes_kernel(const int32_t* __restrict__ child, const __half* __restrict__ sig, int sig_stride,
                                   int sig_off, int64_t n_entries, int64_t capacity, uint32_t* __restrict__ nodes,
                                   unsigned long long* __restrict__ n_leaves, int* __restrict__ bad) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned leaf = 0;
    if (e < n_entries) {
        const int32_t skip = child[e];
        if (skip == 0) {
            nodes[e] = RTO_LEAF_FLAG | __half_as_ushort(sig[e * sig_stride + sig_off]);
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

// Quantised file (scripts/compress_octree.py:68-119, decode n3tree.cpp:279-340): payload slot p = k*n_basis + b holds
//   b <  n_ret : data_retained[b][e][k]
//   b >= n_ret : quant_colors[b-n_ret][quant_map[b-n_ret][e]][k]
// gathered straight into the padded payload plane (the AoS `data` array is never materialised).
__global__ void build_payload_quant_kernel(const __half* __restrict__ colors, const uint16_t* __restrict__ map,
                                           const __half* __restrict__ retained, int n_q, int n_ret, int data_dim,
                                           int stride, int64_t n_entries, __half* __restrict__ payload) {
    const int chunks = stride / 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_entries * chunks) return;
    const int64_t e = i / chunks;
    const int c = (int)(i - e * chunks);
    const int n_basis = n_q + n_ret;
    __align__(16) __half v[8];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const int p = c * 8 + s;
        __half x = __float2half(0.f);
        const int k = p / n_basis, b = p - k * n_basis;
        if (p < data_dim - 1 && k < 3) {   // slots past 3*n_basis (none in svox files) stay zero like the reference's
            if (b < n_ret) {
                x = retained[((int64_t)b * n_entries + e) * 3 + k];
            } else {
                const int j = b - n_ret;
                const uint32_t id = map[(int64_t)j * n_entries + e];
                x = colors[((int64_t)j * 65536 + id) * 3 + k];
            }
        }
        v[s] = x;
    }
    *reinterpret_cast<uint4*>(payload + e * stride + c * 8) = *reinterpret_cast<const uint4*>(v);
}

// ------------------------------------------------------------------------------------------ depth / structure check
// One breadth-first level: thread = (frontier node, octant); internal entries append their child to the next frontier.
// The host adds the level sizes up: more than `capacity` visited nodes means the links form a cycle or share subtrees
// (the same criterion as the former host walk).  Order inside a frontier is irrelevant.
__global__ void bfs_level_kernel(const uint32_t* __restrict__ nodes, const uint32_t* __restrict__ cur, uint32_t n_cur,
                                 uint32_t* __restrict__ nxt, uint32_t cap_nxt, unsigned* __restrict__ n_nxt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t w = RTO_LEAF_FLAG;
    if (i < (int64_t)n_cur * 8) w = nodes[(int64_t)cur[i >> 3] * 8 + (i & 7)];
    const bool internal = !(w & RTO_LEAF_FLAG);
    const unsigned m = __ballot_sync(0xffffffffu, internal);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(n_nxt, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (internal) {
        const unsigned pos = base + __popc(m & ((1u << lane) - 1u));
        if (pos < cap_nxt) nxt[pos] = w;
    }
}

// --------------------------------------------------------------------------------------------- sparse brick grid
// Thread i of the top pass handles ONE level-K cell.  Cells are enumerated in the order in which the host builder
// (rto_grid_host.h: explicit stack, children pushed 0..7 and popped last-first) reaches the level-K nodes: Morton
// digits of levels 1..K-1 complemented, the level-K digit as is.  Brick ids handed out by a prefix sum over i are
// then IDENTICAL to the host builder's, so both builders produce the same tables bit for bit.
__device__ __forceinline__ void grid_cell_of(uint32_t i, int K, uint32_t& x, uint32_t& y, uint32_t& z) {
    const uint32_t m = i ^ (((1u << (3 * (K - 1))) - 1u) << 3);   // un-complement the upper digits
    x = y = z = 0;
    for (int l = 0; l < K; ++l) {   // digit l (0 = finest level of the table)
        const uint32_t o = (m >> (3 * l)) & 7u;
        x |= ((o >> 2) & 1u) << l;
        y |= ((o >> 1) & 1u) << l;
        z |= (o & 1u) << l;
    }
}

__global__ void grid_top_kernel(const uint32_t* __restrict__ nodes, int K, uint32_t n_cells, uint32_t* __restrict__ top,
                                uint32_t* __restrict__ cell_node, uint32_t* __restrict__ block_count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t node_at_K = 0u;   // 0 = the cell is covered by a leaf (node 0 is the root, never a level-K child)
    if (i < n_cells) {
        uint32_t x, y, z;
        grid_cell_of(i, K, x, y, z);
        uint32_t node = 0u;
        uint32_t word = 0u;
        int d = 0;
        for (d = 1; d <= K; ++d) {
            const int sh = K - d;
            const uint32_t oct = (((x >> sh) & 1u) << 2) | (((y >> sh) & 1u) << 1) | ((z >> sh) & 1u);
            word = nodes[node * 8u + oct];
            if (word & RTO_LEAF_FLAG) break;
            node = word;
        }
        const size_t t = (((size_t)x << K) | y) << K | z;
        if (d <= K) top[t] = RTO_LEAF_FLAG | ((uint32_t)(127 + d) << 23) | (word & 0xffffu);
        else node_at_K = node;
        cell_node[i] = node_at_K;
    }
    const int cnt = __syncthreads_count(node_at_K != 0u);
    if (threadIdx.x == 0) block_count[blockIdx.x] = (uint32_t)cnt;
}

// exclusive prefix sum of the per-block counts, one 1024-thread block (at most 2^24 / 256 = 65536 values)
__global__ void grid_scan_kernel(uint32_t* __restrict__ block_count, uint32_t n_blocks, uint32_t* __restrict__ total) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_blocks; base += blockDim.x) {
        const uint32_t idx = base + threadIdx.x;
        const uint32_t v = idx < n_blocks ? block_count[idx] : 0u;
        uint32_t inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) warp_sum[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            uint32_t s = warp_sum[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += n;
            }
            warp_sum[lane] = s;   // inclusive over warps
        }
        __syncthreads();
        const uint32_t before = carry + (wid ? warp_sum[wid - 1] : 0u) + inc - v;
        if (idx < n_blocks) block_count[idx] = before;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void grid_assign_kernel(const uint32_t* __restrict__ cell_node, const uint32_t* __restrict__ block_offset,
                                   int K, uint32_t n_cells, uint32_t* __restrict__ top, uint32_t* __restrict__ brick_node) {
    __shared__ uint32_t warp_cnt[8];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t node = i < n_cells ? cell_node[i] : 0u;
    const unsigned m = __ballot_sync(0xffffffffu, node != 0u);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[wid] = __popc(m);
    __syncthreads();
    if (node == 0u) return;
    uint32_t id = block_offset[blockIdx.x] + __popc(m & ((1u << lane) - 1u));
    for (int w = 0; w < wid; ++w) id += warp_cnt[w];
    uint32_t x, y, z;
    grid_cell_of(i, K, x, y, z);
    top[(((size_t)x << K) | y) << K | z] = id;
    brick_node[id] = node;
}

// one 512-thread block per brick, one thread per finest-level cell: at most 3 more look-ups below the level-K node
__global__ void grid_brick_kernel(const uint32_t* __restrict__ nodes, const uint32_t* __restrict__ brick_node, int K,
                                  uint32_t* __restrict__ bricks, uint8_t* __restrict__ bricks8) {
    const uint32_t b = blockIdx.x;
    const uint32_t c = threadIdx.x;   // == brick_cell_index(lx, ly, lz)
    const uint32_t lx = c >> 6, ly = (c >> 3) & 7u, lz = c & 7u;
    uint32_t node = brick_node[b];
    uint32_t out = 0u;   // stays 0 only if the tree were deeper than max_depth (excluded by the depth pass)
    for (int j = 1; j <= 3; ++j) {
        const int sh = 3 - j;
        const uint32_t oct = (((lx >> sh) & 1u) << 2) | (((ly >> sh) & 1u) << 1) | ((lz >> sh) & 1u);
        const uint32_t word = nodes[node * 8u + oct];
        if (word & RTO_LEAF_FLAG) {
            out = RTO_LEAF_FLAG | ((uint32_t)(127 + K + j) << 23) | (word & 0xffffu);
            break;
        }
        node = word;
    }
    bricks[(size_t)b * 512 + c] = out;
    bricks8[(size_t)b * 512 + c] = brick_byte(out);
}

// Leaf-id planes (rto_ray.cuh GridDev::leaf_top / leaf_bricks), derived from the node words and the finished top table
// only — whichever builder filled the grid.  Thread t = level-K cell (x-major index, as the marcher forms it): a leaf
// reached within K look-ups gives leaf_top[t]; otherwise top[t] is the cell's brick id and the level-K node is noted for
// the brick pass, which descends at most 3 more levels per finest-level cell.
__global__ void grid_leaf_top_kernel(const uint32_t* __restrict__ nodes, const uint32_t* __restrict__ top, int K,
                                     uint32_t n_cells, uint32_t* __restrict__ leaf_top, uint32_t* __restrict__ brick_node) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cells) return;
    const uint32_t mask = (1u << K) - 1u;
    const uint32_t x = t >> (2 * K), y = (t >> K) & mask, z = t & mask;
    uint32_t node = 0u;
    for (int d = 1; d <= K; ++d) {
        const int sh = K - d;
        const uint32_t e = node * 8u + ((((x >> sh) & 1u) << 2) | (((y >> sh) & 1u) << 1) | ((z >> sh) & 1u));
        const uint32_t word = nodes[e];
        if (word & RTO_LEAF_FLAG) { leaf_top[t] = e; return; }
        node = word;
    }
    leaf_top[t] = 0u;
    brick_node[top[t]] = node;
}
__global__ void grid_leaf_brick_kernel(const uint32_t* __restrict__ nodes, const uint32_t* __restrict__ brick_node,
                                       uint32_t* __restrict__ leaf_bricks) {
    const uint32_t c = threadIdx.x;   // == brick_cell_index(lx, ly, lz)
    const uint32_t lx = c >> 6, ly = (c >> 3) & 7u, lz = c & 7u;
    uint32_t node = brick_node[blockIdx.x];
    uint32_t out = 0u;
    for (int j = 1; j <= 3; ++j) {
        const int sh = 3 - j;
        const uint32_t e = node * 8u + ((((lx >> sh) & 1u) << 2) | (((ly >> sh) & 1u) << 1) | ((lz >> sh) & 1u));
        const uint32_t word = nodes[e];
        if (word & RTO_LEAF_FLAG) { out = e; break; }
        node = word;
    }
    leaf_bricks[(size_t)blockIdx.x * 512 + c] = out;
}

// March table of the fused-index marcher (rto_ray.cuh FusedIdx): leaf words as they are, brick ids as biased cell offsets.
__global__ void grid_march_top_kernel(const uint32_t* __restrict__ top, int K, uint32_t n_cells, uint32_t* __restrict__ top_m) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cells) return;
    const uint32_t mask = (1u << K) - 1u;
    top_m[t] = march_top_entry(top[t], t >> (2 * K), (t >> K) & mask, t & mask);
}

// --------------------------------------------------------------------------------------------------- host driver
namespace {
struct DevBuf {   // frees on scope exit unless released
    void* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    template <class T> T* as() const { return static_cast<T*>(p); }
    template <class T> T* release() { T* r = static_cast<T*>(p); p = nullptr; return r; }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
};

struct Fail {
    std::string& err;
    int operator()(int code, const std::string& msg) const { err = msg; return code; }
    int cuda(cudaError_t e, const char* what) const {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? RTO_ERR_NOMEM : RTO_ERR_CUDA;
    }
};
#define RTO_TRY(expr, what) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail.cuda(e_, what); } while (0)

inline unsigned blocks_for(int64_t n, int B) { return (unsigned)((n + B - 1) / B); }
}  // namespace

// max look-ups to a leaf; -1: cycle / shared subtrees.  nodes already range-checked by build_nodes_kernel.
static int tree_depth_device(const uint32_t* nodes, int64_t capacity, int* depth_out, int64_t* launches, const Fail& fail) {
    DevBuf fa, fb, cnt;
    RTO_TRY(fa.alloc((size_t)capacity * sizeof(uint32_t)), "cudaMalloc(frontier)");
    RTO_TRY(fb.alloc((size_t)capacity * sizeof(uint32_t)), "cudaMalloc(frontier)");
    RTO_TRY(cnt.alloc(sizeof(unsigned)), "cudaMalloc");
    RTO_TRY(cudaMemset(fa.p, 0, sizeof(uint32_t)), "cudaMemset");   // frontier 0 = {root}
    uint32_t* cur = fa.as<uint32_t>();
    uint32_t* nxt = fb.as<uint32_t>();
    unsigned n_cur = 1;
    int64_t visited = 1;
    int depth = 0;
    while (n_cur) {
        ++depth;
        if (depth > RTO_COORD_BITS + 1) break;   // deeper than any supported tree: the caller rejects it
        RTO_TRY(cudaMemset(cnt.p, 0, sizeof(unsigned)), "cudaMemset");
        bfs_level_kernel<<<blocks_for((int64_t)n_cur * 8, 256), 256>>>(nodes, cur, n_cur, nxt, (uint32_t)capacity, cnt.as<unsigned>());
        RTO_TRY(cudaGetLastError(), "bfs_level_kernel");
        ++*launches;
        unsigned n_nxt = 0;
        RTO_TRY(cudaMemcpy(&n_nxt, cnt.p, sizeof n_nxt, cudaMemcpyDeviceToHost), "D2H");
        visited += n_nxt;
        if (visited > capacity) { *depth_out = -1; return RTO_OK; }
        uint32_t* t = cur; cur = nxt; nxt = t;
        n_cur = n_nxt;
    }
    *depth_out = depth;
    return RTO_OK;
}

static int grid_build_device(const uint32_t* nodes, int max_depth, TreeBuilt& out, int64_t* launches, const Fail& fail) {
    if (max_depth < 4 || max_depth > 11) return RTO_OK;   // K = D-3 in [1, 8]; other depths march through the tree
    const int K = max_depth - 3;
    const uint32_t n_cells = 1u << (3 * K);
    const int B = 256;
    const unsigned nb = blocks_for(n_cells, B);
    DevBuf top, cell_node, block_count, total, brick_node, bricks, bricks8;
    RTO_TRY(top.alloc((size_t)n_cells * sizeof(uint32_t)), "cudaMalloc(grid top)");
    RTO_TRY(cell_node.alloc((size_t)n_cells * sizeof(uint32_t)), "cudaMalloc(grid scratch)");
    RTO_TRY(block_count.alloc((size_t)nb * sizeof(uint32_t)), "cudaMalloc(grid scratch)");
    RTO_TRY(total.alloc(sizeof(uint32_t)), "cudaMalloc");
    grid_top_kernel<<<nb, B>>>(nodes, K, n_cells, top.as<uint32_t>(), cell_node.as<uint32_t>(), block_count.as<uint32_t>());
    RTO_TRY(cudaGetLastError(), "grid_top_kernel");
    grid_scan_kernel<<<1, 1024>>>(block_count.as<uint32_t>(), nb, total.as<uint32_t>());
    RTO_TRY(cudaGetLastError(), "grid_scan_kernel");
    *launches += 2;
    uint32_t n_bricks = 0;
    RTO_TRY(cudaMemcpy(&n_bricks, total.p, sizeof n_bricks, cudaMemcpyDeviceToHost), "D2H");
    if ((size_t)n_bricks >= ((size_t)1 << 23)) return RTO_OK;   // the marcher indexes brick words with 32 bits (grid_lookup)
    RTO_TRY(brick_node.alloc((size_t)n_bricks * sizeof(uint32_t)), "cudaMalloc(grid scratch)");
    RTO_TRY(bricks.alloc((size_t)(n_bricks ? n_bricks : 1) * 512 * sizeof(uint32_t)), "cudaMalloc(grid bricks)");
    RTO_TRY(bricks8.alloc((size_t)(n_bricks ? n_bricks : 1) * 512), "cudaMalloc(grid byte bricks)");
    if (n_bricks) {
        grid_assign_kernel<<<nb, B>>>(cell_node.as<uint32_t>(), block_count.as<uint32_t>(), K, n_cells, top.as<uint32_t>(),
                                      brick_node.as<uint32_t>());
        RTO_TRY(cudaGetLastError(), "grid_assign_kernel");
        grid_brick_kernel<<<n_bricks, 512>>>(nodes, brick_node.as<uint32_t>(), K, bricks.as<uint32_t>(), bricks8.as<uint8_t>());
        RTO_TRY(cudaGetLastError(), "grid_brick_kernel");
        *launches += 2;
    }
    RTO_TRY(cudaDeviceSynchronize(), "grid build");
    out.grid_top = top.release<uint32_t>();
    out.grid_bricks = bricks.release<uint32_t>();
    out.grid_bricks8 = bricks8.release<uint8_t>();
    out.grid_K = K;
    out.n_bricks = n_bricks;
    return RTO_OK;
}

// The former host-side builder (rto_grid_host.h), kept selectable with RTO_GRID_BUILD=host as the cross-check of the
// device builder (tests/test_gpu_tree.py compares the two tables bit for bit).  Needs the dense host arrays.
static int grid_build_host(const TreeSource& s, int max_depth, TreeBuilt& out, const Fail& fail) {
    std::vector<uint32_t> top, bricks;
    int K = 0;
    if (!build_grid_host(s.child, static_cast<const uint16_t*>(s.data_f16), s.data_dim, s.capacity, max_depth, top, bricks, K) ||
        bricks.size() / 512 >= ((size_t)1 << 23))
        return RTO_OK;
    std::vector<uint8_t> bricks8;
    grid_bytes_host(bricks, bricks8);
    DevBuf dt, db, d8;
    RTO_TRY(dt.alloc(top.size() * sizeof(uint32_t)), "cudaMalloc(grid top)");
    RTO_TRY(db.alloc((bricks.empty() ? 512 : bricks.size()) * sizeof(uint32_t)), "cudaMalloc(grid bricks)");
    RTO_TRY(d8.alloc(bricks.empty() ? 512 : bricks.size()), "cudaMalloc(grid byte bricks)");
    RTO_TRY(cudaMemcpy(dt.p, top.data(), top.size() * sizeof(uint32_t), cudaMemcpyHostToDevice), "H2D grid top");
    if (!bricks.empty()) {
        RTO_TRY(cudaMemcpy(db.p, bricks.data(), bricks.size() * sizeof(uint32_t), cudaMemcpyHostToDevice), "H2D grid bricks");
        RTO_TRY(cudaMemcpy(d8.p, bricks8.data(), bricks8.size(), cudaMemcpyHostToDevice), "H2D grid byte bricks");
    }
    out.grid_top = dt.release<uint32_t>();
    out.grid_bricks = db.release<uint32_t>();
    out.grid_bricks8 = d8.release<uint8_t>();
    out.grid_K = K;
    out.n_bricks = (int64_t)(bricks.size() / 512);
    return RTO_OK;
}

// leaf-id planes for a finished grid (either builder).  Skipped (planes stay nullptr, collisions then descend the tree) when
// the brick-cell reference would not fit 31 bits, or with RTO_LEAF_PLANES=0.
static int grid_leaf_build_device(const uint32_t* nodes, TreeBuilt& out, int64_t* launches, const Fail& fail) {
    const char* off = getenv("RTO_LEAF_PLANES");
    if (out.grid_K <= 0 || out.n_bricks >= ((int64_t)1 << 22) || (off && off[0] == '0')) return RTO_OK;
    const uint32_t n_cells = 1u << (3 * out.grid_K);
    const size_t nb = (size_t)(out.n_bricks ? out.n_bricks : 1);
    DevBuf lt, lb, bn;
    RTO_TRY(lt.alloc((size_t)n_cells * sizeof(uint32_t)), "cudaMalloc(grid leaf top)");
    RTO_TRY(lb.alloc(nb * 512 * sizeof(uint32_t)), "cudaMalloc(grid leaf bricks)");
    RTO_TRY(bn.alloc(nb * sizeof(uint32_t)), "cudaMalloc(grid scratch)");
    grid_leaf_top_kernel<<<blocks_for(n_cells, 256), 256>>>(nodes, out.grid_top, out.grid_K, n_cells, lt.as<uint32_t>(), bn.as<uint32_t>());
    RTO_TRY(cudaGetLastError(), "grid_leaf_top_kernel");
    ++*launches;
    if (out.n_bricks) {
        grid_leaf_brick_kernel<<<(unsigned)out.n_bricks, 512>>>(nodes, bn.as<uint32_t>(), lb.as<uint32_t>());
        RTO_TRY(cudaGetLastError(), "grid_leaf_brick_kernel");
        ++*launches;
    }
    // march table (fused-index marcher): needs the leaf-id planes (its collisions are cell references) and brick
    // offsets below 2^31 - bias; RTO_FUSED_INDEX=0 skips it (the v9 loop then runs)
    DevBuf tm;
    const char* fo = getenv("RTO_FUSED_INDEX");
    const bool fused = out.n_bricks <= RTO_FUSED_MAX_BRICKS && !(fo && fo[0] == '0');
    if (fused) {
        RTO_TRY(tm.alloc((size_t)n_cells * sizeof(uint32_t)), "cudaMalloc(grid march table)");
        grid_march_top_kernel<<<blocks_for(n_cells, 256), 256>>>(out.grid_top, out.grid_K, n_cells, tm.as<uint32_t>());
        RTO_TRY(cudaGetLastError(), "grid_march_top_kernel");
        ++*launches;
    }
    RTO_TRY(cudaDeviceSynchronize(), "grid leaf planes");
    out.grid_leaf_top = lt.release<uint32_t>();
    out.grid_leaf_bricks = lb.release<uint32_t>();
    if (fused) out.grid_top_m = tm.release<uint32_t>();
    return RTO_OK;
}

void tree_built_free(TreeBuilt& b) {
    cudaFree(b.grid_leaf_top); cudaFree(b.grid_leaf_bricks); cudaFree(b.grid_top_m);
    cudaFree(b.nodes); cudaFree(b.payload); cudaFree(b.grid_top); cudaFree(b.grid_bricks); cudaFree(b.grid_bricks8);
    b = TreeBuilt{};
}

int build_tree_device(const TreeSource& s, TreeBuilt& out, std::string& err, int64_t* launches) {
    const Fail fail{err};
    out = TreeBuilt{};
    const int64_t n_entries = s.capacity * 8;
    const int stride = ((s.data_dim - 1) + 7) / 8 * 8;
    const bool quant = s.data_f16 == nullptr;
    const int B = 256;
    DevBuf nodes, payload;
    RTO_TRY(nodes.alloc(n_entries * sizeof(uint32_t)), "cudaMalloc(nodes)");
    RTO_TRY(payload.alloc((size_t)n_entries * stride * sizeof(__half)), "cudaMalloc(payload)");
    {   // staging buffers of the source arrays live only inside this block
        DevBuf d_child, d_cnt, d_bad, d_data, d_colors, d_map, d_sigma, d_ret;
        RTO_TRY(d_child.alloc(n_entries * sizeof(int32_t)), "cudaMalloc(child)");
        RTO_TRY(d_cnt.alloc(sizeof(unsigned long long)), "cudaMalloc");
        RTO_TRY(d_bad.alloc(sizeof(int)), "cudaMalloc");
        RTO_TRY(cudaMemset(d_cnt.p, 0, sizeof(unsigned long long)), "cudaMemset");
        RTO_TRY(cudaMemset(d_bad.p, 0, sizeof(int)), "cudaMemset");
        RTO_TRY(cudaMemcpy(d_child.p, s.child, n_entries * sizeof(int32_t), cudaMemcpyHostToDevice), "H2D child");
        const int64_t chunks_total = n_entries * (stride / 8);
        if (!quant) {
            const size_t bytes = (size_t)n_entries * s.data_dim * sizeof(__half);
            RTO_TRY(d_data.alloc(bytes), "cudaMalloc(data)");
            RTO_TRY(cudaMemcpy(d_data.p, s.data_f16, bytes, cudaMemcpyHostToDevice), "H2D data");
            build_nodes_kernel<<<blocks_for(n_entries, B), B>>>(d_child.as<int32_t>(), d_data.as<__half>(), s.data_dim,
                                                                s.data_dim - 1, n_entries, s.capacity, nodes.as<uint32_t>(),
                                                                d_cnt.as<unsigned long long>(), d_bad.as<int>());
            RTO_TRY(cudaGetLastError(), "build_nodes");
            build_payload_kernel<<<blocks_for(chunks_total, B), B>>>(d_data.as<__half>(), s.data_dim, stride, n_entries,
                                                                     payload.as<__half>());
            RTO_TRY(cudaGetLastError(), "build_payload");
        } else {
            const size_t col_bytes = (size_t)s.n_q * 65536 * 3 * sizeof(__half);
            const size_t map_bytes = (size_t)s.n_q * n_entries * sizeof(uint16_t);
            const size_t ret_bytes = (size_t)s.n_ret * n_entries * 3 * sizeof(__half);
            RTO_TRY(d_colors.alloc(col_bytes), "cudaMalloc(quant_colors)");
            RTO_TRY(d_map.alloc(map_bytes), "cudaMalloc(quant_map)");
            RTO_TRY(d_sigma.alloc(n_entries * sizeof(__half)), "cudaMalloc(sigma)");
            RTO_TRY(d_ret.alloc(ret_bytes), "cudaMalloc(data_retained)");
            if (col_bytes) RTO_TRY(cudaMemcpy(d_colors.p, s.quant_colors, col_bytes, cudaMemcpyHostToDevice), "H2D quant_colors");
            if (map_bytes) RTO_TRY(cudaMemcpy(d_map.p, s.quant_map, map_bytes, cudaMemcpyHostToDevice), "H2D quant_map");
            RTO_TRY(cudaMemcpy(d_sigma.p, s.sigma_f16, n_entries * sizeof(__half), cudaMemcpyHostToDevice), "H2D sigma");
            if (ret_bytes) RTO_TRY(cudaMemcpy(d_ret.p, s.retained_f16, ret_bytes, cudaMemcpyHostToDevice), "H2D data_retained");
            build_nodes_kernel<<<blocks_for(n_entries, B), B>>>(d_child.as<int32_t>(), d_sigma.as<__half>(), 1, 0, n_entries,
                                                                s.capacity, nodes.as<uint32_t>(),
                                                                d_cnt.as<unsigned long long>(), d_bad.as<int>());
            RTO_TRY(cudaGetLastError(), "build_nodes");
            build_payload_quant_kernel<<<blocks_for(chunks_total, B), B>>>(d_colors.as<__half>(), d_map.as<uint16_t>(),
                                                                           d_ret.as<__half>(), s.n_q, s.n_ret, s.data_dim,
                                                                           stride, n_entries, payload.as<__half>());
            RTO_TRY(cudaGetLastError(), "build_payload_quant");
        }
        *launches += 2;
        unsigned long long n_leaves = 0;
        int bad = 0;
        RTO_TRY(cudaMemcpy(&n_leaves, d_cnt.p, sizeof n_leaves, cudaMemcpyDeviceToHost), "tree build");
        RTO_TRY(cudaMemcpy(&bad, d_bad.p, sizeof bad, cudaMemcpyDeviceToHost), "tree build");
        if (bad) return fail(RTO_ERR_INVALID, "malformed tree: child offset leaves the node array or the links form a cycle");
        out.n_leaves = (int64_t)n_leaves;
    }
    int depth = 0;
    if (int rc = tree_depth_device(nodes.as<uint32_t>(), s.capacity, &depth, launches, fail)) return rc;
    if (depth < 0) return fail(RTO_ERR_INVALID, "malformed tree: child offset leaves the node array or the links form a cycle");
    if (depth > RTO_COORD_BITS) {
        char buf[96];
        snprintf(buf, sizeof buf, "tree depth %d exceeds %d levels", depth, RTO_COORD_BITS);
        return fail(RTO_ERR_UNSUPPORTED, buf);
    }
    out.max_depth = depth;
    out.stride = stride;
    // sparse brick grid for the marching loop: RTO_DISABLE_GRID=1 skips it, RTO_GRID_BUILD=host uses the host builder
    const char* off = getenv("RTO_DISABLE_GRID");
    const char* how = getenv("RTO_GRID_BUILD");
    if (!(off && off[0] == '1')) {
        int rc;
        if (how && how[0] == 'h' && !quant) rc = grid_build_host(s, depth, out, fail);
        else rc = grid_build_device(nodes.as<uint32_t>(), depth, out, launches, fail);
        if (rc == RTO_OK) rc = grid_leaf_build_device(nodes.as<uint32_t>(), out, launches, fail);
        if (rc) { tree_built_free(out); return rc; }
    }
    out.nodes = nodes.release<uint32_t>();
    out.payload = payload.release<__half>();
    return RTO_OK;
}

}  // namespace rto
