// rto_render_kernel.cuh — ray generation + regular-tracking octree traversal + SH shade/composite + aux/image writes.
// (kernel template + per-SPP launcher; instantiated one SPP per translation unit by rto_render_spp.cu so that the
// variants compile in parallel, dispatched by rto_render.cu)
//
// One launch replaces the reference's render_kernel<SPP> (renderer/src/cuda/volrend.cu:84-213) and everything it
// inlines (rt_core.cuh:195-332, n3tree_query.hpp:13-48, lumisphere.hpp:38-81, pcg32.h).  Design:
//   * a warp owns an 8x4 pixel tile (the reference maps 32 consecutive x to a warp) => coherent rays per warp,
//     aux/image rows written as full 32 B sectors; persistent blocks claim 16x8 super-tiles from a global counter;
//   * production marching loop (GRID = 10 + K): sparse brick grid, one table load + one BYTE load per step (depth | dense
//     flag), the 4-byte leaf word (sigma) only in cells with non-zero sigma, table indices formed by the fp adder
//     (rto_ray.cuh FusedIdx); GRID = 1..3 are the earlier loops, kept selectable for A/B runs and as cross-checks;
//   * tree walker (GRID = 0; trace builds, trees without a grid): per-ray ancestor stack in shared memory, integer-coordinate
//     descent resumed at the common ancestor (rto_ray.cuh) instead of a root restart per step;
//   * thresholds, hit list, counts and the optical-depth state live in per-ray shared-memory scratch (the reference keeps
//     them in 176 B of local memory per thread);
//   * SH payload is a separate fp16 plane padded to 64 B per leaf and is touched only for collided leaves.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>

#include "rto_internal.h"
#include "rto_ray.cuh"

namespace rto {

#ifndef RTO_RENDER_MIN_BLOCKS
#define RTO_RENDER_MIN_BLOCKS 10  // __launch_bounds__ min blocks/SM of the production kernels (10 -> 48 registers, no spill in the loop)
#endif

#ifndef RTO_TILE_W
#define RTO_TILE_W 8
#define RTO_TILE_H 4
#endif
constexpr int kTileW = RTO_TILE_W, kTileH = RTO_TILE_H;      // pixels per warp-tile (8x4 = one lane per pixel; smaller tiles leave
                                                             // lanes idle: a diagnostic of the lock-step cost, tools/tile_log.py)
#ifndef RTO_BLOCK_WARPS
#define RTO_BLOCK_WARPS 4            // warps per block = warp tiles per super-tile (4: 2x2 tiles = 16x8 px, 8: 4x2 = 32x8 px)
#endif
constexpr int kBlockWarps = RTO_BLOCK_WARPS;
constexpr int kSuperX = kBlockWarps / 2, kSuperY = 2;   // warp tiles per super-tile in x / y
constexpr int kBlockThreads = 32 * kBlockWarps;
constexpr int kDefaultBlocksPerSM = 32 / kBlockWarps;      // RESIDENT blocks per SM (the build keeps 48 registers = an occupancy limit of 10).
                                                           // Tuned on B200 with the v10 loop (profiles/r02_ab_blocks_per_sm_v10*.json), single
                                                           // stream ms | 4 frames in flight: 10 blocks 0.1979 | 6580, 9: 0.1931 | 6590,
                                                           // 8: 0.1927 | 6655, 7: 0.2049 | 6685, 6: 0.2248 | 6656; 8 blocks of a 64-register
                                                           // build: 0.1910 | 6257 (its CTAs keep the neighbouring frames' kernels off the SM).
                                                           // (v6 loop, round 1: 8 blocks / 64 regs 5100, 10 / 48 regs 5280, 12 / 40 regs 5275.)

// Per-ray scratch in shared memory, word w of thread t at base[w * kBlockThreads + t] (conflict-free):
//   [0, D]            ancestor stack (D = tree max depth)
//   [D+1, D+1+SPP]    sorted thresholds + FLT_MAX sentinel
//   [.., +SPP)        collided leaf ids      [.., +SPP) collision counts      [.., +3) scratch floats
template <int SPP>
struct SmemRay {
    uint32_t* base;   // &smem[threadIdx.x]
    int off_dst;      // D + 1
    __device__ __forceinline__ uint32_t& stack(int l) { return base[l * kBlockThreads]; }
    __device__ __forceinline__ float& dst(int i) { return reinterpret_cast<float*>(base)[(off_dst + i) * kBlockThreads]; }
    __device__ __forceinline__ uint32_t& hit_leaf(int i) { return base[(off_dst + SPP + 1 + i) * kBlockThreads]; }
    __device__ __forceinline__ float& hit_cnt(int i) { return reinterpret_cast<float*>(base)[(off_dst + 2 * SPP + 1 + i) * kBlockThreads]; }
    // two floats the grid walker touches only when sigma > sigma_thresh (optical depth so far, delta_scale): parked here so
    // the marching loop's registers go to loop invariants instead
    __device__ __forceinline__ float& scratch(int i) { return reinterpret_cast<float*>(base)[(off_dst + 3 * SPP + 1 + i) * kBlockThreads]; }
    static __host__ __device__ int words(int max_depth) { return max_depth + 1 + 3 * SPP + 1 + 3; }   // max_depth = -1: no stack
};

#ifdef RTO_TILE_LOG
// Diagnostic build only (tools/build_variant.sh tilelog "-DRTO_TILE_LOG"): every warp tile logs
// {tile id, SM id, start ns, end ns, max steps over its lanes, sum of steps, marching-loop ns, hits} into rto_tile_log.
static __device__ unsigned long long* rto_tile_log = nullptr;   // one per translation unit (= per SPP)
__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned smid() {
    unsigned r;
    asm volatile("mov.u32 %0, %smid;" : "=r"(r));
    return r;
}
#endif

// Persistent kernel.  Work unit = a 16x8 pixel SUPER-TILE (2x2 warp tiles) claimed by a block from a global counter
// (centre rows first); the block's warps pull the four 8x4 warp tiles of the current super-tile from a shared-memory
// state word without any barrier, so (a) sibling warps march spatially adjacent rays at the same time and share node
// sectors in L1, and (b) a warp slot never idles behind a slower sibling: it moves on and claims the next super-tile.
//   s_state = (super_tile_id << 8) | tiles_taken ; taken == 4 means "next taker refills".
__device__ __forceinline__ bool next_tile(unsigned* s_state, int* g_counter, int lane, int& sid, int& sub) {
    unsigned r = 0;
    if (lane == 0) {
        for (;;) {
            const unsigned old = atomicAdd(s_state, 1u);
            const unsigned k = old & 0xffu;
            if (k < (unsigned)kBlockWarps) { r = (old & ~0xffu) | k; break; }
            if (k == (unsigned)kBlockWarps) {   // this warp refills: claim a new super-tile, publish it with sub-tile 0 taken by itself
                const unsigned nsid = (unsigned)atomicAdd(g_counter, 1);
                atomicExch(s_state, (nsid << 8) | 1u);
                r = nsid << 8;
                break;
            }
            while ((*reinterpret_cast<volatile unsigned*>(s_state) & 0xffu) > (unsigned)kBlockWarps) __nanosleep(32);   // refill in flight
        }
    }
    r = __shfl_sync(0xffffffffu, r, 0);
    sid = (int)(r >> 8);
    sub = (int)(r & 0xffu);
    return true;
}

// Colour of a finished ray from its collided leaves (rt_core.cuh:277-331), background composite (volrend.cu:174-179) and the
// aux [8][H][W] / image [H][W][4] (/ RGBA8) stores (volrend.cu:187-212).  mem.hit_leaf(i) holds LEAF indices here.
template <int SPP, class Mem>
__device__ __forceinline__ void shade_composite_write(const RenderArgs& a, Mem& mem, const float (&vdir)[3], int idx, uint32_t sh_nums) {
    const FrameParams& fp = a.fp;
    float out0 = 0.f, out1 = 0.f, out2 = 0.f, out3 = 0.f;
    if (sh_nums > 0) {
        // accumulate colour (rt_core.cuh:277-331)
        const int bd = a.tree.basis_dim;
        const __half* __restrict__ sh = a.tree.sh;
        const int stride = a.tree.sh_stride;
        if (bd == 9) {
            float b[9];
            sh_basis(9, vdir, b);
            for (int i = 0; i < (int)sh_nums; ++i) {
                const uint4* q = reinterpret_cast<const uint4*>(sh + (size_t)mem.hit_leaf(i) * stride);
                const float c_i = mem.hit_cnt(i);
                uint32_t w[16];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint4 v = __ldg(q + k);
                    w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
                }
#define HV(k) f_half_bits_to_float((w[(k) >> 1] >> (((k) & 1) * 16)) & 0xffffu)
                float rgb[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
#define MB(k) (b[k] * HV(9 * c + (k)))
                    float tmp = b[0] * HV(9 * c);
                    tmp += MB(4) + MB(5) + MB(6) + MB(7) + MB(8);
                    tmp += MB(1) + MB(2) + MB(3);
#undef MB
                    rgb[c] = f_div(c_i, 1.f + f_exp(-tmp));
                }
#undef HV
                out0 += rgb[0]; out1 += rgb[1]; out2 += rgb[2];
                out3 += c_i;
            }
        } else if (bd > 0) {
            float b[25];
#pragma unroll
            for (int k = 0; k < 25; ++k) b[k] = 0.f;
            sh_basis(bd, vdir, b);
            for (int i = 0; i < (int)sh_nums; ++i) {
                const float c_i = mem.hit_cnt(i);
                const __half* h = sh + (size_t)mem.hit_leaf(i) * stride;
                float rgb[3];
                for (int c = 0; c < 3; ++c) {
                    const __half* hc = h + bd * c;
#define MB(k) (b[k] * __half2float(__ldg(hc + (k))))
                    float tmp = b[0] * __half2float(__ldg(hc));
                    if (bd >= 25) tmp += MB(16) + MB(17) + MB(18) + MB(19) + MB(20) + MB(21) + MB(22) + MB(23) + MB(24);
                    if (bd >= 16) tmp += MB(9) + MB(10) + MB(11) + MB(12) + MB(13) + MB(14) + MB(15);
                    if (bd >= 9) tmp += MB(4) + MB(5) + MB(6) + MB(7) + MB(8);
                    if (bd >= 4) tmp += MB(1) + MB(2) + MB(3);
#undef MB
                    rgb[c] = f_div(c_i, 1.f + f_exp(-tmp));
                }
                out0 += rgb[0]; out1 += rgb[1]; out2 += rgb[2];
                out3 += c_i;
            }
        } else {  // RGBA leaves (rt_core.cuh:322-326)
            for (int i = 0; i < (int)sh_nums; ++i) {
                const float c_i = mem.hit_cnt(i);
                const __half* h = sh + (size_t)mem.hit_leaf(i) * stride;
                out0 += __half2float(__ldg(h + 0)) * c_i;
                out1 += __half2float(__ldg(h + 1)) * c_i;
                out2 += __half2float(__ldg(h + 2)) * c_i;
                out3 += c_i;
            }
        }
        constexpr float INV_SPP = 1.0f / SPP;
        out0 = f_mul(out0, INV_SPP); out1 = f_mul(out1, INV_SPP); out2 = f_mul(out2, INV_SPP); out3 = f_mul(out3, INV_SPP);
    }

    // background composite, offscreen branch (volrend.cu:174-179)
    const float remain = f_mul(f_sub(1.f, out3), fp.background);
    out0 = f_add(out0, remain); out1 = f_add(out1, remain); out2 = f_add(out2, remain);

    // aux [8][H][W] (volrend.cu:187-202) and image [H][W][4] (volrend.cu:205-212)
    if (a.aux) {
        const size_t SIZE = (size_t)fp.W * fp.H;
        float* q = a.aux + idx;
        RTO_ST(q, out0); RTO_ST(q + SIZE, out1); RTO_ST(q + 2 * SIZE, out2); RTO_ST(q + 3 * SIZE, out3);
        RTO_ST(q + 4 * SIZE, f_mul(out0, out0)); RTO_ST(q + 5 * SIZE, f_mul(out1, out1));
        RTO_ST(q + 6 * SIZE, f_mul(out2, out2)); RTO_ST(q + 7 * SIZE, f_mul(out3, out3));
    }
    if (a.img) RTO_ST(a.img + idx, make_float4(out0, out1, out2, 1.0f));
    if (a.img8) RTO_ST(a.img8 + idx, rgba8_of(out0, out1, out2, 1.0f));
}

// GRID: march over the sparse brick grid (rto_ray.cuh walk_grid) instead of the ancestor-stack descent.
// GRID: 0 = tree walker, 1 = brick grid read through the 4-byte leaf words, 2 = brick grid read through the byte plane,
// 3 = byte plane + collisions resolved through the leaf-id planes after the march (v9), 10 + K = the same with the
// fused-index look-up for a grid of level K (production; RTO_FUSED_INDEX=0 selects 3, RTO_DEFER_HITS=0 selects 2 and
// RTO_GRID8=0 selects 1 for A/B runs).
// TRACE: write the per-ray traversal record (rto_trace).  TRACE && GRID is the PRODUCTION marcher with the record switched
// on: steps / term / src / t / hits come straight out of walk_grid, and the visited-leaf sequence is produced by locating
// every sample point in the tree as well (walk_grid<VERIFY>), which also cross-checks the grid's depth and sigma per step.
template <int SPP, bool TRACE, int GRID>
__global__ void __launch_bounds__(kBlockThreads, (TRACE ? 4 : (SPP <= 8 ? RTO_RENDER_MIN_BLOCKS : 4)) * 4 / kBlockWarps) render_kernel(const __grid_constant__ RenderArgs a) {
    extern __shared__ uint32_t ray_smem[];
    __shared__ unsigned s_state;
    const int lane = threadIdx.x & 31;
    const FrameParams& fp = a.fp;
    const int rw = a.x1 - a.x0, rh = a.y1 - a.y0;
    const int supers_x = (rw + kSuperX * kTileW - 1) / (kSuperX * kTileW), supers_y = (rh + kSuperY * kTileH - 1) / (kSuperY * kTileH);
    const int n_supers = supers_x * supers_y;
    SmemRay<SPP> mem{ray_smem + threadIdx.x, GRID ? 0 : a.tree.max_depth + 1};
    const uint32_t* __restrict__ nodes = a.tree.nodes;
    if (threadIdx.x == 0) s_state = ((unsigned)atomicAdd(a.tile_counter, 1) << 8);
    __syncthreads();

    for (;;) {
        int sid, sub;
        next_tile(&s_state, a.tile_counter, lane, sid, sub);
        if (sid >= n_supers) break;
        // centre-first row order of super-tiles: 0 -> mid, 1 -> mid-1, 2 -> mid+1, ...  (A feedback order - last frame's
        // per-tile step counts, heaviest first, rebuilt by a counting-sort kernel - was measured on B200: render 0.2728 vs
        // 0.2795 ms, i.e. exactly the 6.5 us the extra sort launch costs; not kept.)
        const int sr = sid / supers_x, sc = sid - sr * supers_x;
        const int mid = supers_y >> 1;
        const int srow = (sr & 1) ? mid - 1 - (sr >> 1) : mid + (sr >> 1);
        const int tc = sc * kSuperX + (sub % kSuperX), row = srow * kSuperY + (sub / kSuperX);
        const int ix = a.x0 + tc * kTileW + (lane & (kTileW - 1));
        const int iy = a.y0 + row * kTileH + (lane / kTileW);
#ifdef RTO_TILE_LOG
        const unsigned long long log_t0 = gtime_ns();
        unsigned log_steps = 0, log_hits = 0;
        unsigned long long log_t1 = log_t0, log_t2 = log_t0;
#endif
        if (ix < a.x1 && iy < a.y1 && lane < kTileW * kTileH) {
            const int idx = iy * fp.W + ix;   // full-frame pixel index: RNG offset and buffer address (volrend.cu:92-95)
            RaySetup rs;
            setup_ray(fp, ix, iy, rs);
            WalkOut wo;
            if (rs.hit) {
                // ctx.rng.advance(idx*SPP) (volrend.cu:157) through the row/column jump-ahead tables
                const AdvanceMap rm = a.adv_rows[iy], cm = a.adv_cols[ix];
                Pcg32 rng{cm.mult * (rm.mult * a.rng_state + rm.plus) + cm.plus, a.rng_inc};
                sorted_thresholds_from<SPP>(rng, mem);
            }
            if (TRACE && a.tr.thresh && rs.hit) {
                for (int i = 0; i < SPP; ++i) a.tr.thresh[(size_t)idx * SPP + i] = mem.dst(i);
            }
            auto sink = [&](uint32_t step, uint32_t leaf) {
                if (a.tr.leaf_seq && (int)step < a.tr.max_seq) a.tr.leaf_seq[(size_t)idx * a.tr.max_seq + step] = (int32_t)leaf;
            };
#ifdef RTO_TILE_LOG
            log_t1 = gtime_ns();
#endif
            if constexpr (GRID >= 10)
                walk_grid_fused<SPP, TRACE, GRID - 10>(nodes, a.tree.grid, mem, rs, fp.step_size, fp.sigma_thresh, wo, sink);
            else if constexpr (GRID != 0)
                walk_grid<SPP, TRACE, GRID >= 2, GRID == 3>(nodes, a.tree.grid, mem, rs, fp.step_size, fp.sigma_thresh, wo, sink);
            else
                walk<SPP, TRACE>(nodes, mem, rs, fp.step_size, fp.sigma_thresh, wo, sink);
            const uint32_t sh_nums = wo.n_hits;
            // cell references -> leaf indices, all lanes together
            if constexpr (GRID >= 10) resolve_hits_fused<SPP, GRID - 10>(a.tree.grid, mem, sh_nums);
            else if constexpr (GRID == 3) resolve_hits<SPP>(a.tree.grid, mem, sh_nums);
#ifdef RTO_TILE_LOG
            log_t2 = gtime_ns();
            log_steps = wo.steps;
            log_hits = wo.n_hits;
#endif

            if (TRACE) {
                const TraceOut& tr = a.tr;
                if (tr.steps) tr.steps[idx] = wo.steps;
                if (tr.term) tr.term[idx] = wo.term;
                if (tr.src_bits) tr.src_bits[idx] = u_bits(wo.src);
                if (tr.t_bits) tr.t_bits[idx] = u_bits(wo.t);
                if (tr.leaf_hash) tr.leaf_hash[idx] = wo.hash;
                if (tr.depth_sum) tr.depth_sum[idx] = wo.depth_sum;
                if (tr.n_hits) tr.n_hits[idx] = sh_nums;
                if (tr.n_loads) tr.n_loads[idx] = wo.n_loads;
                for (int i = 0; i < SPP; ++i) {
                    const bool live = i < (int)sh_nums;
                    if (tr.hit_leaf) tr.hit_leaf[(size_t)idx * SPP + i] = live ? (int32_t)mem.hit_leaf(i) : -1;
                    if (tr.hit_cnt) tr.hit_cnt[(size_t)idx * SPP + i] = live ? (uint32_t)mem.hit_cnt(i) : 0u;
                }
                if (tr.leaf_seq)
                    for (int s = (int)wo.steps; s < tr.max_seq; ++s) tr.leaf_seq[(size_t)idx * tr.max_seq + s] = -1;
            }

            shade_composite_write<SPP>(a, mem, rs.vdir, idx, sh_nums);
        }
        __syncwarp();
#ifdef RTO_TILE_LOG
        if (rto_tile_log) {
            unsigned mx = log_steps, sm = log_steps, hs = log_hits;
            unsigned long long t1 = log_t1, t2 = log_t2;
            for (int o = 16; o; o >>= 1) {
                mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                sm += __shfl_xor_sync(0xffffffffu, sm, o);
                hs += __shfl_xor_sync(0xffffffffu, hs, o);
                t1 = min(t1, __shfl_xor_sync(0xffffffffu, t1, o));
                t2 = max(t2, __shfl_xor_sync(0xffffffffu, t2, o));
            }
            if (lane == 0) {
                unsigned long long* q = rto_tile_log + (size_t)(sid * kBlockWarps + sub) * 8;
                q[0] = (unsigned long long)(sid * kBlockWarps + sub); q[1] = smid(); q[2] = log_t0; q[3] = gtime_ns();
                q[4] = mx; q[5] = sm; q[6] = t2 - t1; q[7] = hs;
            }
        }
#endif
    }
    // the last warp to leave re-arms the counters for the next launch on this context
    if (lane == 0) {
        const int total_warps = gridDim.x * (kBlockThreads / 32);
        if (atomicAdd(a.tile_counter + 1, 1) == total_warps - 1) {
            a.tile_counter[0] = 0;
            a.tile_counter[1] = 0;
            __threadfence();
        }
    }
}

// Resident blocks per SM of the persistent kernel.  More warps raise issue utilisation but every warp then advances
// more slowly, and the frame time is bounded below by the LONGEST ray's serial chain, so the optimum
// is well below the occupancy limit.  RTO_RENDER_BLOCKS_PER_SM overrides the default for tuning.
static int tuned_blocks_per_sm(int occ_limit) {
    static const int env = [] {
        const char* e = getenv("RTO_RENDER_BLOCKS_PER_SM");
        return e ? atoi(e) : 0;
    }();
    int want = env > 0 ? env : kDefaultBlocksPerSM;
    return want < occ_limit ? want : occ_limit;
}

template <int SPP>
static cudaError_t launch_spp(const RenderArgs& a, int trace, cudaStream_t stream) {   // trace: 0 off, 1 tree walker, 2 production marcher
    const int rw = a.x1 - a.x0, rh = a.y1 - a.y0;
    if (rw <= 0 || rh <= 0) return cudaSuccess;
    const bool grid_path = trace != 1 && a.tree.grid.K > 0;
    const char* g8 = getenv("RTO_GRID8");   // read per launch so that one process can A/B the two planes
    const bool grid8 = grid_path && !(g8 && g8[0] == '0') && a.tree.grid.bricks8 != nullptr;
    const char* dh = getenv("RTO_DEFER_HITS");
    const bool defer = grid8 && !(dh && dh[0] == '0') && a.tree.grid.leaf_top != nullptr;
    const char* fi = getenv("RTO_FUSED_INDEX");
    const int K = a.tree.grid.K;
    const bool fused = defer && !(fi && fi[0] == '0') && a.tree.grid.top_m != nullptr && K >= 1 && K <= 8;
    constexpr int kVariants = 12;   // 0..3: walker / words / bytes / deferred ; 4..11: fused index, K = 1..8
    const int v = (trace ? kVariants : 0) + (fused ? 3 + K : grid_path ? (grid8 ? (defer ? 3 : 2) : 1) : 0);
    const size_t smem = (size_t)SmemRay<SPP>::words(grid_path ? -1 : a.tree.max_depth) * kBlockThreads * sizeof(uint32_t);
    // Function attributes, occupancy and the L2 set-aside are per DEVICE, so the cached launch state is indexed by the
    // current device; the one-time set-up of a slot runs under that slot's mutex (several host threads may drive the same
    // device: volrend_headless --gpu_list 0,0, or a pipelined caller with one thread per stream).
    struct DevState {
        std::mutex mu;
        int num_sms = 0;
        size_t smem_set[2 * 12] = {};
        int occ_limit[2 * 12] = {};
        int persist = -1, max_win = 0, max_persist = 0;
        size_t persist_set = 0;   // current cudaLimitPersistingL2CacheSize this library asked for on the device
    };
    static DevState dev_state[kMaxDevices];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    DevState& ds = dev_state[dev >= 0 && dev < kMaxDevices ? dev : 0];
    void (*kern)(RenderArgs) =
        trace ? (defer ? render_kernel<SPP, true, 3> : grid8 ? render_kernel<SPP, true, 2> : grid_path ? render_kernel<SPP, true, 1> : render_kernel<SPP, true, 0>)
              : (defer ? render_kernel<SPP, false, 3> : grid8 ? render_kernel<SPP, false, 2> : grid_path ? render_kernel<SPP, false, 1> : render_kernel<SPP, false, 0>);
    if (fused) {
        static void (*const fused_kern[2][8])(RenderArgs) = {
            {render_kernel<SPP, false, 11>, render_kernel<SPP, false, 12>, render_kernel<SPP, false, 13>, render_kernel<SPP, false, 14>,
             render_kernel<SPP, false, 15>, render_kernel<SPP, false, 16>, render_kernel<SPP, false, 17>, render_kernel<SPP, false, 18>},
            {render_kernel<SPP, true, 11>, render_kernel<SPP, true, 12>, render_kernel<SPP, true, 13>, render_kernel<SPP, true, 14>,
             render_kernel<SPP, true, 15>, render_kernel<SPP, true, 16>, render_kernel<SPP, true, 17>, render_kernel<SPP, true, 18>}};
        kern = fused_kern[trace ? 1 : 0][K - 1];
    }
    std::unique_lock<std::mutex> lock(ds.mu);
    if (smem > ds.smem_set[v] || ds.occ_limit[v] == 0) {   // first launch on this device, or a deeper tree than any seen before
        if ((e = cudaDeviceGetAttribute(&ds.num_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        int occ = 0;
        if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBlockThreads, smem)) != cudaSuccess) return e;
        ds.occ_limit[v] = occ > 0 ? occ : 1;
        ds.smem_set[v] = smem;
        // Shared-memory carve-out: left to the driver, which sizes it for the kernel's occupancy limit (10 blocks) although
        // the launch keeps 8 resident.  Asking for just what the resident blocks need (more L1 for the brick planes) was
        // measured (profiles/r02_ab_smem_carveout.json): render 0.1907 vs 0.1920 ms alone, but 5670 vs 6620 frames/s with four
        // frames in flight — the GuidanceNet / filter CTAs of the neighbouring frames no longer find shared memory on the SM.
        // RTO_SMEM_CARVEOUT=<percent> applies a preference for A/B runs.
        if (const char* ce = getenv("RTO_SMEM_CARVEOUT")) {
            const int pct = atoi(ce);
            if (pct > 0 && pct <= 100 && cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct) != cudaSuccess)
                (void)cudaGetLastError();   // only a preference
        }
    }
    const int n_supers = ((rw + kSuperX * kTileW - 1) / (kSuperX * kTileW)) * ((rh + kSuperY * kTileH - 1) / (kSuperY * kTileH));
    int grid = ds.num_sms * tuned_blocks_per_sm(ds.occ_limit[v]);
    const int need = n_supers;
    if (grid > need) grid = need;
    // L2 persistence window over the brick array (RTO_L2_PERSIST=0 turns it off): keeps as much of the grid as the
    // device allows resident in the 126 MB L2 while the per-frame buffers (aux, maps, image) stream through.
    if (ds.persist < 0) {
        const char* ev = getenv("RTO_L2_PERSIST");
        ds.persist = (ev && ev[0] == '0') ? 0 : 1;
        if (ds.persist) {
            cudaDeviceGetAttribute(&ds.max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
            cudaDeviceGetAttribute(&ds.max_win, cudaDevAttrMaxAccessPolicyWindowSize, dev);
            if (ds.max_persist <= 0) ds.persist = 0;
        }
    }
    const bool use_window = ds.persist && grid_path && a.tree.grid_brick_bytes > 0;
    // the plane the marching loop reads on (almost) every step: the byte bricks when they are in use
    size_t bytes = grid8 ? a.tree.grid_brick_bytes / sizeof(uint32_t) : a.tree.grid_brick_bytes;
    if (use_window) {
        // L2 set-aside for persisting lines: a process-wide DEVICE limit (documented in rtoctree_b200.h).  It is sized to the
        // plane it protects, grown only when a larger tree is rendered, never above the device maximum.
        size_t want = bytes < (size_t)ds.max_persist ? bytes : (size_t)ds.max_persist;
        want = (want + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        if (want > (size_t)ds.max_persist) want = (size_t)ds.max_persist;
        if (want > ds.persist_set) {
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) ds.persist_set = want;
            else { ds.persist = 0; (void)cudaGetLastError(); }
        }
    }
    const size_t persist_set = ds.persist_set;
    const int max_win = ds.max_win;
    lock.unlock();
    if (use_window && persist_set > 0) {
        if (max_win > 0 && bytes > (size_t)max_win) bytes = (size_t)max_win;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kBlockThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeAccessPolicyWindow;
        at[0].val.accessPolicyWindow.base_ptr = grid8 ? (void*)const_cast<uint8_t*>(a.tree.grid.bricks8) : (void*)const_cast<uint32_t*>(a.tree.grid.bricks);
        at[0].val.accessPolicyWindow.num_bytes = bytes;
        at[0].val.accessPolicyWindow.hitRatio = bytes <= persist_set ? 1.0f : (float)persist_set / (float)bytes;
        at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kern, a);
    }
    kern<<<grid, kBlockThreads, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace rto
