// rto_denoise_tc.cu — GuidanceNet forward on the 5th-gen tensor cores (tcgen05 + TMEM) and the fast kernel filter.
//
// The two 3x3 convolutions of the deployed GuidanceNet (denoiser/network.py:123-168: conv 8->32, relu6, conv 32->8,
// relu6, fp16 storage / fp32 accumulate) are the only dense contraction on the render path.  The reference runs
// them through libtorch/cuDNN (src/denoiser/denoiser.cpp:46).  Here each CTA owns a 60 x TH pixel tile (TH = 10) and
// runs both convolutions as implicit GEMMs with M = pixels:
//
//   * the fp32 input tile (+2 px halo: 64 x (TH+4) pixels x 8 planes) arrives by ONE 3-D tensor-map TMA load
//     (cp.async.bulk.tensor.3d, hardware zero fill outside the image) into shared memory that the activation buffer reuses
//     later, and is converted there — conflict-free — to the operand layout;
//   * the tile (+2 px halo) is staged in shared memory PIXEL-MAJOR with a pitch of 64 pixels, 8 fp16 channels =
//     one 16-byte row of a K-major / no-swizzle core matrix.  Because consecutive pixels are consecutive 16-byte
//     rows, the A operand of filter tap (dy,dx) is the SAME buffer viewed through a descriptor whose start address
//     is shifted by (dy*64+dx) pixels — no im2col copy is ever materialised;
//   * M-tiles are runs of 128 consecutive linear pixel positions (they wrap over image rows; the 2 wrap-around
//     columns per row are computed and discarded);
//   * conv1: per M-tile 5 tcgen05.mma, N = 32, K = 16 = two 8-channel filter taps per instruction (the second K chunk
//     is the first one shifted by the descriptor's leading-dimension offset), accumulators in TMEM columns
//     [32*i, 32*i+32);  epilogue 1 (tcgen05.ld -> +bias, relu6, fp16) writes the 32-channel activation back to
//     shared memory as four 8-channel planes (again 16 B per pixel per plane), zero outside the image;
//   * conv2: the three horizontal taps ride in N: D'[r][dx*8+o] = sum_{dy,c} mid[r+dy*64][c] w2[o][c][dy][dx] needs no
//     x-shift of A, so an M-tile (two image rows) costs 3 dy x 2 k-steps = 6 MMAs (M128 N32 K16, LBO = plane stride) instead
//     of 18; accumulators reuse TMEM columns already drained;  epilogue 2 adds D'[r-1][dx=-1] + D'[r][0] + D'[r+1][dx=+1]
//     (warp shuffles; the two rows that straddle a warp boundary go through 64 bytes of shared memory) -> +bias, relu6,
//     fp16 -> fp32 softmax / guidance;
//   * the M-tiles are software-pipelined through single-use mbarriers: one MMA-issuing warp, two 4-warp epilogue teams
//     (see the kernel);  256 TMEM columns and <= 110 KB of shared memory per CTA => two CTAs per SM.
//
// The kernel filter (denoiser/extension/filtering.cu:108-228) then runs as ONE launch for all levels, separated into
// a horizontal and a vertical box sum of e^{g} * (r,g,b,1) (guidance is in [0,6] after relu6, so the reference's
// max-subtraction is not needed for range).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>

#include <cfloat>
#include <cstdint>
#include <cstring>

#include "rto_internal.h"

namespace rto {

namespace tc {
constexpr int TW = 60, PW = 64;                   // output tile width, smem pitch (= TW + 4)
constexpr int THREADS = 288;                      // warps 0-7: staging + epilogues (two teams of 4), warp 8: MMA issue
constexpr int TMEM_COLS = 256;

// Per tile height TH (rows of output pixels per CTA).  Positions are linear indices y*64+x into the staged tile.
template <int TH>
struct Cfg {
    static constexpr int IN_PX = (TH + 4) * PW + 64;                       // staged input pixels (+ slack read only by discarded rows)
    static constexpr int Q1_MIN = PW + 1;                                  // conv1 positions: x in [1,62], y in [1,TH+2]
    static constexpr int N1_TILES = ((TH + 2) * PW - 2 + 127) / 128;
    static constexpr int Q2_MIN = 2 * PW;                                  // conv2 M-tile j = image rows y = 2+2j, 3+2j (x = 0..63)
    static constexpr int N2_TILES = TH / 2;
    static constexpr int MID_PX = 128 * N1_TILES + 128;                    // positions per 8-channel plane of the conv1 activation
    // shared memory map (bytes)
    static constexpr int OFF_IN = 0;
    static constexpr int OFF_MID = OFF_IN + IN_PX * 16;
    static constexpr int OFF_W1 = OFF_MID + 4 * MID_PX * 16;
    static constexpr int OFF_W2 = OFF_W1 + 9 * 512;
    static constexpr int OFF_ZERO = OFF_W2 + 6 * 1024;                      // zero block: must lie ABOVE every operand start address
    static constexpr int OFF_BIAS = OFF_ZERO + 2048;
    static constexpr int OFF_XCH = OFF_BIAS + 60 * 4;                       // (biases: 40 floats + the same 40 values as 20 half2) ; epilogue-2 neighbour exchange: [team][parity][pair][dir][8] floats
    static constexpr int OFF_BAR = OFF_XCH + 2 * 2 * 2 * 2 * 32;            // mbarriers: c1[N1] | mid[N1] | c2[N2] | weights | input tile
    static constexpr int OFF_TMEM = OFF_BAR + 8 * (2 * N1_TILES + N2_TILES + 2);
    // fp32 input tile as the tensor-map TMA load delivers it: [8 planes][TH+4 rows][SW px].  The box starts at image column
    // bx - 4, not bx - 2: a tiled TMA load needs a 16-byte aligned start in the innermost dimension (probed on B200,
    // tools/probe/tma3d_probe.cu: x0 = -2 faults, x0 = -4 zero-fills), so the rows are SW = 68 floats and tile column x sits at
    // staging column x + 2.  The tile lives where the conv1 activations go later (OFF_MID): it is consumed (converted to fp16,
    // pixel-major) before the first MMA is issued.
    static constexpr int OFF_STAGE = OFF_MID;
    static constexpr int SW = PW + 4;
    static constexpr int STAGE_BYTES = 8 * (TH + 4) * SW * 4;
    static_assert(STAGE_BYTES <= 4 * MID_PX * 16 && OFF_STAGE % 128 == 0, "the fp32 staging tile fits the activation buffer, 128-byte aligned");
    static constexpr int SMEM_BYTES = OFF_TMEM + 8;
    static_assert(N1_TILES * 32 <= TMEM_COLS, "conv1 accumulators must fit the TMEM allocation");
    static_assert(TH % 2 == 0 && N2_TILES <= N1_TILES, "conv2 tiles are row pairs and reuse the TMEM columns of conv1 tile j");
    static_assert(128 * N1_TILES + 129 < IN_PX && Q2_MIN + 128 * N2_TILES + PW < MID_PX, "operand reads stay inside the buffers");
    static constexpr int TMEM_ALLOC = N1_TILES * 32 <= 128 ? 128 : 256;    // columns allocated (power of two)
    static_assert(2 * SMEM_BYTES <= 227 * 1024, "two CTAs per SM");
    static_assert(OFF_W1 % 16 == 0 && OFF_BIAS % 16 == 0, "bulk-copy destinations are 16-byte aligned");
};

// packed weights in global memory: [w1: 9 taps][32 out][8 in] fp16 |
// [w2: 3 dy][2 k-steps][2 chunks][32 rows = dx*8 + out (24 used)][8 in] fp16 | b1 [32] fp32 | b2 [8] fp32 | b1 [32] b2 [8] fp16
constexpr int PK_W1 = 0, PK_W2 = 9 * 512, PK_BIAS = PK_W2 + 6 * 1024, PK_BYTES = PK_BIAS + 60 * 4;
}  // namespace tc

size_t denoise_tc_packed_bytes() { return tc::PK_BYTES; }

__global__ void pack_weights_kernel(const NetDev net, unsigned char* __restrict__ out) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    __half* w1 = reinterpret_cast<__half*>(out + tc::PK_W1);
    __half* w2 = reinterpret_cast<__half*>(out + tc::PK_W2);
    float* bias = reinterpret_cast<float*>(out + tc::PK_BIAS);
    if (tid < 9 * 32 * 8) {  // [tap][co][ci]  <- w1[co][ci][tap]
        const int ci = tid % 8, co = (tid / 8) % 32, t = tid / 256;
        w1[tid] = net.w1[(co * 8 + ci) * 9 + t];
    }
    if (tid < 3 * 2 * 2 * 32 * 8) {  // [dy][k-step s][chunk c][row n = dx*8+co][ci8] <- w2[co][(2s+c)*8+ci][dy*3+dx], rows 24..31 zero
        const int ci = tid % 8, n = (tid / 8) % 32, c = (tid / 256) % 2, ks = (tid / 512) % 2, dy = tid / 1024;
        const int dx = n / 8, co = n % 8;
        w2[tid] = dx < 3 ? net.w2[(co * 32 + (2 * ks + c) * 8 + ci) * 9 + dy * 3 + dx] : __float2half(0.f);
    }
    if (tid < 32) bias[tid] = __half2float(net.b1[tid]);
    if (tid < 8) bias[32 + tid] = __half2float(net.b2[tid]);
    __half* bias_h = reinterpret_cast<__half*>(bias + 40);   // the same biases as fp16 pairs (epilogues add them with HADD2)
    if (tid < 32) bias_h[tid] = net.b1[tid];
    if (tid < 8) bias_h[32 + tid] = net.b2[tid];
}

cudaError_t denoise_tc_pack_weights(const NetDev& net, void* packed_dev, cudaStream_t stream) {
    pack_weights_kernel<<<(3 * 2 * 2 * 32 * 8 + 255) / 256, 256, 0, stream>>>(net, static_cast<unsigned char*>(packed_dev));
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// 8 rows x 16 B core matrices; rows 16 B apart, 8-row groups SBO apart, the two 8-element K chunks LBO apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t start_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((start_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, M = 128
__device__ __forceinline__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine); bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// 3-D tiled tensor-map load global -> shared (TMA): box = the tensor map's boxDim, coordinates in elements (may be negative or
// past the end: out-of-bounds elements are ZERO-filled by the hardware), completion as transaction bytes on `bar`
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                 ::"r"(dst_smem), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n"
        "DONE_%=:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// two conv outputs -> packed fp16 activations, with the reference's rounding points (rto_internal.h NetDev::fused_bias):
//   fused = 0 : half(float(half(acc)) + bias) ; fused = 1 : half(acc + bias) ; then relu6 (exact on fp16 values).
// `hi` = the upper clamp as a half2: 6 inside the image, 0 for a position outside it (its activation must read as zero
// padding for conv2): min(max(h, 0), 0) = 0, so the padding costs no select.
// fused = 0 in PACKED fp16: half(acc) + bias as ONE half2 add.  The reference path rounds float(half(acc)) + float(bias) to
// fp16; both operands are fp16 values, the fp32 sum carries 24 >= 2*11 + 2 significant bits, so rounding it to fp16 equals
// the correctly rounded fp16 sum (double rounding is innocuous for addition at that width) — the same bits as HADD2.
__device__ __forceinline__ uint32_t act_h2(float acc0, float acc1, float b0, float b1, __half2 b01, int fused,
                                           __half2 hi = __half2{__half_raw{0x4600}, __half_raw{0x4600}}) {
    __half2 h;
    if (fused) {
        h = __floats2half2_rn(acc0 + b0, acc1 + b1);   // .x (low 16 bits) = channel 0
    } else {
        h = __hadd2_rn(__floats2half2_rn(acc0, acc1), b01);   // b01 = (b0, b1) as stored: fp16
    }
    h = __hmin2(__hmax2(h, __float2half2_rn(0.f)), hi);
    uint32_t u;
    memcpy(&u, &h, 4);
    return u;
}

__device__ __forceinline__ uint4 pack8(float c0, float c1, float c2, float c3, float c4, float c5, float c6, float c7) {
    const __half2 h0 = __floats2half2_rn(c0, c1), h1 = __floats2half2_rn(c2, c3), h2 = __floats2half2_rn(c4, c5), h3 = __floats2half2_rn(c6, c7);
    uint4 v;
    memcpy(&v.x, &h0, 4); memcpy(&v.y, &h1, 4); memcpy(&v.z, &h2, 4); memcpy(&v.w, &h3, 4);
    return v;
}

// One CTA = one 60 x TH tile, software-pipelined over its 128-position M-tiles with single-use mbarriers:
//   warp 8 (one thread) issues every tcgen05.mma: conv1 tile i -> commit c1[i];  conv2 tile j once the activation
//     tiles j..j+2 it reads are in shared memory (mid[]) -> commit c2[j];
//   warps 0-7 form two epilogue teams of 4 warps (a warp can only read its own 32 TMEM lanes): team k handles conv1
//     tiles k, k+2, ... (TMEM -> +b1, relu6, fp16 -> shared memory, arrive on mid[i]) and then conv2 tiles k, k+2, ...
//     (TMEM -> +b2, relu6 -> softmax / guidance -> global).
// The tensor pipe therefore runs conv1 of later tiles and conv2 of earlier tiles underneath both epilogues.
template <int TH>
__global__ void __launch_bounds__(tc::THREADS, 2)
guidance_net_tc_kernel(const __grid_constant__ CUtensorMap aux_map, const unsigned char* __restrict__ packed, const DenoiseArgs d,
                       int fused_bias, int exp_guidance, int use_tma) {
    using namespace tc;
    using C = Cfg<TH>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = d.W, H = d.H;
    const int bx = blockIdx.x * TW, by = d.y0 + blockIdx.y * TH;   // image coords of output pixel (x=2, y=2) of the tile
    const size_t HW = (size_t)W * H;
    const uint32_t s_base = smem_u32(smem);
    const uint32_t bar_c1 = s_base + C::OFF_BAR, bar_mid = bar_c1 + 8 * C::N1_TILES, bar_c2 = bar_mid + 8 * C::N1_TILES;
    const uint32_t bar_w = bar_c2 + 8 * C::N2_TILES, bar_in = bar_w + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEM);

    // ---- one-time setup: TMEM allocation (warp 0), mbarriers (one thread)
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(s_base + C::OFF_TMEM), "n"(C::TMEM_ALLOC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        for (int i = 0; i < C::N1_TILES; ++i) { mbar_init(bar_c1 + 8 * i, 1); mbar_init(bar_mid + 8 * i, 128); }
        for (int j = 0; j < C::N2_TILES; ++j) mbar_init(bar_c2 + 8 * j, 1);
        mbar_init(bar_w, 1);
        mbar_init(bar_in, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        if (use_tma) {   // the whole fp32 input tile: 8 planes x (TH+4) rows x 64 px starting at image (bx-2, by-2), zero outside
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_in), "n"(C::STAGE_BYTES) : "memory");
            tma_load_3d(s_base + C::OFF_STAGE, &aux_map, bx - 4, by - 2, 0, bar_in);
        }
        // packed weights + biases: two bulk async copies (TMA engine, no registers, no thread instructions) signalling bar_w
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar_w), "n"(PK_BYTES) : "memory");
        bulk_g2s(s_base + C::OFF_W1, packed, PK_BIAS, bar_w);
        bulk_g2s(s_base + C::OFF_BIAS, packed + PK_BIAS, PK_BYTES - PK_BIAS, bar_w);
    }
    // ---- stage weights, zero block and the input tile (fp32 planes -> fp16, 8 channels = 16 B per pixel)
    {
        uint4* z = reinterpret_cast<uint4*>(smem + C::OFF_ZERO);
        if (tid < 128) z[tid] = make_uint4(0, 0, 0, 0);
        uint4* in = reinterpret_cast<uint4*>(smem + C::OFF_IN);
        const bool vec = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(d.aux) & 15) == 0;
        if (use_tma) {
            // fp32 planes in shared memory -> fp16, 8 channels = one 16-byte row per pixel.  Consecutive threads take
            // consecutive pixels: 4-byte reads and 16-byte writes are both bank-conflict free.
            __syncthreads();         // bar_in was initialised by thread 32 a moment ago
            mbar_wait(bar_in, 0);
            const float* st = reinterpret_cast<const float*>(smem + C::OFF_STAGE);
            constexpr int PS = (TH + 4) * C::SW;              // plane stride of [plane][y][column]
            for (int p = tid; p < (TH + 4) * PW; p += THREADS) {
                const float* a = st + (p >> 6) * C::SW + (p & (PW - 1)) + 2;   // tile pixel (x = p & 63, y = p >> 6) = staging column x + 2
                in[p] = pack8(a[0], a[PS], a[2 * PS], a[3 * PS], a[4 * PS], a[5 * PS], a[6 * PS], a[7 * PS]);
            }
            for (int p = (TH + 4) * PW + tid; p < C::IN_PX; p += THREADS) in[p] = make_uint4(0, 0, 0, 0);
        } else if (vec) {
            // item = (row y, aligned group of 4 pixels): gx0 = bx - 4 + 4g covers tile-local x = 4g-2 .. 4g+1
            constexpr int GROUPS = PW / 4 + 1;
            for (int it = tid; it < (TH + 4) * GROUPS; it += THREADS) {
                const int y = it / GROUPS, g = it - y * GROUPS;
                const int gy = by + y - 2, gx0 = bx - 4 + 4 * g;
                float4 c[8];
                if (gy >= 0 && gy < H && gx0 >= 0 && gx0 < W) {
                    const float* a = d.aux + (size_t)gy * W + gx0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) c[k] = __ldg(reinterpret_cast<const float4*>(a + k * HW));
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) c[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const int x0 = 4 * g - 2;
                uint4* row = in + y * PW;
                if (x0 >= 0) {
                    row[x0] = pack8(c[0].x, c[1].x, c[2].x, c[3].x, c[4].x, c[5].x, c[6].x, c[7].x);
                    row[x0 + 1] = pack8(c[0].y, c[1].y, c[2].y, c[3].y, c[4].y, c[5].y, c[6].y, c[7].y);
                }
                if (x0 + 2 < PW) {
                    row[x0 + 2] = pack8(c[0].z, c[1].z, c[2].z, c[3].z, c[4].z, c[5].z, c[6].z, c[7].z);
                    row[x0 + 3] = pack8(c[0].w, c[1].w, c[2].w, c[3].w, c[4].w, c[5].w, c[6].w, c[7].w);
                }
            }
            for (int p = (TH + 4) * PW + tid; p < C::IN_PX; p += THREADS) in[p] = make_uint4(0, 0, 0, 0);
        } else {
            for (int p = tid; p < C::IN_PX; p += THREADS) {
                const int x = p & (PW - 1), y = p >> 6;
                const int gx = bx + x - 2, gy = by + y - 2;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (y < TH + 4 && gx >= 0 && gx < W && gy >= 0 && gy < H) {
                    const float* a = d.aux + (size_t)gy * W + gx;
                    v = pack8(__ldg(a), __ldg(a + HW), __ldg(a + 2 * HW), __ldg(a + 3 * HW), __ldg(a + 4 * HW), __ldg(a + 5 * HW),
                              __ldg(a + 6 * HW), __ldg(a + 7 * HW));
                }
                in[p] = v;
            }
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(bar_w, 0);     // weights and biases have landed (long before this point in practice)
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        if (lane == 0) {
            const uint32_t zero = s_base + C::OFF_ZERO;
            // ---- conv1: per M-tile 5 MMAs, D[128 x 32] += A[128 x 16] * B[32 x 16]^T.  The input has only 8 channels, so
            // the two 8-wide K chunks of one MMA carry two filter taps: chunk 1 = chunk 0 shifted by LBO (one pixel to the
            // right: 16 B, or one row down: 1024 B) in A and by whole 512-byte tap blocks in B.  Tap 8 pairs with zeros.
            constexpr uint32_t idesc1 = make_idesc(32);
#pragma unroll 1
            for (int i = 0; i < C::N1_TILES; ++i) {
                const uint32_t pa = s_base + C::OFF_IN + (uint32_t)(C::Q1_MIN + 128 * i) * 16u;
                const uint32_t pb = s_base + C::OFF_W1;
                const uint32_t dcol = tmem + i * 32;
#pragma unroll
                for (int r = 0; r < 3; ++r) {   // taps (dy, -1) + (dy, 0), dy = r - 1
                    const uint32_t a0 = pa + (uint32_t)(((r - 1) * PW - 1) * 16);
                    umma_f16(dcol, make_desc(a0, 16, 128), make_desc(pb + (3 * r) * 512, 512, 128), idesc1, r > 0);
                }
                {   // taps (-1, +1) + (0, +1)
                    const uint32_t a0 = pa + (uint32_t)((-PW + 1) * 16);
                    umma_f16(dcol, make_desc(a0, PW * 16, 128), make_desc(pb + 2 * 512, 3 * 512, 128), idesc1, 1);
                }
                {   // tap (+1, +1) + zero chunk
                    const uint32_t a0 = pa + (uint32_t)((PW + 1) * 16), b0 = pb + 8 * 512;
                    umma_f16(dcol, make_desc(a0, zero - a0, 128), make_desc(b0, zero - b0, 128), idesc1, 1);
                }
                umma_commit(bar_c1 + 8 * i);
            }
            // ---- conv2: the three horizontal taps ride in N.  D'[r][dx*8+o] = sum_{dy,c} mid[r + dy*64][c] * w2[o][c][dy][dx]
            // needs no x-shift of A, so an M-tile costs 3 dy x 2 k-steps = 6 MMAs (M128 N32 K16) instead of 18; the epilogue
            // adds D'[r-1][dx=-1] + D'[r][dx=0] + D'[r+1][dx=+1].  M-tile j = two image rows (positions 128+128j ..), so both of
            // its end rows are wrap columns whose outputs are discarded.  Tile j reads activation positions of conv1 tiles
            // j-1..j+1 and its accumulator reuses the TMEM columns of conv1 tile j, drained by then.
            constexpr uint32_t idesc2 = make_idesc(32);
            int ready = 0;
#pragma unroll 1
            for (int j = 0; j < C::N2_TILES; ++j) {
                const int need = j + 1 < C::N1_TILES - 1 ? j + 1 : C::N1_TILES - 1;
                for (; ready <= need; ++ready) mbar_wait(bar_mid + 8 * ready, 0);
                tc_fence_after();
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint32_t a0 = s_base + C::OFF_MID + (uint32_t)(2 * ks * C::MID_PX + C::Q2_MIN + 128 * j + (dy - 1) * PW) * 16u;
                        const uint32_t b0 = s_base + C::OFF_W2 + (dy * 2 + ks) * 1024;
                        umma_f16(tmem + j * 32, make_desc(a0, C::MID_PX * 16, 128), make_desc(b0, 512, 128), idesc2, (dy | ks) > 0);
                    }
                }
                umma_commit(bar_c2 + 8 * j);
            }
        }
        __syncwarp();   // lanes 1-31 wait for the issuing lane before the block-wide barrier below
    } else {
        const int team = warp >> 2, sub = warp & 3;
        const int row = sub * 32 + lane;
        const uint32_t tlane = tmem + ((uint32_t)(sub * 32) << 16);
        // ---- epilogue 1: +b1, relu6, fp16 -> four 8-channel planes, zero outside the image
        {
            const float* bias = reinterpret_cast<const float*>(smem + C::OFF_BIAS);
            const __half2* bias_h2 = reinterpret_cast<const __half2*>(bias + 40);
            for (int i = team; i < C::N1_TILES; i += 2) {
                mbar_wait(bar_c1 + 8 * i, 0);
                tc_fence_after();
                uint32_t r[32];
                tmem_ld32(tlane + i * 32, r);
                const int q = C::Q1_MIN + 128 * i + row;
                const int x = q & (PW - 1), y = q >> 6;
                const int gx = bx + x - 2, gy = by + y - 2;
                const bool inside = gx >= 0 && gx < W && gy >= 0 && gy < H;
                const __half2 hi = __float2half2_rn(inside ? 6.f : 0.f);   // outside the image the activation is conv2's zero padding
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t pk[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int ch = c * 8 + 2 * j;
                        pk[j] = act_h2(__uint_as_float(r[ch]), __uint_as_float(r[ch + 1]), bias[ch], bias[ch + 1], bias_h2[ch >> 1], fused_bias, hi);
                    }
                    *reinterpret_cast<uint4*>(smem + C::OFF_MID + (c * C::MID_PX + q) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
                fence_async_smem();    // generic-proxy stores -> visible to the tensor core's async-proxy reads
                tc_fence_before();     // orders this thread's tcgen05.ld before the MMAs that reuse the columns
                mbar_arrive(bar_mid + 8 * i);
            }
        }
        // ---- epilogue 2: combine the three dx blocks of neighbouring rows, +b2, relu6, fp16 -> float ; softmax / guidance
        {
            const float* bias = reinterpret_cast<const float*>(smem + C::OFF_BIAS) + 32;
            const __half2* bias_h2 = reinterpret_cast<const __half2*>(smem + C::OFF_BIAS + 40 * 4) + 16;
            float4* xch = reinterpret_cast<float4*>(smem + C::OFF_XCH) + team * 16;   // [parity][pair][dir] x 2 float4
            int it = 0;
            for (int j = team; j < C::N2_TILES; j += 2, ++it) {
                mbar_wait(bar_c2 + 8 * j, 0);
                tc_fence_after();
                uint32_t r[32];
                tmem_ld32(tlane + j * 32, r);
                // row r of the tile = position q; its output needs block 0 (dx=-1) of row r-1 and block 2 (dx=+1) of row r+1
                float lft[8], rgt[8];
#pragma unroll
                for (int o = 0; o < 8; ++o) {
                    lft[o] = __shfl_up_sync(0xffffffffu, __uint_as_float(r[o]), 1);
                    rgt[o] = __shfl_down_sync(0xffffffffu, __uint_as_float(r[16 + o]), 1);
                }
                // rows 31|32 and 95|96 of the tile (x = 31|32) sit in different warps: swap through shared memory
                float4* xb = xch + (it & 1) * 8 + (sub >> 1) * 4;
                if (!(sub & 1) && lane == 31) {
                    xb[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
                    xb[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
                }
                if ((sub & 1) && lane == 0) {
                    xb[2] = make_float4(__uint_as_float(r[16]), __uint_as_float(r[17]), __uint_as_float(r[18]), __uint_as_float(r[19]));
                    xb[3] = make_float4(__uint_as_float(r[20]), __uint_as_float(r[21]), __uint_as_float(r[22]), __uint_as_float(r[23]));
                }
                if (team == 0) asm volatile("bar.sync 1, 128;\n" ::: "memory");   // named barrier of this 4-warp team
                else asm volatile("bar.sync 2, 128;\n" ::: "memory");
                if ((sub & 1) && lane == 0) {
                    const float4 u0 = xb[0], u1 = xb[1];
                    lft[0] = u0.x; lft[1] = u0.y; lft[2] = u0.z; lft[3] = u0.w; lft[4] = u1.x; lft[5] = u1.y; lft[6] = u1.z; lft[7] = u1.w;
                }
                if (!(sub & 1) && lane == 31) {
                    const float4 u0 = xb[2], u1 = xb[3];
                    rgt[0] = u0.x; rgt[1] = u0.y; rgt[2] = u0.z; rgt[3] = u0.w; rgt[4] = u1.x; rgt[5] = u1.y; rgt[6] = u1.z; rgt[7] = u1.w;
                }
                const int q = C::Q2_MIN + 128 * j + row;
                const int x = q & (PW - 1), y = q >> 6;
                const int gx = bx + x - 2, gy = by + y - 2;
                if (x >= 2 && x < TW + 2 && gx < W && gy < H && gy < d.y1) {
                    float o[8];
#pragma unroll
                    for (int c = 0; c < 8; c += 2) {
                        const float acc0 = (lft[c] + __uint_as_float(r[8 + c])) + rgt[c];
                        const float acc1 = (lft[c + 1] + __uint_as_float(r[9 + c])) + rgt[c + 1];
                        const uint32_t u = act_h2(acc0, acc1, bias[c], bias[c + 1], bias_h2[c >> 1], fused_bias);
                        __half2 h;
                        memcpy(&h, &u, 4);
                        const float2 f = __half22float2(h);
                        o[c] = f.x;
                        o[c + 1] = f.y;
                    }
                    const float mx = fmaxf(fmaxf(o[0], o[1]), fmaxf(o[2], o[3]));
                    float e[4], sum = 0.f;
#pragma unroll
                    for (int l = 0; l < 4; ++l) { e[l] = __expf(o[l] - mx); sum += e[l]; }
                    const float inv = __fdividef(1.0f, sum);   // logits are fp16 values in [0,6]: one fp16 ulp moves a weight by 4e-3
                    const size_t p = (size_t)gy * W + gx;
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        RTO_ST(d.weight_map + l * HW + p, e[l] * inv);
                        // exp_guidance: the map feeds filter_sep_kernel only, which needs E_l = e^{g_l} at every pixel of every
                        // window: computed HERE once per pixel (the filter's threads overlap and would each redo it three times)
                        RTO_ST(d.guidance_map + l * HW + p, exp_guidance ? __expf(o[4 + l]) : o[4 + l]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(C::TMEM_ALLOC) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ separable filter
// out(p) = sum_l w_l(p) * [sum_{q in N_l(p)} E_l(q) rgb(q)] / [sum_{q in N_l(p)} E_l(q)],  E_l = exp(g_l), 0 outside the
// image (filtering.cu:108-228 with exp(g - max)/sum == exp(g)/sum; valid because relu6 bounds g to [0,6]).
// sum_{q in N_l(p)} E_l(q) (rgb(q),1) is a BOX sum of the premultiplied field E_l*(r,g,b,1), so it separates: a horizontal pass (registers, inputs straight from global/L2, 4 adjacent outputs per
// thread) writes H_l[row][col] to shared memory, a vertical pass adds 2S+1 rows of H_l.  Work per pixel drops from
// 164 taps x 4 FMA to 24 x 4 FMA (x 32/24 halo rows) + 24 x 4 FADD.  Summation order differs from the exact kernel
// (fp32, <= 81 positive terms: relative 1e-6), well inside the 1e-3 image tolerance; rto_filter keeps the exact kernel.
namespace fs {
constexpr int BW = 32, BH = 24, R = 4, SROWS = BH + 2 * R;        // 32 staged rows
constexpr int THREADS = 256;                                      // pass 1: 32 rows x 8 groups of 4 px ; pass 2: 32 cols x 8 groups of 3 rows
constexpr int SMEM_BYTES = 4 * SROWS * BW * 16;                    // H[4][32][32] float4 = 64 KB
static_assert(SROWS * (BW / 4) == THREADS && BW * (BH / 3) == THREADS, "thread mapping");
}  // namespace fs

// 12 consecutive pixels of one plane row starting at xs (xs % 4 == 0).  GUARD = false: the CTA's whole halo is inside the
// image and rows are 16-byte aligned, so no bounds logic at all (the case for all but the border tiles).
template <bool GUARD>
__device__ __forceinline__ void load12(const float* __restrict__ plane, int W, bool vec, bool rowin, size_t rowoff, int xs,
                                       float (&v)[12]) {
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int x = xs + 4 * s;
        if (!GUARD || (rowin && vec && x >= 0 && x + 3 < W)) {
            const float4 t = RTO_LD_LAST(reinterpret_cast<const float4*>(plane + rowoff + x));
            v[4 * s] = t.x; v[4 * s + 1] = t.y; v[4 * s + 2] = t.z; v[4 * s + 3] = t.w;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[4 * s + k] = (rowin && x + k >= 0 && x + k < W) ? RTO_LD_LAST(plane + rowoff + x + k) : 0.f;
        }
    }
}

// pass 1: horizontal sums of E_l * (r,g,b,1) for staged row r, outputs x = bx + 4*xg + j
template <bool GUARD>
__device__ __forceinline__ void filter_pass1(const float* __restrict__ aux, const float* __restrict__ guidance, int W, int H,
                                             int bx, int by, int tid, float4* __restrict__ Hs) {
    using namespace fs;
    const size_t HW = (size_t)W * H;
    const bool vec = (W & 3) == 0;                                // planes and rows 16-byte aligned
    const int r = tid >> 3, xg = tid & 7;
    const int gy = by + r - R;
    const int xs = bx + 4 * xg - R;
    const bool rowin = !GUARD || (gy >= 0 && gy < H);
    const size_t rowoff = rowin ? (size_t)gy * W : 0;
    float cr[12], cg[12], cb[12];
    load12<GUARD>(aux, W, vec, rowin, rowoff, xs, cr);
    load12<GUARD>(aux + HW, W, vec, rowin, rowoff, xs, cg);
    load12<GUARD>(aux + 2 * HW, W, vec, rowin, rowoff, xs, cb);
    const int sw = (xg >> 1) & 3;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const int S = l + 1;
        float e[12];   // E_l = e^{g_l}, already exponentiated by the GuidanceNet epilogue; load12 yields 0 outside the image
        load12<GUARD>(guidance + l * HW, W, vec, rowin, rowoff, xs, e);
        // the four windows [R+j-S, R+j+S], j = 0..3, share the core [R+3-S, R+S] (2S-2 taps, none for S = 1): it is summed
        // once and every output adds its own three taps — 8S+40 instead of 32S+16 operations per level
        float4 core = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = R + 3 - S; i <= R + S; ++i) {
            core.x = fmaf(e[i], cr[i], core.x); core.y = fmaf(e[i], cg[i], core.y); core.z = fmaf(e[i], cb[i], core.z); core.w += e[i];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 a = core;
#pragma unroll
            for (int i = R + j - S; i <= R + j + S; ++i) {
                if (i >= R + 3 - S && i <= R + S) continue;   // core tap (compile-time after unrolling)
                a.x = fmaf(e[i], cr[i], a.x); a.y = fmaf(e[i], cg[i], a.y); a.z = fmaf(e[i], cb[i], a.z); a.w += e[i];
            }
            Hs[(l * SROWS + r) * BW + 4 * xg + (j ^ sw)] = a;   // swizzle: conflict-free 16-byte stores
        }
    }
}

__global__ void __launch_bounds__(fs::THREADS, 3) filter_sep_kernel(const float* __restrict__ aux, const float* __restrict__ weight,
                                                                   const float* __restrict__ guidance /* e^{g}: see the net epilogue */, int W, int H, int y0,
                                                                   int y1, float4* __restrict__ out, uchar4* __restrict__ out8) {
    using namespace fs;
    extern __shared__ __align__(16) unsigned char fsm[];
    float4* Hs = reinterpret_cast<float4*>(fsm);                  // [4][SROWS][BW], columns swizzled inside groups of 4
    const int tid = threadIdx.x;
    // Block rows are anchored at ABSOLUTE image rows (multiples of BH), not at the band start: the shared-core sums below add a
    // pixel's window rows in an order that depends on its place in the 3-row group, and a pixel must get the same bits whether
    // it is produced by a full-frame launch or by a band of the tile split.
    const int bx = blockIdx.x * BW, by = (y0 / BH + blockIdx.y) * BH;
    const size_t HW = (size_t)W * H;
    const bool interior = (W & 3) == 0 && bx >= R && bx + BW + R <= W && by >= R && by + BH + R <= H &&
                          (reinterpret_cast<uintptr_t>(aux) & 15) == 0 && (reinterpret_cast<uintptr_t>(guidance) & 15) == 0;
    if (interior) filter_pass1<false>(aux, guidance, W, H, bx, by, tid, Hs);
    else filter_pass1<true>(aux, guidance, W, H, bx, by, tid, Hs);
    __syncthreads();
    {   // ---- pass 2: vertical sums, weights, output
        const int x = tid & 31, ty = (tid >> 5) * 3;
        const int slot = (x & ~3) | ((x & 3) ^ ((x >> 3) & 3));
        const int gx = bx + x;
        float o[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) o[k][0] = o[k][1] = o[k][2] = 0.f;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int S = l + 1;
            // the three row windows [k-S, k+S], k = 0..2, share the core rows [2-S, S]: summed once, then two rows each
            float4 acc[3];
            float4 core = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int rr = 2 - S; rr <= S; ++rr) {
                const float4 v = Hs[(l * SROWS + ty + R + rr) * BW + slot];
                core.x += v.x; core.y += v.y; core.z += v.z; core.w += v.w;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) acc[k] = core;
#pragma unroll
            for (int rr = -S; rr <= 2 + S; ++rr) {
                if (rr >= 2 - S && rr <= S) continue;   // core row
                const float4 v = Hs[(l * SROWS + ty + R + rr) * BW + slot];
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (rr - k >= -S && rr - k <= S) { acc[k].x += v.x; acc[k].y += v.y; acc[k].z += v.z; acc[k].w += v.w; }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int gy = by + ty + k;
                if (gx < W && gy < H && gy >= y0 && gy < y1) {
                    const float w = __fdividef(RTO_LD_LAST(weight + l * HW + (size_t)gy * W + gx), acc[k].w);
                    o[k][0] += acc[k].x * w; o[k][1] += acc[k].y * w; o[k][2] += acc[k].z * w;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int gy = by + ty + k;
            if (gx < W && gy < H && gy >= y0 && gy < y1) {
                RTO_ST(out + (size_t)gy * W + gx, make_float4(o[k][0], o[k][1], o[k][2], 1.0f));
                if (out8) RTO_ST(out8 + (size_t)gy * W + gx, rgba8_of(o[k][0], o[k][1], o[k][2], 1.0f));   // the bytes `-o` writes to the PNG
            }
        }
    }
}

// Tensor map of an aux buffer [8][H][W] fp32 for the kernel's input-tile load: dims (W, H, 8), box (68, TH+4, 8), zero fill.
// Returns false when the buffer cannot be described (row pitch or base not 16-byte aligned): the kernel then stages through
// registers.  Pure host-side encoding (a few hundred ns), done per launch.
// cuTensorMapEncodeTiled is a DRIVER entry point: it is looked up through the runtime (cudaGetDriverEntryPoint) so that the
// library carries no link-time dependency on libcuda.so.1 and still loads — and fails loudly at the first compute call — on
// a machine without a driver (tests/test_abi.py::test_no_gpu_fails_loudly).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static const EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

static bool make_aux_map(CUtensorMap* tm, const float* aux, int W, int H, int th) {
    if ((W & 3) != 0 || (reinterpret_cast<uintptr_t>(aux) & 15) != 0) return false;
    const EncodeTiledFn encode = encode_tiled_fn();
    if (!encode) return false;
    static const bool off = [] { const char* e = getenv("RTO_NET_TMA"); return e && e[0] == '0'; }();   // A/B switch
    if (off) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, 8};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * (cuuint64_t)H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)tc::PW + 4, (cuuint32_t)(th + 4), 8};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(aux), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int TH>
static cudaError_t launch_net_th(const NetDev& net, const void* packed, const DenoiseArgs& d, int rows, bool exp_guidance, cudaStream_t stream) {
    // per device: the opt-in shared-memory size is a per-device function attribute.  Atomic flags: several host threads may
    // drive the same device (volrend_headless --gpu_list 0,0); setting the attribute twice is harmless.
    static std::atomic<bool> attr_set[kMaxDevices] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    dev = dev >= 0 && dev < kMaxDevices ? dev : 0;
    if (!attr_set[dev]) {
        e = cudaFuncSetAttribute(guidance_net_tc_kernel<TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<TH>::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    dim3 grid((d.W + tc::TW - 1) / tc::TW, (rows + TH - 1) / TH);
    alignas(64) CUtensorMap tm;
    memset(&tm, 0, sizeof tm);
    const bool use_tma = make_aux_map(&tm, d.aux, d.W, d.H, TH);
    guidance_net_tc_kernel<TH><<<grid, tc::THREADS, tc::Cfg<TH>::SMEM_BYTES, stream>>>(tm, static_cast<const unsigned char*>(packed), d, net.fused_bias,
                                                                                       exp_guidance ? 1 : 0, use_tma ? 1 : 0);
    return cudaGetLastError();
}

// Tile height.  Measured on B200 with four frames in flight (bench workload): TH 8 / 10 / 12 -> 4747 / 4760 / 4715 frames/s,
// TH 14 -> 4197: the kernel's own time is the same (45 us) for all of them, but at TH = 14 two CTAs take 218 KB of
// shared memory per SM and nothing of the neighbouring frames' kernels can be co-resident.  RTO_NET_TILE_H overrides.
// exp_guidance: write e^{guidance} instead of guidance (the form filter_sep_kernel consumes; rto_net_forward keeps guidance)
cudaError_t launch_guidance_net_tc(const NetDev& net, const void* packed, const DenoiseArgs& d, bool exp_guidance, cudaStream_t stream) {
    const int rows = d.y1 - d.y0;
    if (rows <= 0) return cudaSuccess;
    static const int th = [] {   // thread-safe one-time initialisation
        const char* v = getenv("RTO_NET_TILE_H");
        const int t = v ? atoi(v) : 10;
        return (t == 6 || t == 8 || t == 10 || t == 12 || t == 14) ? t : 10;
    }();
    switch (th) {
        case 6: return launch_net_th<6>(net, packed, d, rows, exp_guidance, stream);
        case 8: return launch_net_th<8>(net, packed, d, rows, exp_guidance, stream);
        case 12: return launch_net_th<12>(net, packed, d, rows, exp_guidance, stream);
        case 14: return launch_net_th<14>(net, packed, d, rows, exp_guidance, stream);
        default: return launch_net_th<10>(net, packed, d, rows, exp_guidance, stream);
    }
}

cudaError_t launch_filter_fast(const float* aux, const float* weight, const float* guidance, int W, int H, int y0, int y1,
                               float4* out, uchar4* out8, cudaStream_t stream) {
    const int rows = y1 - y0;
    if (rows <= 0) return cudaSuccess;
    static std::atomic<bool> attr_set[kMaxDevices] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    dev = dev >= 0 && dev < kMaxDevices ? dev : 0;
    if (!attr_set[dev]) {
        e = cudaFuncSetAttribute(filter_sep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fs::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    dim3 grid((W + fs::BW - 1) / fs::BW, (y1 + fs::BH - 1) / fs::BH - y0 / fs::BH);   // block rows anchored at multiples of BH (see the kernel)
    filter_sep_kernel<<<grid, fs::THREADS, fs::SMEM_BYTES, stream>>>(aux, weight, guidance, W, H, y0, y1, out, out8);
    return cudaGetLastError();
}

}  // namespace rto
