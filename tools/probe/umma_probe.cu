// Diagnostic (not part of the product): one tcgen05.mma (M=128, N=32, K=16, f16 -> f32) on operands placed in shared
// memory with configurable core-matrix strides / start offsets, to pin the K-major SWIZZLE_NONE descriptor semantics
// that rto_denoise_tc.cu relies on (shifted start addresses, arbitrary LBO, overlapping chunks).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probe/umma_probe tools/probe/umma_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t start, uint32_t f_lbo, uint32_t f_sbo) {
    return (uint64_t)((start >> 4) & 0x3fffu) | ((uint64_t)((f_lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((f_sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
struct Cfg { int a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo, swap_fields, N; };

__global__ void probe(const __half* A /*[128+pad rows][16]*/, const __half* B /*[N][16]*/, float* D /*[128][N]*/, Cfg c) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;  // zero everything
    __syncthreads();
    // A element (r,k) at a_off + (r%8)*16 + (r/8)*a_sbo + (k/8)*a_lbo + (k%8)*2
    for (int i = tid; i < 128 * 16; i += blockDim.x) {
        int r = i / 16, k = i % 16;
        *reinterpret_cast<__half*>(sm + c.a_off + (r % 8) * 16 + (r / 8) * c.a_sbo + (k / 8) * c.a_lbo + (k % 8) * 2) = A[i];
    }
    for (int i = tid; i < c.N * 16; i += blockDim.x) {
        int r = i / 16, k = i % 16;
        *reinterpret_cast<__half*>(sm + c.b_off + (r % 8) * 16 + (r / 8) * c.b_sbo + (k / 8) * c.b_lbo + (k % 8) * 2) = B[i];
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;\n" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        uint32_t a0 = smem_u32(sm) + c.a_off, b0 = smem_u32(sm) + c.b_off;
        uint64_t ad = c.swap_fields ? make_desc(a0, c.a_sbo, c.a_lbo) : make_desc(a0, c.a_lbo, c.a_sbo);
        uint64_t bd = c.swap_fields ? make_desc(b0, c.b_sbo, c.b_lbo) : make_desc(b0, c.b_lbo, c.b_sbo);
        uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\nW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\nD_%=:\n\t}\n" ::"r"(smem_u32(&bar)), "r"(0u) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (warp < 4) {
        uint32_t r[32];
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        for (int j = 0; j < c.N; ++j) D[(warp * 32 + lane) * c.N + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;\n" ::"r"(tmem) : "memory");
}

int main() {
    const int N = 32;
    std::vector<__half> hA(128 * 16), hB(N * 16);
    std::vector<float> fA(128 * 16), fB(N * 16), ref(128 * N), out(128 * N);
    srand(1);
    for (int i = 0; i < 128 * 16; ++i) { float v = (rand() % 17 - 8) / 8.f; hA[i] = __float2half(v); fA[i] = v; }
    for (int i = 0; i < N * 16; ++i) { float v = (rand() % 13 - 6) / 4.f; hB[i] = __float2half(v); fB[i] = v; }
    for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < 16; ++k) s += fA[m * 16 + k] * fB[n * 16 + k]; ref[m * N + n] = s; }
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, out.size() * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct { const char* name; Cfg c; } tests[] = {
        {"V0 canonical aligned (lbo=K-chunk stride, sbo=8-row stride)", {0, 2048, 128, 8192, 512, 128, 0, N}},
        {"V1 same data, descriptor fields swapped", {0, 2048, 128, 8192, 512, 128, 1, N}},
        {"V2 A start shifted by 5 rows (80 B), B by 16 B", {80, 2048, 128, 8192 + 16, 512, 128, 0, N}},
        {"V3 irregular large LBO (A 40016 B, B 50032 B)", {0, 40016, 128, 8192, 50032, 128, 0, N}},
        {"V4 sbo=256 (rows groups 256 B apart), lbo=128 (chunks interleaved per group)", {0, 128, 256, 16384, 128, 256, 0, N}},
        {"V5 A start 16*1093, lbo 60000 (like conv1 zero-chunk)", {16 * 1093, 60000, 128, 100000 - 100000 % 16, 512, 128, 0, N}},
    };
    for (auto& t : tests) {
        cudaMemset(dD, 0xff, out.size() * 4);
        probe<<<1, 128, 200 * 1024>>>(dA, dB, dD, t.c);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
        double mx = 0; int bad = 0;
        for (int i = 0; i < 128 * N; ++i) { double d = fabs(out[i] - ref[i]); if (!(d <= 1e-3)) ++bad; if (d > mx || d != d) mx = d; }
        printf("%-80s : %s max|err| %.4g, wrong %d / %d  (D[0][0..3] = %.3f %.3f %.3f %.3f ; ref %.3f %.3f %.3f %.3f)\n", t.name,
               cudaGetErrorString(e), mx, bad, 128 * N, out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
    }
    return 0;
}
