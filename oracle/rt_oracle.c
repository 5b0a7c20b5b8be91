/*
 * TEST INFRASTRUCTURE — not part of the product.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's library.
 *
 * rt_oracle.c — CPU restatement (plain C, scalar, one ray at a time) of the RT-Octree render hot path:
 *
 *   pcg32 seed / advance / next_uint          renderer/3rdparty/pcg32.h:53-68,103-112,145-166
 *   screen2worlddir, maybe_world2ndc          renderer/src/cuda/volrend.cu:24-56
 *   render_kernel<SPP> (offscreen branch)     renderer/src/cuda/volrend.cu:84-213
 *   _dda_world, _dda_unit, _get_delta_scale   renderer/include/volrend/cuda/rt_core.cuh:19-65
 *   sample_dst<SPP>                           renderer/include/volrend/cuda/rt_core.cuh:67-193
 *   trace_ray<float,SPP>                      renderer/include/volrend/cuda/rt_core.cuh:195-332
 *   query_single_from_root                    renderer/include/volrend/internal/n3tree_query.hpp:13-48
 *   maybe_precalc_basis (SH branch)           renderer/include/volrend/internal/lumisphere.hpp:38-81
 *   GuidanceNetCompact forward (deployed)     denoiser/network.py:123-168,170-208 (fp16 graph, Appendix B)
 *   applying<.,16,32,SUPPORT> (filter fwd)    denoiser/extension/filtering.cu:108-228, 701-717
 *
 * Arithmetic follows, op for op, the PTX nvcc 12.9 emits for the reference kernel at -O3 (default
 * -fmad=true): every fused multiply-add the compiler formed is an explicit fmaf() here, every
 * non-fused multiply/add is a plain C operation, and this file MUST be compiled with
 * -ffp-contract=off so the C compiler forms no others.  The fp64 detours of the reference (invdir,
 * _dda_world, SH basis constants) are kept.  On IEEE hardware +,-,*,/,sqrt,fma,cvt are correctly
 * rounded on both sides, so given the same thresholds dst[] the traversal (leaf sequence, step count,
 * termination index, accumulated optical depth bits) is reproducible BIT-FOR-BIT against the GPU.
 * The one thing a CPU cannot reproduce is MUFU lg2.approx used for dst[] (rt_core.cuh:75 `__logf`):
 * callers either inject GPU-computed thresholds (`thresh`) or accept log2f()-based ones.
 *
 * PARITY PINNING: the reference ships no tests/golden vectors.  This restatement is
 * pinned against (1) oracle/_ref/libref_cpu.so = the reference's own rt_core.cuh host-compiled
 * (tests/test_oracle_pin.py, CPU) and (2) the unmodified reference CUDA binary oracle/_ref/ref_driver
 * on the GPU box (tests/test_gpu_vs_reference.py).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RTO_MAX_SPP 32
#define RTO_BASIS_MAX 25

/* ------------------------------------------------------------------------------------------------ */
/* fp16 helpers (gcc _Float16: round-to-nearest-even conversions, IEEE binary16)                     */
typedef _Float16 f16;
static inline float h2f(uint16_t bits) { f16 h; memcpy(&h, &bits, 2); return (float)h; }
static inline float round_f16(float x) { return (float)(f16)x; }

/* ------------------------------------------------------------------------------------------------ */
/* pcg32 (pcg32.h:53-68, 145-166)                                                                    */
#define PCG32_MULT 0x5851f42d4c957f2dULL
typedef struct { uint64_t state, inc; } pcg32_t;

static inline uint32_t pcg32_next_uint(pcg32_t* r) {
    uint64_t old = r->state;
    r->state = old * PCG32_MULT + r->inc;
    uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((~rot + 1u) & 31));
}
void rto_oracle_pcg32_seed(uint64_t initstate, uint64_t initseq, uint64_t* state, uint64_t* inc) {
    pcg32_t r;
    r.state = 0u;
    r.inc = (initseq << 1u) | 1u;
    pcg32_next_uint(&r);
    r.state += initstate;
    pcg32_next_uint(&r);
    *state = r.state;
    *inc = r.inc;
}
static inline void pcg32_advance(pcg32_t* r, uint64_t delta) {
    uint64_t cur_mult = PCG32_MULT, cur_plus = r->inc, acc_mult = 1u, acc_plus = 0u;
    while (delta > 0) {
        if (delta & 1) {
            acc_mult *= cur_mult;
            acc_plus = acc_plus * cur_mult + cur_plus;
        }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        delta /= 2;
    }
    r->state = acc_mult * r->state + acc_plus;
}
uint64_t rto_oracle_pcg32_advance(uint64_t state, uint64_t inc, uint64_t delta) {
    pcg32_t r = {state, inc};
    pcg32_advance(&r, delta);
    return r.state;
}
/* next_float (pcg32.h:103-112) */
static inline float pcg32_next_float(pcg32_t* r) {
    uint32_t u = (pcg32_next_uint(r) >> 9) | 0x3f800000u;
    float f;
    memcpy(&f, &u, 4);
    return f - 1.0f;
}

/* sorted thresholds for one pixel; CPU stand-in for `-__logf(1 - u)` = lg2.approx(1-u) * -ln2          */
static void sample_dst_cpu(pcg32_t* rng, int spp, float* dst) {
    for (int i = 0; i < spp; ++i) {
        float u = pcg32_next_float(rng);
        float t = log2f(1.0f - u) * -0.6931472f;
        int j = i; /* insertion sort ascending (rt_core.cuh:67-185; the result is order-independent) */
        while (j > 0 && dst[j - 1] > t) { dst[j] = dst[j - 1]; --j; }
        dst[j] = t;
    }
    dst[spp] = FLT_MAX;
}

/* the uniform draws themselves are exact integer work: expose them so tests can check the GPU's RNG */
void rto_oracle_uniform_bits(uint64_t state, uint64_t inc, int pix_begin, int pix_end, int spp, uint32_t* out) {
    for (int idx = pix_begin; idx < pix_end; ++idx) {
        pcg32_t r = {state, inc};
        pcg32_advance(&r, (uint64_t)(int64_t)(idx * spp));
        for (int i = 0; i < spp; ++i) out[(size_t)(idx - pix_begin) * spp + i] = pcg32_next_uint(&r);
    }
}

/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    const int32_t* child;   /* [cap*8] relative node offsets, 0 = leaf                               */
    const uint16_t* data;   /* [cap*8*data_dim] fp16 bits, AoS as in tree.npz                         */
    int data_dim, basis_dim; /* basis_dim<=0: RGBA format                                             */
    float offset[3], scale[3];
    float ndc_width, ndc_height, ndc_focal; /* ndc_width<=0: off                                     */
} tree_t;

typedef struct {
    uint32_t* steps;      /* [n] loop iterations (leaf visits)                                        */
    int32_t* term;        /* [n] step index at which the SPP-th collision ended the ray, else -1      */
    uint32_t* src_bits;   /* [n] fp32 bits of the accumulated optical depth `src` at exit             */
    uint32_t* t_bits;     /* [n] fp32 bits of `t` at exit                                             */
    uint64_t* leaf_hash;  /* [n] FNV-1a (64-bit) over the int32 leaf indices visited, in order        */
    uint32_t* depth_sum;  /* [n] sum over steps of the number of child look-ups                       */
    uint32_t* n_hits;     /* [n] sh_nums = number of (not nec. distinct) collided leaves              */
    int32_t* hit_leaf;    /* [n][spp] flat leaf index (node*8+child) per collision entry, -1 padded   */
    uint32_t* hit_cnt;    /* [n][spp] collisions counted in that entry                                */
    int32_t* leaf_seq;    /* [n][max_seq] first max_seq visited leaf indices, -1 padded (may be NULL) */
    int max_seq;
} trace_t;

#define FNV_OFFSET 0xcbf29ce484222325ULL
#define FNV_PRIME 0x100000001b3ULL
static inline uint64_t fnv_i32(uint64_t h, int32_t v) {
    uint32_t u = (uint32_t)v;
    for (int b = 0; b < 4; ++b) { h ^= (u >> (8 * b)) & 0xffu; h *= FNV_PRIME; }
    return h;
}

/* maybe_precalc_basis, SH branch (lumisphere.hpp:38-81): double literals => fp64 products, fp32 results */
static void sh_basis(int basis_dim, const float* dir, float* out) {
    out[0] = (float)0.28209479177387814;
    const float x = dir[0], y = dir[1], z = dir[2];
    const float xx = x * x, yy = y * y, zz = z * z;
    const float xy = x * y, yz = y * z, xz = x * z;
    if (basis_dim >= 25) {
        out[16] = (float)(2.5033429417967046 * xy * (xx - yy));
        out[17] = (float)(-1.7701307697799304 * yz * (3 * xx - yy));
        out[18] = (float)(0.9461746957575601 * xy * (7 * zz - 1.f));
        out[19] = (float)(-0.6690465435572892 * yz * (7 * zz - 3.f));
        out[20] = (float)(0.10578554691520431 * (zz * (35 * zz - 30) + 3));
        out[21] = (float)(-0.6690465435572892 * xz * (7 * zz - 3));
        out[22] = (float)(0.47308734787878004 * (xx - yy) * (7 * zz - 1.f));
        out[23] = (float)(-1.7701307697799304 * xz * (xx - 3 * yy));
        out[24] = (float)(0.6258357354491761 * (xx * (xx - 3 * yy) - yy * (3 * xx - yy)));
    }
    if (basis_dim >= 16) {
        out[9] = (float)(-0.5900435899266435 * y * (3 * xx - yy));
        out[10] = (float)(2.890611442640554 * xy * z);
        out[11] = (float)(-0.4570457994644658 * y * (4 * zz - xx - yy));
        out[12] = (float)(0.3731763325901154 * z * (2 * zz - 3 * xx - 3 * yy));
        out[13] = (float)(-0.4570457994644658 * x * (4 * zz - xx - yy));
        out[14] = (float)(1.445305721320277 * z * (xx - yy));
        out[15] = (float)(-0.5900435899266435 * x * (xx - 3 * yy));
    }
    if (basis_dim >= 9) {
        out[4] = (float)(1.0925484305920792 * xy);
        out[5] = (float)(-1.0925484305920792 * yz);
        out[6] = (float)(0.31539156525252005 * (2.0 * zz - xx - yy));
        out[7] = (float)(-1.0925484305920792 * xz);
        out[8] = (float)(0.5462742152960396 * (xx - yy));
    }
    if (basis_dim >= 4) {
        out[1] = (float)(-0.4886025119029199 * y);
        out[2] = (float)(0.4886025119029199 * z);
        out[3] = (float)(-0.4886025119029199 * x);
    }
}

/* One ray.  `dst` holds spp sorted thresholds + FLT_MAX sentinel.  Writes out[4] = rgb*,alpha.        */
static void trace_ray(const tree_t* tree, float* dir, const float* vdir, const float* cen, int spp,
                      float step_size, float sigma_thresh, const float* dst, float* out, trace_t* tr,
                      size_t ridx) {
    /* _get_delta_scale (rt_core.cuh:53-65) */
    float d0 = dir[0] * tree->scale[0], d1 = dir[1] * tree->scale[1], d2 = dir[2] * tree->scale[2];
    float nrm = sqrtf(fmaf(d2, d2, fmaf(d0, d0, d1 * d1)));
    const float delta_scale = 1.0f / nrm; /* rcp.rn */
    dir[0] = d0 * delta_scale; dir[1] = d1 * delta_scale; dir[2] = d2 * delta_scale;
    const float tmax_bg = 1e9f / delta_scale; /* rt_core.cuh:208 */

    /* invdir (rt_core.cuh:212-215): double add + double reciprocal, rounded to float */
    float invdir[3];
    for (int i = 0; i < 3; ++i) invdir[i] = (float)(1.0 / ((double)dir[i] + 1e-9));

    /* _dda_world with render_bbox = {0,0,0,1,1,1} (rt_core.cuh:19-36; render_options.hpp:30) in double */
    float tmin = 0.0f, tmax = 1e4f;
    for (int i = 0; i < 3; ++i) {
        float t1 = (float)((((double)0.0f + 1e-6) - (double)cen[i]) * (double)invdir[i]);
        float t2 = (float)((((double)1.0f - 1e-6) - (double)cen[i]) * (double)invdir[i]);
        tmin = fmaxf(tmin, fminf(t1, t2));
        tmax = fminf(tmax, fmaxf(t1, t2));
    }
    tmax = fminf(tmax, tmax_bg);

    uint32_t steps = 0, depth_sum = 0, sh_nums = 0, nspp = 0;
    int32_t term = -1;
    float src = 0.0f, t = tmin;
    uint64_t hash = FNV_OFFSET;
    int64_t hit_leaf[RTO_MAX_SPP];
    float cnts[RTO_MAX_SPP];
    for (int i = 0; i < spp; ++i) { cnts[i] = 0.f; hit_leaf[i] = -1; }

    if (!(tmax < 0 || tmin > tmax)) {
        while (t < tmax) {
            /* pos = cen + t*dir : fma (PTX) ; clamp (n3tree_query.hpp:17-19) */
            float p[3];
            for (int i = 0; i < 3; ++i) {
                float v = fmaf(t, dir[i], cen[i]);
                v = fminf(v, 1.f - 1e-6f);
                p[i] = fmaxf(v, 0.f);
            }
            /* query_single_from_root, N = 2 */
            int64_t ptr = 0, sub_ptr;
            float cube_sz = 2.0f;
            uint32_t depth = 0;
            for (;;) {
                float index = 0.f;
                for (int i = 0; i < 3; ++i) {
                    p[i] *= 2.0f;
                    const float fl = floorf(p[i]);
                    index = fmaf(index, 2.0f, fl); /* i==0: 0*2+fl (exact either way) */
                    p[i] -= fl;
                }
                sub_ptr = ptr + (int32_t)index;
                const int64_t skip = tree->child[sub_ptr];
                ++depth;
                if (skip == 0) break;
                cube_sz *= 2.0f;
                ptr += skip * 8;
            }
            /* _dda_unit (rt_core.cuh:38-51): t1 = -x*invdir (mul), t2 = t1 + invdir (add, not fused) */
            float tu = 1e4f;
            for (int i = 0; i < 3; ++i) {
                float t1 = -p[i] * invdir[i];
                float t2 = t1 + invdir[i];
                tu = fminf(tu, fmaxf(t1, t2));
            }
            const float t_subcube = tu / cube_sz;
            const float delta_t = t_subcube + step_size;
            const float sigma = h2f(tree->data[(size_t)sub_ptr * tree->data_dim + tree->data_dim - 1]);

            hash = fnv_i32(hash, (int32_t)sub_ptr);
            if (tr && tr->leaf_seq && (int)steps < tr->max_seq)
                tr->leaf_seq[ridx * tr->max_seq + steps] = (int32_t)sub_ptr;
            depth_sum += depth;
            ++steps;

            if (sigma > sigma_thresh) {
                /* delta = (delta_t*delta_scale)*sigma ; src+delta contracted to ONE fma whose value is both
                 * compared with dst[] and stored back as src (PTX of rt_core.cuh:252-266) */
                const float a = delta_scale * delta_t;
                const float s_new = fmaf(a, sigma, src);
                if (s_new >= dst[nspp]) {
                    hit_leaf[sh_nums] = sub_ptr;
                    float* cnt = &cnts[sh_nums];
                    ++sh_nums;
                    do { *cnt += 1.0f; ++nspp; } while (s_new >= dst[nspp]);
                    if ((int)nspp == spp) { term = (int32_t)(steps - 1); src = s_new; break; }
                }
                src = s_new;
            }
            t = t + delta_t;
        }
    }

    if (tr) {
        if (tr->steps) tr->steps[ridx] = steps;
        if (tr->term) tr->term[ridx] = term;
        if (tr->src_bits) memcpy(&tr->src_bits[ridx], &src, 4);
        if (tr->t_bits) memcpy(&tr->t_bits[ridx], &t, 4);
        if (tr->leaf_hash) tr->leaf_hash[ridx] = hash;
        if (tr->depth_sum) tr->depth_sum[ridx] = depth_sum;
        if (tr->n_hits) tr->n_hits[ridx] = sh_nums;
        for (int i = 0; i < spp; ++i) {
            if (tr->hit_leaf) tr->hit_leaf[ridx * spp + i] = (int32_t)hit_leaf[i];
            if (tr->hit_cnt) tr->hit_cnt[ridx * spp + i] = (uint32_t)cnts[i];
        }
        if (tr->leaf_seq)
            for (int s = (int)steps; s < tr->max_seq; ++s) tr->leaf_seq[ridx * tr->max_seq + s] = -1;
    }
    if (sh_nums == 0) return;

    /* accumulate colour (rt_core.cuh:277-331) */
    float basis[RTO_BASIS_MAX];
    for (int i = 0; i < RTO_BASIS_MAX; ++i) basis[i] = 0.f;
    const int bd = tree->basis_dim;
    if (bd > 0) sh_basis(bd, vdir, basis);
    for (uint32_t i = 0; i < sh_nums; ++i) {
        const uint16_t* tv = tree->data + (size_t)hit_leaf[i] * tree->data_dim;
        if (bd > 0) {
            int off = 0;
            for (int c = 0; c < 3; ++c) {
#define MB(k) (basis[k] * h2f(tv[off + (k)]))
                float tmp = basis[0] * h2f(tv[off]);
                if (bd >= 25) tmp += MB(16) + MB(17) + MB(18) + MB(19) + MB(20) + MB(21) + MB(22) + MB(23) + MB(24);
                if (bd >= 16) tmp += MB(9) + MB(10) + MB(11) + MB(12) + MB(13) + MB(14) + MB(15);
                if (bd >= 9) tmp += MB(4) + MB(5) + MB(6) + MB(7) + MB(8);
                if (bd >= 4) tmp += MB(1) + MB(2) + MB(3);
#undef MB
                out[c] += cnts[i] / (1.f + exp2f(-tmp * 1.442695f)); /* __expf = ex2.approx(x*log2e) */
                off += bd;
            }
        } else {
            for (int j = 0; j < 3; ++j) out[j] += h2f(tv[j]) * cnts[i];
        }
        out[3] += cnts[i];
    }
    const float inv_spp = 1.0f / (float)spp;
    out[0] *= inv_spp; out[1] *= inv_spp; out[2] *= inv_spp; out[3] *= inv_spp;
}

/*
 * Render pixels [pix_begin, pix_end) of one frame (render_kernel<SPP>, offscreen branch).
 *   c2w12    column-major 4x3 (right, up, back, centre), as uploaded by Camera::_update (camera.cpp:72-73)
 *   rng_*    frame-level pcg32 state (ctx.rng, passed by value into the kernel: volrend.cu:90,157)
 *   thresh   optional [pix_end-pix_begin][spp] sorted thresholds (GPU lg2.approx values); NULL = CPU log2f
 *   aux      [8][H][W]; img [H][W][4] (may be NULL).  Only the pixel range is written.
 *   trace arrays are indexed by (idx - pix_begin).
 * returns 0, -1 for an spp the reference does not instantiate (volrend.cu:266-278).
 */
int rto_oracle_render(const int32_t* child, const uint16_t* data, int data_dim, int basis_dim,
                      const float* offset, const float* scale, float ndc_width, float ndc_height,
                      float ndc_focal, const float* c2w12, int W, int H, float fx, float fy, int spp,
                      float step_size, float sigma_thresh, float background, uint64_t rng_state,
                      uint64_t rng_inc, int pix_begin, int pix_end, const float* thresh, float* aux,
                      float* img, uint32_t* tr_steps, int32_t* tr_term, uint32_t* tr_src_bits,
                      uint32_t* tr_t_bits, uint64_t* tr_leaf_hash, uint32_t* tr_depth_sum,
                      uint32_t* tr_n_hits, int32_t* tr_hit_leaf, uint32_t* tr_hit_cnt, int32_t* tr_leaf_seq,
                      int max_seq) {
    static const int ok_spp[] = {1, 2, 3, 4, 6, 8, 16, 32};
    int ok = 0;
    for (unsigned i = 0; i < sizeof(ok_spp) / sizeof(int); ++i) ok |= (ok_spp[i] == spp);
    if (!ok) return -1;
    tree_t tree;
    tree.child = child; tree.data = data; tree.data_dim = data_dim; tree.basis_dim = basis_dim;
    for (int i = 0; i < 3; ++i) { tree.offset[i] = offset[i]; tree.scale[i] = scale[i]; }
    tree.ndc_width = ndc_width; tree.ndc_height = ndc_height; tree.ndc_focal = ndc_focal;
    trace_t tr = {tr_steps, tr_term, tr_src_bits, tr_t_bits, tr_leaf_hash, tr_depth_sum, tr_n_hits,
                  tr_hit_leaf, tr_hit_cnt, tr_leaf_seq, max_seq};
    const int has_trace = tr_steps || tr_term || tr_src_bits || tr_t_bits || tr_leaf_hash || tr_depth_sum ||
                          tr_n_hits || tr_hit_leaf || tr_hit_cnt || tr_leaf_seq;
    const float* m = c2w12;
    const size_t SIZE = (size_t)W * H;
    for (int idx = pix_begin; idx < pix_end; ++idx) {
        const int ix = idx % W, iy = idx / W;
        /* screen2worlddir (volrend.cu:24-34) in the PTX's op order */
        const float x = ((float)ix - (float)W * 0.5f) / fx;
        const float y = (-((float)iy - (float)H * 0.5f)) / fy;
        float o[3], dir[3], cen[3], vdir[3];
        for (int k = 0; k < 3; ++k) o[k] = fmaf(x, m[k], y * m[3 + k]) - m[6 + k];
        float inv = 1.0f / sqrtf(fmaf(o[2], o[2], fmaf(o[0], o[0], o[1] * o[1])));
        for (int k = 0; k < 3; ++k) { dir[k] = o[k] * inv; vdir[k] = dir[k]; cen[k] = m[9 + k]; }
        if (ndc_width > 0.f) { /* maybe_world2ndc (volrend.cu:36-56) */
            const float t = (-(cen[2] + 1.0f)) / dir[2];
            for (int k = 0; k < 3; ++k) cen[k] = fmaf(t, dir[k], cen[k]);
            const float k0 = (ndc_focal * -2.0f) / ndc_width;
            const float k1 = (ndc_focal * -2.0f) / ndc_height;
            const float c0 = cen[0] / cen[2], c1 = cen[1] / cen[2];
            const float nd0 = k0 * (dir[0] / dir[2] - c0);
            const float nd1 = k1 * (dir[1] / dir[2] - c1);
            const float nd2 = -2.0f / cen[2];
            const float nc2 = 2.0f / cen[2] + 1.0f;
            cen[0] = k0 * c0; cen[1] = k1 * c1; cen[2] = nc2;
            const float n = 1.0f / sqrtf(fmaf(nd2, nd2, fmaf(nd0, nd0, nd1 * nd1)));
            dir[0] = nd0 * n; dir[1] = nd1 * n; dir[2] = nd2 * n;
        }
        for (int k = 0; k < 3; ++k) cen[k] = fmaf(tree.scale[k], cen[k], tree.offset[k]); /* :142-144 */

        float dst[RTO_MAX_SPP + 1];
        if (thresh) {
            memcpy(dst, thresh + (size_t)(idx - pix_begin) * spp, sizeof(float) * spp);
            dst[spp] = FLT_MAX;
        } else {
            pcg32_t rng = {rng_state, rng_inc};
            pcg32_advance(&rng, (uint64_t)(int64_t)(idx * spp)); /* volrend.cu:157 (int product) */
            sample_dst_cpu(&rng, spp, dst);
        }
        float out[4] = {0.f, 0.f, 0.f, 0.f};
        trace_ray(&tree, dir, vdir, cen, spp, step_size, sigma_thresh, dst, out, has_trace ? &tr : NULL,
                  (size_t)(idx - pix_begin));
        /* background composite, offscreen (volrend.cu:174-179) */
        const float remain = background * (1.f - out[3]);
        out[0] += remain; out[1] += remain; out[2] += remain;
        if (aux) { /* volrend.cu:187-202 */
            aux[idx] = out[0];
            aux[idx + SIZE] = out[1];
            aux[idx + 2 * SIZE] = out[2];
            aux[idx + 3 * SIZE] = out[3];
            aux[idx + 4 * SIZE] = out[0] * out[0];
            aux[idx + 5 * SIZE] = out[1] * out[1];
            aux[idx + 6 * SIZE] = out[2] * out[2];
            aux[idx + 7 * SIZE] = out[3] * out[3];
        }
        if (img) { /* volrend.cu:205-212 */
            img[4 * (size_t)idx + 0] = out[0];
            img[4 * (size_t)idx + 1] = out[1];
            img[4 * (size_t)idx + 2] = out[2];
            img[4 * (size_t)idx + 3] = 1.0f;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------ */
/* GuidanceNet, deployed form (network.py:152-168, graph of Appendix B):                            */
/*   x = half(aux); for each layer: x = relu6(half(half(conv3x3_same(x, W)) + b)); y = float(x);    */
/*   weight = softmax(y[:L]); guidance = y[L:]                                                       */
/* conv accumulates in fp32 over (ci, ky, kx) in that loop order (library-internal order is unpinned);*/
/* fp16 rounding points: fused_bias = 0 follows ATen's cuDNN path (cudnn_convolution rounds to fp16,  */
/* then output.add_(bias) rounds again) = what the reference deploys on the GPU; fused_bias = 1 rounds*/
/* once after adding the bias in fp32 = what PyTorch's CPU fp16 conv does (pinned by                 */
/* tests/golden/guidance_net_ref.npz, generated from the reference module on CPU).                   */
static void conv3x3_f16(const float* in /*[Ci][H][W] fp16-valued*/, int Ci, int Co, int H, int W,
                        const uint16_t* w /*[Co][Ci][3][3]*/, const uint16_t* b, float* out /*[Co][H][W]*/,
                        int fused_bias) {
    float* wf = (float*)malloc(sizeof(float) * (size_t)Co * Ci * 9);
    for (int i = 0; i < Co * Ci * 9; ++i) wf[i] = h2f(w[i]);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            for (int co = 0; co < Co; ++co) {
                float acc = 0.f;
                for (int ci = 0; ci < Ci; ++ci)
                    for (int ky = 0; ky < 3; ++ky) {
                        const int yy = y + ky - 1;
                        if (yy < 0 || yy >= H) continue;
                        for (int kx = 0; kx < 3; ++kx) {
                            const int xx = x + kx - 1;
                            if (xx < 0 || xx >= W) continue;
                            acc = fmaf(in[((size_t)ci * H + yy) * W + xx], wf[((co * Ci + ci) * 3 + ky) * 3 + kx], acc);
                        }
                    }
                float v = fused_bias ? round_f16(acc + h2f(b[co])) : round_f16(round_f16(acc) + h2f(b[co]));
                v = fminf(fmaxf(v, 0.f), 6.f);
                out[((size_t)co * H + y) * W + x] = v;
            }
    free(wf);
}

/* aux [in_ch][H][W] fp32 -> weight [L][H][W], guidance [L][H][W] fp32 ; two layers (num_layers = 2)   */
int rto_oracle_guidance_net(const float* aux, int in_ch, int mid_ch, int L, int H, int W,
                            const uint16_t* w1, const uint16_t* b1, const uint16_t* w2, const uint16_t* b2,
                            float* weight, float* guidance, int fused_bias) {
    const size_t HW = (size_t)H * W;
    float* x0 = (float*)malloc(sizeof(float) * in_ch * HW);
    float* x1 = (float*)malloc(sizeof(float) * mid_ch * HW);
    float* x2 = (float*)malloc(sizeof(float) * 2 * L * HW);
    if (!x0 || !x1 || !x2) return -1;
    for (size_t i = 0; i < in_ch * HW; ++i) x0[i] = round_f16(aux[i]);
    conv3x3_f16(x0, in_ch, mid_ch, H, W, w1, b1, x1, fused_bias);
    conv3x3_f16(x1, mid_ch, 2 * L, H, W, w2, b2, x2, fused_bias);
    for (size_t p = 0; p < HW; ++p) {
        float mx = -FLT_MAX, sum = 0.f, e[16];
        for (int l = 0; l < L; ++l) mx = fmaxf(mx, x2[l * HW + p]);
        for (int l = 0; l < L; ++l) { e[l] = expf(x2[l * HW + p] - mx); sum += e[l]; }
        for (int l = 0; l < L; ++l) weight[l * HW + p] = e[l] / sum;
        for (int l = 0; l < L; ++l) guidance[l * HW + p] = x2[(L + l) * HW + p];
    }
    free(x0); free(x1); free(x2);
    return 0;
}

/* filtering forward (filtering.cu:108-228 per level, :441-470 over levels, support = level+1)        */
/* img_in [H][W][4], weight/guidance [L][H][W], img_out [H][W][4]                                    */
int rto_oracle_filter(const float* img_in, const float* weight, const float* guidance, int L, int H, int W,
                      float* img_out) {
    if (L < 1 || L > 6) return -1; /* filtering.cu:338-367: supports 1..6 */
    const size_t HW = (size_t)H * W;
    for (int level = 0; level < L; ++level) {
        const int S = level + 1;
        const float* g = guidance + level * HW;
        const float* wm = weight + level * HW;
#pragma omp parallel for schedule(static)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                float mx = -FLT_MAX;
                for (int dy = -S; dy <= S; ++dy)
                    for (int dx = -S; dx <= S; ++dx) {
                        const int yy = y + dy, xx = x + dx;
                        const float gv = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? g[(size_t)yy * W + xx] : -FLT_MAX;
                        mx = fmaxf(mx, gv);
                    }
                float r = 0.f, gg = 0.f, b = 0.f, ksum = 0.f;
                for (int dy = -S; dy <= S; ++dy)
                    for (int dx = -S; dx <= S; ++dx) {
                        const int yy = y + dy, xx = x + dx;
                        if (!(yy >= 0 && yy < H && xx >= 0 && xx < W)) continue; /* k = exp(-FLT_MAX-mx) = 0 */
                        const float k = exp2f((g[(size_t)yy * W + xx] - mx) * 1.442695f);
                        ksum += k;
                        const float* q = img_in + 4 * ((size_t)yy * W + xx);
                        r = fmaf(q[0], k, r); gg = fmaf(q[1], k, gg); b = fmaf(q[2], k, b);
                    }
                const float w = wm[(size_t)y * W + x] * (1.0f / ksum);
                float* o = img_out + 4 * ((size_t)y * W + x);
                if (level == 0) { o[0] = r * w; o[1] = gg * w; o[2] = b * w; o[3] = 1.0f; }
                else { o[0] += r * w; o[1] += gg * w; o[2] += b * w; }
            }
    }
    return 0;
}
