/*
 * rtoctree_b200.h — C ABI of librtoctree_b200.so, the B200-native (sm_100a) RT-Octree render hot path.
 *
 * The reference (LumiOwO/RT-Octree) has no FFI layer: its boundary is the C++ API of two static libraries
 * (`volrend`, `volrend_denoiser`) plus the `volrend_headless` CLI and file formats.  This header
 * is the thin C layer the north star asks for; every entry point names the reference interface it replaces.
 * The C++ classes in rt_octree_b200/host/volrend_b200.hpp mirror the reference classes one-to-one on top of it,
 * and rt_octree_b200/capi.py binds it with ctypes for the tests and bench.py.  See INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every function returns RTO_OK (0) or a negative rto_status and
 * leaves a message for rto_last_error() (thread-local).  `stream` is a cudaStream_t passed as void* (NULL = the
 * legacy default stream); all render/denoise calls are asynchronous on it, exactly like launch_renderer /
 * Denoiser::denoise.  There is NO CPU fallback: without a usable CUDA device every call that needs one fails
 * with RTO_ERR_CUDA.
 */
#ifndef RTOCTREE_B200_H_
#define RTOCTREE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTO_ABI_VERSION 2

typedef enum rto_status {
    RTO_OK = 0,
    RTO_ERR_INVALID = -1,      /* bad argument / malformed tree (the reference throws std::runtime_error) */
    RTO_ERR_UNSUPPORTED = -2,  /* e.g. spp not in {1,2,3,4,6,8,16,32}: renderer/src/cuda/volrend.cu:266-278 */
    RTO_ERR_CUDA = -3,         /* CUDA runtime error (the reference prints, cudaDeviceReset()s and exits:
                                  renderer/src/cuda/common.cu:8-21; a library must not exit) */
    RTO_ERR_NOMEM = -4
} rto_status;

/* DataFormat::format, renderer/include/volrend/data_format.hpp:9-16 */
typedef enum rto_data_format { RTO_FORMAT_RGBA = 0, RTO_FORMAT_SH = 1, RTO_FORMAT_SG = 2, RTO_FORMAT_ASG = 3 } rto_data_format;

typedef struct rto_tree rto_tree;       /* replaces volrend::N3Tree (device side)   include/volrend/n3tree.hpp:24-106 */
typedef struct rto_context rto_context; /* replaces volrend::RenderContext          include/volrend/render_context.hpp:14-120 */
typedef struct rto_net rto_net;         /* replaces volrend::Denoiser               include/volrend/denoiser/denoiser.hpp:11-21 */

/* Camera + CameraSpec (include/volrend/camera.hpp:16-68, internal/data_spec.hpp:11-26).  c2w is the column-major
 * 4x3 camera-to-world transform Camera::_update uploads (src/camera.cpp:72-73): right, up, back, centre. */
typedef struct rto_camera {
    int width, height;
    float fx, fy;
    float c2w[12];
} rto_camera;

/* RenderOptions (include/volrend/render_options.hpp:13-78).  The JSON binding carries 11 keys (:61-77); keys that
 * only drive the GUI (show_grid, grid_max_depth, probe*) are parsed by the host layer and are not needed here.
 * enable_probe must be 0 (lumisphere probe is out of scope). */
typedef struct rto_render_options {
    float step_size;              /* 1e-4 */
    float sigma_thresh;           /* 1e-2 */
    float stop_thresh;            /* parsed, unused by the CUDA path (only shaders/rt.frag:314) */
    float background_brightness;  /* 1.0 */
    int denoise;                  /* 1: final image is produced by rto_denoise, rto_render writes only aux */
    int spp;                      /* 1,2,3,4,6,8,16,32 */
    int enable_probe;             /* must be 0 */
} rto_render_options;
void rto_render_options_default(rto_render_options* opt); /* reference defaults, spp = 1, denoise = 1 */

typedef struct rto_tree_info {
    int64_t capacity;        /* nodes */
    int N, data_dim, format, basis_dim;
    int max_depth;           /* max child look-ups to reach a leaf */
    int64_t n_leaves;
    int64_t node_bytes, payload_bytes; /* HBM footprint of the SoA layout */
    int payload_stride_halfs;
    int grid_level;          /* K of the sparse brick grid used by the marching loop (0 = none) */
    int64_t n_bricks, grid_bytes;
    float offset[3], scale[3];
    float ndc_width, ndc_height, ndc_focal;
} rto_tree_info;

/* Per-ray traversal record for the bit-exact parity tests (all DEVICE pointers, any may be NULL), indexed by
 * the full-frame pixel index iy*W+ix.  Definitions in oracle/rt_oracle.c (trace_t). */
typedef struct rto_trace {
    uint32_t* steps;      /* leaf visits */
    int32_t* term;        /* step index of the SPP-th collision, or -1 */
    uint32_t* src_bits;   /* fp32 bits of the accumulated optical depth at exit */
    uint32_t* t_bits;     /* fp32 bits of t at exit */
    uint64_t* leaf_hash;  /* FNV-1a 64 over visited leaf indices */
    uint32_t* depth_sum;  /* sum of child look-ups the REFERENCE's root-restart query would do (sum of leaf depths) */
    uint32_t* n_hits;     /* collided-leaf entries (sh_nums) */
    uint32_t* n_loads;    /* node words this implementation actually loaded (ancestor-resume descent) */
    int32_t* hit_leaf;    /* [n][spp] */
    uint32_t* hit_cnt;    /* [n][spp] */
    int32_t* leaf_seq;    /* [n][max_seq] */
    float* thresh;        /* [n][spp] sorted thresholds dst[] (lg2.approx based) */
    int max_seq;
    int marcher;          /* which marching loop writes the record: 0 = tree walker (ancestor-stack descent),
                             1 = the PRODUCTION brick-grid marcher (the loop rto_render runs; the leaf visited at every step
                             is located through the tree as well, which also cross-checks the grid's depth and sigma: a
                             disagreement is reported as term = -777).  Trees without a brick grid (depth < 4 or > 11) are
                             marched by the tree walker in both modes, exactly like rto_render does. */
} rto_trace;

const char* rto_last_error(void);
int rto_abi_version(void);
/* cudaSetDevice — main_headless.cpp:234-238 (`--gpu`).  */
int rto_set_device(int device);
int rto_device_count(int* count);
/* cudaStreamSynchronize(stream) (stream == NULL: cudaDeviceSynchronize) — what Timer::record / cudaMemcpy do implicitly */
int rto_synchronize(void* stream);

/* ---- tree : N3Tree::load_npz result -> N3Tree::load_cuda (src/n3tree.cpp:228-362, src/cuda/n3tree.cu:9-41) ----
 * Host arrays exactly as they sit in tree.npz: child int32 [capacity][N][N][N] (relative node offsets, 0 = leaf),
 * data fp16 [capacity][N][N][N][data_dim] (last = sigma).  The call uploads them and builds every HBM plane ON THE GPU:
 * node words with embedded sigma, the padded fp16 payload plane, the structure check (offsets in range, no cycles,
 * depth) and the sparse brick grid the marching loop reads.  N must be 2. */
int rto_tree_create(rto_tree** out, const int32_t* child, const void* data_f16, int64_t capacity, int N,
                    int data_dim, int format, int basis_dim, const float offset[3], const float scale[3]);
/* Quantised / svox-compressed tree.npz (producer renderer/scripts/compress_octree.py:68-119; the reference decodes it on the
 * host into the dense `data` array, renderer/src/n3tree.cpp:279-340).  Here the compressed arrays are uploaded as they sit
 * in the file and gathered ON THE GPU straight into the payload plane:
 *   quant_colors fp16 [n_quant][65536][3], quant_map u16 [n_quant][capacity][N][N][N], sigma fp16 [capacity][N][N][N],
 *   data_retained fp16 [n_retained][capacity][N][N][N][3] (NULL when n_retained == 0);
 *   colour slot j + n_retained + k*(n_quant+n_retained) of a leaf = quant_colors[j][quant_map[j][leaf]][k]   (:301-316),
 *   colour slot j + k*(n_quant+n_retained)               of a leaf = data_retained[j][leaf][k]                (:325-338). */
int rto_tree_create_quantized(rto_tree** out, const int32_t* child, int64_t capacity, int N, int data_dim, int format,
                              int basis_dim, const float offset[3], const float scale[3], const void* quant_colors_f16,
                              const uint16_t* quant_map, int n_quant, const void* sigma_f16, const void* data_retained_f16,
                              int n_retained);
/* Copy one device plane of the loaded tree back to the host (inspection / parity tests; the reference keeps the host
 * arrays in N3Tree::data_ / child_ instead).  `bytes` must equal the plane size: nodes = rto_tree_info.node_bytes,
 * payload = payload_bytes, grid top / leaf top / march top = 4 << (3*grid_level), grid bricks / leaf bricks = n_bricks * 2048,
 * byte bricks = n_bricks * 512 (the leaf planes hold 0 bytes when they were not built). */
typedef enum rto_tree_plane { RTO_PLANE_NODES = 0, RTO_PLANE_PAYLOAD = 1, RTO_PLANE_GRID_TOP = 2, RTO_PLANE_GRID_BRICKS = 3,
                              RTO_PLANE_GRID_BRICKS8 = 4,
                              RTO_PLANE_GRID_LEAF_TOP = 5,    /* u32 per level-K cell: flat leaf index (node*8+octant) of a leaf cell */
                              RTO_PLANE_GRID_LEAF_BRICKS = 6, /* u32 per brick cell: flat leaf index of the covering leaf */
                              RTO_PLANE_GRID_MARCH_TOP = 7    /* u32 per level-K cell: leaf word, or brick*512 + 2^18 - (512x + 64y + 8z) */
} rto_tree_plane;
int rto_tree_read_plane(const rto_tree* tree, int plane, void* host_dst, size_t bytes);
/* main_headless.cpp:400-405 / n3tree.hpp:69-71: NDC warp for forward-facing (llff) scenes; width<=0 disables. */
int rto_tree_set_ndc(rto_tree* tree, float ndc_width, float ndc_height, float ndc_focal);
int rto_tree_get_info(const rto_tree* tree, rto_tree_info* info);
void rto_tree_destroy(rto_tree* tree); /* N3Tree::~N3Tree -> free_cuda */

/* ---- context : RenderContext::update / freeResource (render_context.hpp:42-120) ----
 * Owns aux [8][H][W] fp32, the output image [H][W][4] fp32 (linear memory instead of a cudaArray surface), the
 * GuidanceNet scratch maps, the pcg32 state (seed 20230418, :16) and the three-stage event timer (:122-213). */
int rto_context_create(rto_context** out, int width, int height);
void rto_context_destroy(rto_context* ctx);
float* rto_context_aux(rto_context* ctx);    /* device pointer, RenderContext::aux_buffer */
float* rto_context_image(rto_context* ctx);  /* device pointer, float4 per pixel */
/* ctx.rng: pcg32(seed) ; advance(delta) with the reference default delta = 2^32 (pcg32.h:145) */
int rto_context_rng_seed(rto_context* ctx, uint64_t seed);
int rto_context_rng_advance(rto_context* ctx, int64_t delta);
/* Pure function of the frame index: state = pcg32(20230418) advanced by (warmup + frame) * 2^32, which is what
 * main_headless.cpp:469-479,506 leaves in ctx.rng when it renders pose `frame` (warmup = 100 there).  Lets any
 * rank of a frame-sharded job reproduce the single-GPU image bit for bit. */
int rto_context_rng_set_frame(rto_context* ctx, int64_t warmup, int64_t frame);
int rto_context_rng_get(const rto_context* ctx, uint64_t* state, uint64_t* inc);
/* cudaMemcpy(…, DeviceToHost) of aux (main_headless.cpp:516-517) / image (:526-534); async on `stream`. */
int rto_context_read_aux(rto_context* ctx, float* host_dst, void* stream);
/* upload a stored guidance buffer (`buf_<name>.bin` of --write_buffer, fp32 [8][H][W]; denoiser/dataset.py:161-163 reads the
 * same layout) so that rto_denoise can run on it without rendering */
int rto_context_write_aux(rto_context* ctx, const float* host_src, void* stream);
int rto_context_read_image(rto_context* ctx, float* host_dst, void* stream);
/* the same image as RGBA8 [H][W][4], converted on the device exactly like the reference CLI converts on the host before
 * writing a PNG, `(uint8_t)(v * 255)` (main_headless.cpp:524-541): a quarter of the device->host bytes */
int rto_context_read_image_rgba8(rto_context* ctx, unsigned char* host_dst, void* stream);
/* device pointer of that RGBA8 copy ([H][W][4] bytes; allocated by the first call).  Once it exists, the kernels that produce
 * the image (rto_render with denoise off, rto_denoise's filter) write it from the same epilogue, so the read-back above
 * costs no extra launch. */
unsigned char* rto_context_image_rgba8(rto_context* ctx);
/* rows [y0, y1) of that RGBA8 copy into the SAME rows of a full-frame host buffer ([H][W][4] u8): the read-back of one band of a
 * tile split whose GPUs each deliver their own rows to the host (N PCIe links in parallel) instead of assembling the frame on
 * one GPU first.  The rows must have been produced on this context (rto_render_rect with denoise off, or rto_denoise_rows)
 * after rto_context_image_rgba8 created the copy, with no image target set. */
int rto_context_read_rows_rgba8(rto_context* ctx, unsigned char* host_frame, int y0, int y1, void* stream);
int rto_context_read_image_rows(rto_context* ctx, float* host_frame, int y0, int y1, void* stream);   /* the same for the float4 image */

/* ---- render : volrend::launch_renderer(tree, cam, options, ctx, stream, offscreen=true)
 *               include/volrend/cuda/renderer_kernel.hpp:11-16, src/cuda/volrend.cu:236-285 ----
 * CONTRACT (multi-stream callers): a context carries the work counter of the persistent render kernel and the frame's
 * buffers, so at most ONE render / denoise may be in flight per context; pipelined callers use one context per stream
 * (the reference has a single RenderContext and a single blocking stream, main_headless.cpp:441-447).  Trees and nets are
 * read-only and may be shared by any number of contexts, streams and host threads of the same device.
 * SIDE EFFECT: the first render of a tree with a brick grid raises the DEVICE limit cudaLimitPersistingL2CacheSize to the
 * size of the plane the marching loop reads (16 MB for the bench tree; never lowered) so that the per-frame streaming
 * buffers cannot evict it from L2; RTO_L2_PERSIST=0 in the environment disables this. */
int rto_render(rto_context* ctx, const rto_tree* tree, const rto_camera* cam, const rto_render_options* opt,
               void* stream);
/* Same, restricted to the pixel rectangle [x0,x1) x [y0,y1) (single-frame tile split); pixel
 * indices, RNG offsets and buffer addresses stay full-frame. */
int rto_render_rect(rto_context* ctx, const rto_tree* tree, const rto_camera* cam, const rto_render_options* opt,
                    int x0, int y0, int x1, int y1, void* stream);
/* Same kernel with the per-ray traversal record switched on (parity tests only); trace->marcher selects the loop. */
int rto_render_trace(rto_context* ctx, const rto_tree* tree, const rto_camera* cam, const rto_render_options* opt,
                     const rto_trace* trace, void* stream);

/* ---- denoiser : Denoiser(ts_module_path) / Denoiser::denoise(cam, ctx, stream)  (src/denoiser/denoiser.cpp:8-71)
 * Weights are the four fp16 tensors of the deployed GuidanceNet (network.py:123-168), exported once from the
 * reference's ts_*.ts by tools/make_ts_module.py --export: w1 [mid][in][3][3], b1 [mid], w2 [2L][mid][3][3], b2 [2L]. */
int rto_net_create(rto_net** out, const void* w1_f16, const void* b1_f16, const void* w2_f16, const void* b2_f16,
                   int in_ch, int mid_ch, int levels);
void rto_net_destroy(rto_net* net);
/* implementation selector: 0 = auto (tensor-core kernel when the net is the shipped 8->32->8, L=4 shape),
 * 1 = force the generic CUDA-core kernels (bring-up / cross-check path) */
int rto_net_set_impl(rto_net* net, int impl);
/* fp16 rounding of the conv bias add: 0 (default) = half(half(acc)+b), ATen's cuDNN path (cudnn_convolution, then
 * output.add_(bias)) that the reference runs on the GPU; 1 = half(acc+b), what PyTorch's CPU fp16 conv computes. */
int rto_net_set_bias_mode(rto_net* net, int fused);
int rto_denoise(rto_context* ctx, const rto_net* net, void* stream);
int rto_denoise_rows(rto_context* ctx, const rto_net* net, int y0, int y1, void* stream);
/* The two halves on their own, on caller-provided DEVICE buffers:
 * ts_module.forward(aux) -> (weight_map [L][H][W], guidance_map [L][H][W])           denoiser.cpp:46-48 */
int rto_net_forward(const rto_net* net, const float* aux_dev, int width, int height, float* weight_dev,
                    float* guidance_dev, void* stream);
/* denoiser::filtering(stream, weight_map, guidance_map, img_in, img_out)  denoiser/extension/filtering.h:7-13
 * img_in / img_out: [H][W][4] fp32 device buffers; L in 1..6 (filtering.cu:338-367). */
int rto_filter(const float* weight_dev, const float* guidance_dev, const float* img_in_dev, int levels, int width,
               int height, float* img_out_dev, void* stream);

/* ---- training side of the same op: Filtering::forward with requires_grad / Filtering::backward,
 * denoiser/extension/filtering.cu:580-699 behind `_denoiser.filtering_autograd` (bindings.cpp:5-13).  One image per call
 * (the reference loops over the batch on the host too).  rgb_filtered [L][H][W][4], max_map / inv_kernel_sum [L][H][W] are
 * the per-level tensors the reference saves for backward (:207-217); grads are [L][H][W].  The backward GATHERS over the
 * symmetric window instead of the reference's scatter + atomicAdd (:230-301): same sums, deterministic order. */
int rto_filter_forward_save(const float* weight_dev, const float* guidance_dev, const float* img_in_dev, int levels, int width,
                            int height, float* img_out_dev, float* rgb_filtered_dev, float* max_map_dev,
                            float* inv_kernel_sum_dev, void* stream);
int rto_filter_backward(const float* grad_output_dev, const float* img_in_dev, const float* weight_dev, const float* guidance_dev,
                        const float* rgb_filtered_dev, const float* max_map_dev, const float* inv_kernel_sum_dev, int levels,
                        int width, int height, float* grad_weight_dev, float* grad_guidance_dev, void* stream);

/* ---- single-frame tile split over several GPUs (no counterpart in the reference, which is single-GPU) ----
 * Each GPU renders a row band (+ halo: rto_render_rect) and denoises it (rto_denoise_rows); instead of gathering the bands
 * afterwards, the kernels that PRODUCE the final image — the filter epilogue, or the render kernel when the denoiser is off —
 * store their rows straight into one destination image, which may live in another GPU's memory: peer-direct stores over
 * NVLink / NVSwitch, fused with the last kernel of the frame.  `image_dev` ([H][W][4] fp32) and the optional `rgba8_dev`
 * ([H][W][4] u8) are full-frame device pointers: another context's rto_context_image / rto_context_image_rgba8 in the same
 * process (after rto_peer_enable), or a pointer obtained with rto_ipc_open in another process.  NULL restores the
 * context's own buffers.  The caller orders the bands' completion before reading the destination (events / barrier), then
 * calls rto_context_mark_image_written on the destination context so that its RGBA8 copy is known to be current. */
int rto_context_set_image_target(rto_context* ctx, float* image_dev, unsigned char* rgba8_dev);
int rto_context_mark_image_written(rto_context* ctx, int rgba8_too);
int rto_peer_enable(int peer_device);                                  /* cudaDeviceEnablePeerAccess from the current device */
int rto_ipc_export(const void* dev_ptr, unsigned char handle[64]);    /* cudaIpcGetMemHandle (dev_ptr: base of an allocation) */
int rto_ipc_open(const unsigned char handle[64], void** dev_ptr);     /* cudaIpcOpenMemHandle, peer access enabled lazily */
int rto_ipc_close(void* dev_ptr);
int rto_event_create(void** event);                                    /* cross-stream / cross-device ordering without host syncs */
int rto_event_create_timed(void** event);                              /* with timing: per-band device times for the band balancer */
int rto_event_elapsed_ms(void* start, void* end, float* ms);           /* waits for `end`, then cudaEventElapsedTime */
int rto_event_record(void* event, void* stream);
int rto_stream_wait_event(void* stream, void* event);
int rto_event_destroy(void* event);

/* ---- pipelined callers (no counterpart in the reference, whose driver owns one blocking stream, main_headless.cpp:445) ----
 * Non-blocking streams and pinned host memory without linking the CUDA runtime into the host program. */
int rto_stream_create(void** stream);
int rto_stream_destroy(void* stream);
int rto_host_alloc(void** ptr, size_t bytes);   /* cudaHostAlloc: destinations of asynchronous read-backs */
int rto_host_free(void* ptr);

/* One frame of the path — launch_renderer, Denoiser::denoise and the read-backs of main_headless.cpp:506-541 — captured once
 * as a CUDA graph on `ctx` and replayed with ONE launch per frame.  Per frame only the camera transform and ctx.rng (the state
 * rto_context_rng_* left at the time of the call) change; tree, net, options, focal lengths and the optional PINNED host
 * destinations (rto_host_alloc; NULL = no copy) are fixed at creation.  Same kernels, same results as the separate calls. */
typedef struct rto_frame rto_frame;
typedef struct rto_frame_desc {
    const rto_tree* tree;
    const rto_net* net;            /* may be NULL when opt.denoise == 0 */
    rto_render_options opt;
    float fx, fy;
    unsigned char* host_rgba8;     /* [H][W][4] u8   (rto_context_read_image_rgba8) */
    float* host_image;             /* [H][W][4] f32  (rto_context_read_image) */
    float* host_aux;               /* [8][H][W] f32  (rto_context_read_aux, the --write_buffer copy) */
} rto_frame_desc;
int rto_frame_create(rto_frame** out, rto_context* ctx, const rto_frame_desc* desc);
int rto_frame_launch(rto_frame* frame, const float c2w[12], void* stream);
/* rto_context_rng_set_frame(ctx, warmup, frame_index) + rto_frame_launch in one call: the whole per-frame host work of a
 * frame-sharded or pipelined driver (pose `frame_index` of the job, rng state as main_headless.cpp leaves it for that pose) */
int rto_frame_launch_indexed(rto_frame* frame, const float c2w[12], int64_t warmup, int64_t frame_index, void* stream);
/* The host loop of a pipelined driver in ONE call (volrend_headless --pipe N; main_headless.cpp:441-541 is the serial
 * original): frame i of [first, first + count) runs on slot k = i % n_slots, i.e. on frames[k] / streams[k], with pose
 * c2w + 12 * (i % n_poses) and the rng state of pose i (rto_frame_launch_indexed).  Before a slot is reused the call waits
 * for its stream — the slot's previous frame has reached its pinned host destinations — and, if `retired` is given,
 * reports that frame as retired(user, frame_index, slot); the callback is the place to consume the host buffers and must
 * return only when they may be overwritten.  Frames still in flight when the call returns are reported by the next call on
 * the same slots, or waited for and reported by this one with drain != 0.  Returns the first error; nothing else is launched
 * after it. */
typedef void (*rto_frame_retired_fn)(void* user, int64_t frame_index, int slot);
int rto_frame_sequence(rto_frame* const* frames, void* const* streams, int n_slots, const float* c2w, int64_t n_poses,
                       int64_t warmup, int64_t first, int64_t count, int drain, rto_frame_retired_fn retired, void* user);
void rto_frame_destroy(rto_frame* frame);

/* ---- timer : RenderContext::Timer (render_context.hpp:122-213) ----
 * With timing enabled rto_render / rto_denoise bracket their launches with cudaEvents on `stream`;
 * rto_timer_record synchronises on the last stop event and accumulates (Timer::record).  ms[0..2] = mean
 * render / net / filter ms per recorded frame (Timer::report); FPS = 1000 / (ms[0]+ms[1]+ms[2]). */
int rto_timer_enable(rto_context* ctx, int enable);
int rto_timer_reset(rto_context* ctx);
int rto_timer_record(rto_context* ctx, int denoise);
int rto_timer_report(const rto_context* ctx, float ms[3], int* frames);

/* number of kernels this library has launched since load (bench.py's gpu_launches evidence) */
int64_t rto_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RTOCTREE_B200_H_ */
