// volrend_b200.hpp — C++ host-side mirror of the reference's public classes for the headless render path, implemented
// on top of the C ABI (include/rtoctree_b200.h).  Same names, argument meaning and error behaviour as
//   volrend::N3Tree            renderer/include/volrend/n3tree.hpp:24-106      (+ src/n3tree.cpp:111-362)
//   volrend::Camera            renderer/include/volrend/camera.hpp:16-68
//   volrend::RenderOptions     renderer/include/volrend/render_options.hpp:13-78
//   volrend::RenderContext     renderer/include/volrend/render_context.hpp:14-213 (incl. Timer)
//   volrend::launch_renderer   renderer/include/volrend/cuda/renderer_kernel.hpp:11-16
//   volrend::Denoiser          renderer/include/volrend/denoiser/denoiser.hpp:11-21
// so that a driver written against the reference (e.g. main_headless.cpp) ports by swapping the include.  No glm, no
// libtorch, no OpenGL: Camera::transform is a plain column-major float[12]; streams are passed as void*.
#pragma once
#include <array>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rtoctree_b200.h"
#include "json_min.hpp"
#include "npz.hpp"

namespace volrend {

inline void rto_check(int rc, const char* what) {
    if (rc == RTO_OK) return;
    std::string msg = std::string(what) + ": " + rto_last_error();
    if (rc == RTO_ERR_CUDA) {   // the reference's cuda(...) macro prints and exits (src/cuda/common.cu:8-21)
        fprintf(stderr, "%s\n", msg.c_str());
        std::exit(rc);
    }
    throw std::runtime_error(msg);
}

// ---------------------------------------------------------------------------------------------- DataFormat
struct DataFormat {
    enum { RGBA, SH, SG, ASG, _COUNT } format = RGBA;
    int basis_dim = -1;
    void parse(const std::string& str) {   // src/n3tree.cpp:55-78
        size_t nonalph = std::string::npos;
        for (size_t i = 0; i < str.size(); ++i)
            if (!std::isalpha((unsigned char)str[i])) { nonalph = i; break; }
        if (nonalph != std::string::npos) {
            basis_dim = std::atoi(str.c_str() + nonalph);
            const std::string t = str.substr(0, nonalph);
            format = t == "ASG" ? ASG : t == "SG" ? SG : t == "SH" ? SH : RGBA;
        } else {
            basis_dim = -1;
            format = RGBA;
        }
    }
    std::string to_string() const {
        static const char* names[] = {"RGBA", "SH", "SG", "ASG"};
        std::string out = format < _COUNT ? names[format] : "UNKNOWN";
        if (~basis_dim) out.append(std::to_string(basis_dim));
        return out;
    }
};

// -------------------------------------------------------------------------------------------------- N3Tree
struct N3Tree {
    N3Tree() = default;
    explicit N3Tree(const std::string& path) { open(path); }
    N3Tree(const N3Tree&) = delete;
    N3Tree& operator=(const N3Tree&) = delete;
    ~N3Tree() { rto_tree_destroy(device); }

    int N = 0;
    int data_dim = 0;
    DataFormat data_format;
    int capacity = 0;
    std::array<float, 3> scale{}, offset{};
    bool use_ndc = false;
    float ndc_width = 0, ndc_height = 0, ndc_focal = 0;
    mutable rto_tree* device = nullptr;   // replaces device.{data,child,offset,scale}: SoA layout owned by the library

    bool is_data_loaded() const { return device != nullptr; }

    // N3Tree::open -> cnpy::npz_load -> load_npz -> load_cuda   (src/n3tree.cpp:111-154, 228-362; src/cuda/n3tree.cu:9-41)
    void open(const std::string& path) {
        rto_tree_destroy(device);
        device = nullptr;
        if (!std::ifstream(path)) {
            printf("Can't load because file does not exist: %s\n", path.c_str());
            return;
        }
        rtohost::npz_t npz = rtohost::npz_load(path);
        load_npz(npz);
        const std::string pb = path.substr(0, path.size() - 4) + "_poses_bounds.npy";
        use_ndc = bool(std::ifstream(pb));
        if (use_ndc) {   // src/n3tree.cpp:131-148
            fprintf(stderr, "INFO: Found poses_bounds.npy for NDC: %s\n", pb.c_str());
            rtohost::NpyArray a = rtohost::npy_load(pb);
            auto get = [&](size_t i) { return a.word_size == 4 ? (double)a.data<float>()[i] : a.data<double>()[i]; };
            ndc_height = (float)get(4);
            ndc_width = (float)get(9);
            ndc_focal = (float)get(14);
        }
    }

    // N3Tree::open_mem (n3tree.hpp:32, src/n3tree.cpp:156-172): the same file as a memory image ("for web mostly" there)
    void open_mem(const char* data, uint64_t size) {
        rto_tree_destroy(device);
        device = nullptr;
        rtohost::npz_t npz = rtohost::npz_load_mem(reinterpret_cast<const unsigned char*>(data), (size_t)size);
        load_npz(npz);
        use_ndc = false;
    }

    // call after changing use_ndc / ndc_* (main_headless.cpp:400-405 sets them after construction)
    void sync_ndc() const {
        if (device) rto_check(rto_tree_set_ndc(device, use_ndc ? ndc_width : -1.f, ndc_height, ndc_focal), "rto_tree_set_ndc");
    }

    // Everything load_npz derives from the file, before the upload (also used by `volrend_headless --dry_run`).
    struct HostArrays {
        int N = 0, data_dim = 0, capacity = 0;
        DataFormat data_format;
        std::array<float, 3> scale{}, offset{};
        const int32_t* child = nullptr;
        const unsigned char* data = nullptr;      // fp16 [cap][N][N][N][data_dim]
        size_t n_child = 0;
        std::vector<unsigned char> decoded;       // owns `data` for quantised files decoded on the host (--dry_run)
        // quantised files (src/n3tree.cpp:279-340): the compressed arrays as they sit in the npz
        bool quantized = false;
        const uint16_t *quant_colors = nullptr, *quant_map = nullptr, *sigma = nullptr, *retained = nullptr;
        int n_quant = 0, n_retained = 0;
    };

    // N3Tree::load_npz (src/n3tree.cpp:228-362) without the upload.  Quantised files: the load path hands the compressed
    // arrays to rto_tree_create_quantized (codebook gather on the GPU); `decode_on_host` materialises the dense `data`
    // array on the host instead, for `volrend_headless --dry_run`, which has to work without a device.
    static void decode_npz(rtohost::npz_t& npz, HostArrays& h, bool decode_on_host) {
        auto need = [&](const char* k) -> rtohost::NpyArray& {
            auto it = npz.find(k);
            if (it == npz.end()) throw std::runtime_error(std::string("tree.npz: missing key '") + k + "'");
            return it->second;
        };
        h.data_dim = (int)need("data_dim").scalar_as_double();
        if (npz.count("data_format")) {
            h.data_format.parse(npz["data_format"].as_string());
        } else if (h.data_dim == 4) {
            h.data_format.format = DataFormat::RGBA;
            h.data_format.basis_dim = -1;
            fprintf(stderr, "INFO: Legacy file with no format specifier; spherical basis disabled\n");
        } else {
            h.data_format.format = DataFormat::SH;
            h.data_format.basis_dim = (h.data_dim - 1) / 3;
            fprintf(stderr, "INFO: Legacy file with no format specifier; autodetect spherical harmonics order\n");
        }
        fprintf(stderr, "INFO: Data format %s\n", h.data_format.to_string().c_str());
        if (npz.count("invradius3")) {
            const rtohost::NpyArray& s = npz["invradius3"];
            for (int i = 0; i < 3; ++i) h.scale[i] = s.word_size == 4 ? s.data<float>()[i] : (float)s.data<double>()[i];
        } else {
            h.scale[0] = h.scale[1] = h.scale[2] = (float)need("invradius").scalar_as_double();
        }
        {
            const rtohost::NpyArray& o = need("offset");
            for (int i = 0; i < 3; ++i) h.offset[i] = o.word_size == 4 ? o.data<float>()[i] : (float)o.data<double>()[i];
        }
        rtohost::NpyArray& child = need("child");
        if (child.word_size != 4 || child.shape.size() != 4) throw std::runtime_error("child must be int32 [cap,N,N,N]");
        h.N = (int)child.shape[1];
        h.child = child.data<int32_t>();
        if (h.N != 2) fprintf(stderr, "WARNING: N != 2 probably doesn't work.\n");
        const size_t N = (size_t)h.N, cap = child.shape[0], data_dim = (size_t)h.data_dim;
        if (cap == 0 || cap >= ((size_t)1 << 28) || child.shape[2] != N || child.shape[3] != N || h.data_dim < 1)
            throw std::runtime_error("malformed tree.npz: child shape / data_dim");
        const size_t n_child = cap * N * N * N;
        h.n_child = n_child;
        h.capacity = (int)cap;
        // Every array is uploaded with sizes derived from `child` ([cap,N,N,N]); a file whose arrays disagree with it must be
        // rejected HERE, because the C ABI receives raw pointers and cannot see the lengths behind them.
        auto expect = [&](const rtohost::NpyArray& a, const char* name, std::vector<size_t> shape, size_t word) {
            bool ok = a.word_size == word && a.shape.size() == shape.size() && !a.fortran_order;
            for (size_t i = 0; ok && i < shape.size(); ++i) ok = a.shape[i] == shape[i];
            if (!ok) throw std::runtime_error(std::string("malformed tree.npz: '") + name + "' has the wrong shape or type");
        };
        if (child.fortran_order) throw std::runtime_error("malformed tree.npz: 'child' is in Fortran order");
        if (npz.count("quant_colors")) {   // median-cut codebook files, src/n3tree.cpp:279-340
            fprintf(stderr, "INFO: Decoding quantized colors\n");
            const rtohost::NpyArray& qc = npz["quant_colors"];
            if (qc.word_size != 2) throw std::runtime_error("codebook must be stored in half precision");
            const rtohost::NpyArray& qm = need("quant_map");
            if (qm.shape.empty() || qc.shape.empty() || qc.shape[0] != qm.shape[0])
                throw std::runtime_error("codebook and map basis numbers does not match");
            const size_t n_q = qm.shape[0];
            const size_t n_ret = npz.count("data_retained") ? (npz["data_retained"].shape.empty() ? 0 : npz["data_retained"].shape[0]) : 0;
            expect(qc, "quant_colors", {n_q, 65536, 3}, 2);
            expect(qm, "quant_map", {n_q, cap, N, N, N}, 2);
            expect(need("sigma"), "sigma", {cap, N, N, N}, 2);
            if (n_ret) expect(npz["data_retained"], "data_retained", {n_ret, cap, N, N, N, 3}, 2);
            if (n_q + n_ret == 0 || 3 * (n_q + n_ret) > data_dim - 1) throw std::runtime_error("malformed tree.npz: codebook count does not fit data_dim");
            h.quantized = true;
            h.quant_colors = qc.data<uint16_t>(); h.quant_map = qm.data<uint16_t>(); h.sigma = need("sigma").data<uint16_t>();
            h.retained = n_ret ? npz["data_retained"].data<uint16_t>() : nullptr;
            h.n_quant = (int)n_q; h.n_retained = (int)n_ret;
            if (!decode_on_host) return;
            // Host materialisation of the dense array (only `--dry_run` needs it; the product gathers on the GPU).  Slot layout
            // of a leaf record (n3tree.cpp:301-338): colour channel k of basis function b sits at b + k*n_basis, retained basis
            // functions first, then the quantised ones; sigma last.  Written basis-major so each source plane streams once.
            const size_t n_basis = n_q + n_ret;
            h.decoded.assign(n_child * data_dim * 2, 0);
            uint16_t* out = reinterpret_cast<uint16_t*>(h.decoded.data());
            auto colour_slot = [&](size_t leaf, size_t basis, size_t k) -> uint16_t& { return out[leaf * data_dim + basis + k * n_basis]; };
            for (size_t b = 0; b < n_ret; ++b) {
                const uint16_t* src = h.retained + b * n_child * 3;
                for (size_t leaf = 0; leaf < n_child; ++leaf)
                    for (size_t k = 0; k < 3; ++k) colour_slot(leaf, b, k) = src[leaf * 3 + k];
            }
            for (size_t q = 0; q < n_q; ++q) {
                const uint16_t* book = h.quant_colors + q * 65536 * 3;
                const uint16_t* idx = h.quant_map + q * n_child;
                for (size_t leaf = 0; leaf < n_child; ++leaf)
                    for (size_t k = 0; k < 3; ++k) colour_slot(leaf, n_ret + q, k) = book[(size_t)idx[leaf] * 3 + k];
            }
            for (size_t leaf = 0; leaf < n_child; ++leaf) out[leaf * data_dim + data_dim - 1] = h.sigma[leaf];
            h.data = h.decoded.data();
        } else {
            const rtohost::NpyArray& d = need("data");
            if (d.word_size != 2) throw std::runtime_error("data must be stored in half precision");
            expect(d, "data", {cap, N, N, N, data_dim}, 2);
            h.data = d.bytes.data();
        }
    }

   private:
    void load_npz(rtohost::npz_t& npz) {
        HostArrays h;
        decode_npz(npz, h, /*decode_on_host=*/false);
        N = h.N; data_dim = h.data_dim; capacity = h.capacity; data_format = h.data_format; scale = h.scale; offset = h.offset;
        printf("INFO: Scale %f %f %f\n", scale[0], scale[1], scale[2]);
        if (h.quantized)
            rto_check(rto_tree_create_quantized(&device, h.child, capacity, N, data_dim, (int)data_format.format,
                                                data_format.basis_dim, offset.data(), scale.data(), h.quant_colors,
                                                h.quant_map, h.n_quant, h.sigma, h.retained, h.n_retained),
                      "rto_tree_create_quantized");
        else
            rto_check(rto_tree_create(&device, h.child, h.data, capacity, N, data_dim, (int)data_format.format,
                                      data_format.basis_dim, offset.data(), scale.data()),
                      "rto_tree_create");
    }
};

// -------------------------------------------------------------------------------------------------- Camera
static const float CAMERA_DEFAULT_FOCAL_LENGTH = 1111.11f;
struct Camera {
    Camera(int width = 256, int height = 256, float fx = CAMERA_DEFAULT_FOCAL_LENGTH, float fy = -1.f)
        : width(width), height(height), fx(fx < 0.f ? CAMERA_DEFAULT_FOCAL_LENGTH : fx), fy(fy < 0.f ? this->fx : fy) {
        for (float& v : transform) v = 0.f;
    }
    int width, height;
    float fx, fy;
    float transform[12];   // column-major 4x3 c2w: right, up, back, centre  (glm::mat4x3 memory order)
    // the reference uploads the 48-byte transform here (src/camera.cpp:72-73); it travels as a kernel argument now
    void _update(bool = true, bool = true) {}
    rto_camera pod() const {
        rto_camera c;
        c.width = width; c.height = height; c.fx = fx; c.fy = fy;
        memcpy(c.c2w, transform, sizeof transform);
        return c;
    }
};

// ------------------------------------------------------------------------------------------- RenderOptions
struct RenderOptions {
    float step_size = 1e-4f;
    float sigma_thresh = 1e-2f;
    float stop_thresh = 1e-2f;
    float background_brightness = 1.f;
    bool show_grid = false;
    int grid_max_depth = 4;
    bool enable_probe = false;
    float probe[3] = {0.f, 0.f, 1.f};
    int probe_disp_size = 100;
    bool denoise = true;
    int spp = 1;

    // NLOHMANN_DEFINE_TYPE_INTRUSIVE(...): all 11 keys are mandatory (render_options.hpp:61-77)
    static RenderOptions from_json(const rtohost::Json& j) {
        RenderOptions o;
        o.step_size = (float)j.at("step_size").as_number();
        o.sigma_thresh = (float)j.at("sigma_thresh").as_number();
        o.stop_thresh = (float)j.at("stop_thresh").as_number();
        o.background_brightness = (float)j.at("background_brightness").as_number();
        o.show_grid = j.at("show_grid").as_bool();
        o.grid_max_depth = (int)j.at("grid_max_depth").as_number();
        o.enable_probe = j.at("enable_probe").as_bool();
        const rtohost::Json& p = j.at("probe");
        for (int i = 0; i < 3; ++i) o.probe[i] = (float)p[i].as_number();
        o.probe_disp_size = (int)j.at("probe_disp_size").as_number();
        o.denoise = j.at("denoise").as_bool();
        o.spp = (int)j.at("spp").as_number();
        return o;
    }
    rto_render_options pod() const {
        rto_render_options r;
        r.step_size = step_size; r.sigma_thresh = sigma_thresh; r.stop_thresh = stop_thresh;
        r.background_brightness = background_brightness; r.denoise = denoise; r.spp = spp; r.enable_probe = enable_probe;
        return r;
    }
};

// ------------------------------------------------------------------------------------------- RenderContext
struct RenderContext {
    static constexpr int CHANNELS = 8;
    rto_context* handle = nullptr;
    float* aux_buffer = nullptr;   // device [8][H][W]
    float* image = nullptr;        // device [H][W][4]  (the reference writes a cudaArray surface)
    bool offscreen = true;
    int width = 0, height = 0;

    struct Rng {   // ctx.rng: pcg32(20230418) with advance() (render_context.hpp:16; pcg32.h:145)
        RenderContext* c;
        void advance(int64_t delta = (1ll << 32)) { rto_check(rto_context_rng_advance(c->handle, delta), "rng.advance"); }
    } rng{this};

    RenderContext() = default;
    RenderContext(const RenderContext&) = delete;
    ~RenderContext() { freeResource(); }

    void freeResource() {
        rto_context_destroy(handle);
        handle = nullptr;
        aux_buffer = image = nullptr;
    }
    // RenderContext::update(image_arr, depth_arr, width, height): the output image lives in the context here
    void update(int w, int h) {
        freeResource();
        width = w; height = h;
        rto_check(rto_context_create(&handle, w, h), "rto_context_create");
        aux_buffer = rto_context_aux(handle);
        image = rto_context_image(handle);
    }

    struct Timer {   // RenderContext::Timer (render_context.hpp:122-213)
        RenderContext* c;
        void reset(void* /*stream*/) {
            rto_check(rto_timer_enable(c->handle, 1), "timer");
            rto_check(rto_timer_reset(c->handle), "timer");
        }
        // stage events are recorded inside rto_render / rto_denoise; kept for source compatibility
        void render_start() {} void render_stop() {} void torch_start() {} void torch_stop() {}
        void filter_start() {} void filter_stop() {}
        void record(bool denoise) { rto_check(rto_timer_record(c->handle, denoise), "timer.record"); }
        void report() const {
            float ms[3];
            int n = 0;
            rto_check(rto_timer_report(c->handle, ms, &n), "timer.report");
            const float all = ms[0] + ms[1] + ms[2];
            printf("render: %.10f ms per frame\n", ms[0]);
            printf("torch:  %.10f ms per frame\n", ms[1]);
            printf("filter: %.10f ms per frame\n", ms[2]);
            printf("all:    %.10f ms per frame\n", all);
            printf("FPS:    %.10f\n", 1000.f / all);
        }
    };
    Timer timer() { return Timer{this}; }
};

// launch_renderer(tree, cam, options, ctx, stream, offscreen): asynchronous on `stream`; throws std::runtime_error for
// an unsupported SPP (src/cuda/volrend.cu:275-277).  Only the offscreen path exists (no GL interop).
inline void launch_renderer(const N3Tree& tree, const Camera& cam, const RenderOptions& options, RenderContext& ctx,
                            void* stream, bool offscreen = true) {
    if (!offscreen) throw std::runtime_error("launch_renderer: only offscreen rendering is supported");
    if (!tree.device) throw std::runtime_error("launch_renderer: tree not loaded");
    const rto_camera c = cam.pod();
    const rto_render_options o = options.pod();
    rto_check(rto_render(ctx.handle, tree.device, &c, &o, stream), "launch_renderer");
}

// ------------------------------------------------------------------------------------------------ Denoiser
class Denoiser final {
   public:
    // The reference takes a TorchScript file (denoiser.cpp:12-26).  Here `path` is the raw-tensor export of that file
    // made once by tools/make_ts_module.py --export (an .npz with w1,b1,w2,b2 fp16); if a .ts path is given, the
    // export is looked up next to it as <path>.npz or <stem>.npz.
    explicit Denoiser(const std::string& ts_module_path) {
        if (ts_module_path.empty()) throw std::runtime_error("No torchscript module is given to denoiser.");
        std::string p = ts_module_path;
        auto exists = [](const std::string& f) { return bool(std::ifstream(f)); };
        if (p.size() > 3 && p.substr(p.size() - 3) == ".ts") {
            if (exists(p + ".npz")) p = p + ".npz";
            else if (exists(p.substr(0, p.size() - 3) + ".npz")) p = p.substr(0, p.size() - 3) + ".npz";
            else throw std::runtime_error("Error when loading torchscript model from " + ts_module_path +
                                          " (export it once: python tools/make_ts_module.py --export " + ts_module_path +
                                          " --out " + ts_module_path + ".npz)");
        }
        rtohost::npz_t z;
        try {
            z = rtohost::npz_load(p);
        } catch (const std::exception& e) {
            fprintf(stderr, "%s\n", e.what());
            throw std::runtime_error("Error when loading torchscript model from " + ts_module_path);
        }
        for (const char* k : {"w1", "b1", "w2", "b2"})
            if (!z.count(k) || z[k].word_size != 2) throw std::runtime_error(std::string("GuidanceNet export: missing fp16 tensor ") + k);
        const int mid = (int)z["w1"].shape[0], in_ch = (int)z["w1"].shape[1], levels = (int)z["w2"].shape[0] / 2;
        rto_check(rto_net_create(&net_, z["w1"].bytes.data(), z["b1"].bytes.data(), z["w2"].bytes.data(), z["b2"].bytes.data(),
                                 in_ch, mid, levels),
                  "rto_net_create");
        levels_ = levels;
    }
    Denoiser(const Denoiser&) = delete;
    ~Denoiser() { rto_net_destroy(net_); }
    void denoise(const Camera&, RenderContext& ctx, void* stream) { rto_check(rto_denoise(ctx.handle, net_, stream), "denoise"); }
    rto_net* handle() { return net_; }
    int levels() const { return levels_; }   // filter levels L: the tile split renders a halo of 2 + L rows per band

   private:
    rto_net* net_ = nullptr;
    int levels_ = 4;
};

}  // namespace volrend
