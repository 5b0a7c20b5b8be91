// volrend_headless (B200-native) — drop-in for the reference's headless driver, renderer/main_headless.cpp:192-552.
//
//   volrend_headless <tree.npz> <poses> [--options opt.json] [--ts_module weights] [--dataset {blender,tt,llff}]
//                    [-o dir] [--write_buffer] [--gpu id] [-w -h --fx --fy --bg -s -e -a --scale --max_imgs -r -i --draw]
//
// Same flags (opts.cpp:7-31 + main_headless.cpp:202-223), same pose loaders (:255-370), camera conventions (:372-390),
// 100-frame warm-up protocol (:469-479), per-pose rng.advance() (:506), `buf_<basename>.bin` layout (:512-523) and timer
// report (render_context.hpp:190-206).  Additions (all optional): --num_gpus N / --gpu_list a,b,.. shard the poses over
// GPUs (one host thread + one replica of the tree per shard, no communication); --pipe N keeps N frames in flight on N
// (context, stream) pairs, each frame one CUDA-graph launch (--no_graph: separate launches), outputs land in pinned ring
// buffers and are written by --writers threads; --readback {rgba8,float,aux} copies results to the host every frame even
// without -o (timing with the device->host copy inside); --warmup K; --write_float also dumps the final float4 image as
// `img_<basename>.bin`; --dump_poses F; --dry_run parses every input and prints a JSON summary without touching a GPU.
// Every frame's rng state is a pure function of its global index, so all of these modes write identical files.
// Differences, on purpose: the tt pose directory is read in sorted order (the reference iterates it unsorted, :282);
// PNGs are written with zlib directly (the reference needs libpng).
#include <zlib.h>

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <filesystem>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "volrend_b200.hpp"

namespace fs = std::filesystem;
using namespace volrend;

namespace {

struct Mat43 { float m[12]; };   // column-major 4x3 (glm::mat4x3): m[c*3+r]

// --------------------------------------------------------------------------------------------- CLI parsing
struct Args {
    std::map<std::string, std::string> kv;
    std::vector<std::string> positional;
    bool has(const std::string& k) const { return kv.count(k) > 0; }
    std::string str(const std::string& k, const std::string& d = "") const { auto it = kv.find(k); return it == kv.end() ? d : it->second; }
    int i(const std::string& k, int d) const { return has(k) ? atoi(kv.at(k).c_str()) : d; }
    float f(const std::string& k, float d) const { return has(k) ? (float)atof(kv.at(k).c_str()) : d; }
};

Args parse_args(int argc, char** argv) {
    static const std::map<std::string, std::string> shorts = {{"w", "width"}, {"h", "height"}, {"s", "step_size"}, {"e", "stop_thresh"},
                                                              {"a", "sigma_thresh"}, {"o", "write_images"}, {"i", "intrin"}, {"r", "reverse_yz"}};
    static const std::set<std::string> flags = {"reverse_yz", "write_buffer", "help", "dry_run", "write_float", "no_graph", "tile_split", "band_readback"};
    Args a;
    for (int k = 1; k < argc; ++k) {
        std::string t = argv[k];
        if (t.size() >= 2 && t[0] == '-' && !(std::isdigit((unsigned char)t[1]) || t[1] == '.')) {
            std::string key = t.substr(t[1] == '-' ? 2 : 1), val;
            const size_t eq = key.find('=');
            bool has_val = false;
            if (eq != std::string::npos) { val = key.substr(eq + 1); key = key.substr(0, eq); has_val = true; }
            if (shorts.count(key)) key = shorts.at(key);
            if (flags.count(key)) { a.kv[key] = has_val ? val : "1"; continue; }
            if (!has_val) {
                if (k + 1 >= argc) { fprintf(stderr, "Option '%s' is missing an argument\n", t.c_str()); std::exit(1); }
                val = argv[++k];
            }
            a.kv[key] = val;
        } else {
            a.positional.push_back(t);
        }
    }
    return a;
}

void print_help() {
    puts("Headless PlenOctree volume rendering, B200-native RT-Octree path\n"
         "Usage:\n  volrend_headless [OPTION...] npz_file [c2w_txt_4x4...]\n\n"
         "      --file arg          npz file storing octree data\n"
         "      --gpu arg           CUDA device id (default: -1)\n"
         "  -w, --width arg         image width (default: 800)\n"
         "  -h, --height arg        image height (default: 800)\n"
         "      --fx arg            focal length in x direction; -1 = 1111 or default for NDC (default: -1.0)\n"
         "      --fy arg            focal length in y direction; -1 = use fx (default: -1.0)\n"
         "      --bg arg            background brightness 0-1 (default: 1.0)\n"
         "  -s, --step_size arg     step size epsilon added to computed cube size (default: 1e-4)\n"
         "  -e, --stop_thresh arg   early stopping threshold (on remaining intensity) (default: 1e-2)\n"
         "  -a, --sigma_thresh arg  sigma threshold (skip cells with < sigma) (default: 1e-2)\n"
         "      --help              Print this help message\n"
         "  -o, --write_images arg  output directory of images; if empty, DOES NOT save (for timing only)\n"
         "  -i, --intrin arg        intrinsics matrix 4x4 (accepted, unused - as in the reference)\n"
         "  -r, --reverse_yz        use OpenCV camera space convention instead of NeRF\n"
         "      --scale arg         scaling to apply to image (default: 1)\n"
         "      --max_imgs arg      max images to render, default no limit (default: 0)\n"
         "      --options arg       render options json\n"
         "      --dataset arg       dataset type: blender | tt | llff (default: blender)\n"
         "      --ts_module arg     GuidanceNet weights (npz export of the reference's ts_*.ts)\n"
         "      --write_buffer      save auxiliary buffers (buf_<name>.bin). Invalid if output directory is not given.\n"
         "      --num_gpus arg      shard the poses over this many GPUs (default: 1)\n"
         "      --gpu_list arg      comma-separated device ids, one shard per entry (e.g. 0,1,2,3; 0,0 = two shards on GPU 0)\n"
         "      --tile_split        single-frame latency mode: every frame is cut into row bands over the GPUs of --num_gpus /\n"
         "                          --gpu_list; each band's filter stores into GPU 0's image over NVLink (no gather)\n"
         "      --band_readback     with --tile_split: every GPU copies its own band to the host frame (N PCIe links in parallel)\n"
         "                          instead of storing it into the first GPU's image and copying the assembled frame from there\n"
         "      --pipe arg          frames in flight per GPU: N contexts/streams, CUDA-graph frames (default: 1 = reference protocol)\n"
         "      --no_graph          with --pipe: issue the kernels separately instead of one graph launch per frame\n"
         "      --readback arg      copy rgba8 | float | aux to pinned host memory every frame even without -o\n"
         "      --writers arg       file-writer threads used with --pipe (default: 4)\n"
         "      --warmup arg        warm-up frames on pose 0 (default: 100)\n"
         "      --write_float       also save the final float4 image as img_<name>.bin\n"
         "      --dump_poses arg    write every camera transform (after loader + convention) as float32 [n][12]\n"
         "      --dry_run           parse all inputs, print a JSON summary, do not touch the GPU");
}

// ------------------------------------------------------------------------------------------- small vector math
struct V3 { float x, y, z; };
V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
V3 normalize3(V3 v) {   // main_headless.cpp:102-106
    const float n = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return {v.x / n, v.y / n, v.z / n};
}
V3 col(const Mat43& t, int c) { return {t.m[c * 3], t.m[c * 3 + 1], t.m[c * 3 + 2]}; }

// inverse(poses_avg) * pose, in exactly the fp32 operation order of the reference's glm calls, so that the recentred llff
// poses (and through them every ray) are bit-identical to the reference CLI's:
//   glm::inverse(mat4)  = detail::compute_inverse<4,4>  (3rdparty/glm/glm/detail/func_matrix.inl:347-405): 18 2x2 sub-
//                         determinants "Coef", four cofactor columns (a*F - b*F) + c*F, sign flips, the determinant as
//                         (d.x + d.y) + (d.z + d.w) of column 0 times row 0 of the cofactor matrix, one reciprocal,
//                         every element multiplied by it;
//   mat4 * mat4         = ((A0*b0 + A1*b1) + A2*b2) + A3*b3 per column (detail/type_mat4x4.inl:630-648).
// The reference is built without -ffast-math and without FMA contraction (CMakeLists.txt:30,66: x86-64 baseline), as is this
// file (host/Makefile): plain IEEE multiplies and adds in source order.  m[c][r] = column c, row r.
typedef float M4[4][4];
void inverse4_glm(const M4 m, M4 out) {
    const float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
    const float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
    const float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    const float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    const float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    const float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    const float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    const float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
    const float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    const float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    const float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
    const float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    const float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
    const float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
    const float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    const float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    const float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
    const float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    const float Fac0[4] = {Coef00, Coef00, Coef02, Coef03}, Fac1[4] = {Coef04, Coef04, Coef06, Coef07};
    const float Fac2[4] = {Coef08, Coef08, Coef10, Coef11}, Fac3[4] = {Coef12, Coef12, Coef14, Coef15};
    const float Fac4[4] = {Coef16, Coef16, Coef18, Coef19}, Fac5[4] = {Coef20, Coef20, Coef22, Coef23};
    const float Vec0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, Vec1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
    const float Vec2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, Vec3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
    static const float SignA[4] = {+1, -1, +1, -1}, SignB[4] = {-1, +1, -1, +1};
    M4 inv;
    for (int i = 0; i < 4; ++i) {
        const float Inv0 = (Vec1[i] * Fac0[i] - Vec2[i] * Fac1[i]) + Vec3[i] * Fac2[i];
        const float Inv1 = (Vec0[i] * Fac0[i] - Vec2[i] * Fac3[i]) + Vec3[i] * Fac4[i];
        const float Inv2 = (Vec0[i] * Fac1[i] - Vec1[i] * Fac3[i]) + Vec3[i] * Fac5[i];
        const float Inv3 = (Vec0[i] * Fac2[i] - Vec1[i] * Fac4[i]) + Vec2[i] * Fac5[i];
        inv[0][i] = Inv0 * SignA[i];
        inv[1][i] = Inv1 * SignB[i];
        inv[2][i] = Inv2 * SignA[i];
        inv[3][i] = Inv3 * SignB[i];
    }
    const float Dot0[4] = {m[0][0] * inv[0][0], m[0][1] * inv[1][0], m[0][2] * inv[2][0], m[0][3] * inv[3][0]};
    const float Dot1 = (Dot0[0] + Dot0[1]) + (Dot0[2] + Dot0[3]);
    const float OneOverDeterminant = 1.0f / Dot1;
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) out[c][r] = inv[c][r] * OneOverDeterminant;
}

// _recenter_poses (main_headless.cpp:152-189): pose <- mat4x3(inverse(expand(poses_avg)) * expand(pose))
void recenter_poses(std::vector<Mat43>& trans) {
    V3 z{0, 0, 0}, up{0, 0, 0}, cen{0, 0, 0};
    for (const Mat43& t : trans) { z = z + col(t, 2); up = up + col(t, 1); cen = cen + col(t, 3); }
    const float n = (float)trans.size();
    z = normalize3(z / n);
    up = up / n;
    cen = cen / n;
    // _viewmatrix(z, up, pos)
    z = normalize3(z);
    const V3 x = normalize3(cross(up, z));
    const V3 y = normalize3(cross(z, x));
    const M4 c2w = {{x.x, x.y, x.z, 0.0f}, {y.x, y.y, y.z, 0.0f}, {z.x, z.y, z.z, 0.0f}, {cen.x, cen.y, cen.z, 1.0f}};
    M4 A;
    inverse4_glm(c2w, A);
    for (Mat43& p : trans) {
        const M4 B = {{p.m[0], p.m[1], p.m[2], 0.0f}, {p.m[3], p.m[4], p.m[5], 0.0f}, {p.m[6], p.m[7], p.m[8], 0.0f}, {p.m[9], p.m[10], p.m[11], 1.0f}};
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 3; ++r)
                p.m[c * 3 + r] = ((A[0][r] * B[c][0] + A[1][r] * B[c][1]) + A[2][r] * B[c][2]) + A[3][r] * B[c][3];
    }
}

std::string remove_ext(const std::string& s) {
    const size_t p = s.rfind('.');
    return p == std::string::npos ? s : s.substr(0, p);
}

std::string read_text(const std::string& path) {
    std::ifstream f(path);
    if (!f) { fprintf(stderr, "ERROR: '%s' does not exist\n", path.c_str()); std::exit(1); }
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// read_transform_matrices (main_headless.cpp:62-91): a file may hold several 4x4 row-major matrices
int read_transform_matrices(const std::string& path, std::vector<Mat43>& out) {
    std::ifstream ifs(path);
    if (!ifs) { fprintf(stderr, "ERROR: '%s' does not exist\n", path.c_str()); std::exit(1); }
    int cnt = 0;
    while (ifs) {
        Mat43 t;
        float garb;
        ifs >> t.m[0] >> t.m[3] >> t.m[6] >> t.m[9];
        if (!ifs) break;
        ifs >> t.m[1] >> t.m[4] >> t.m[7] >> t.m[10];
        ifs >> t.m[2] >> t.m[5] >> t.m[8] >> t.m[11];
        if (ifs) ifs >> garb >> garb >> garb >> garb;
        ++cnt;
        out.push_back(t);
    }
    return cnt;
}

// ------------------------------------------------------------------------------------------------- PNG (RGBA8)
void put32(std::vector<unsigned char>& v, uint32_t x) { for (int s = 24; s >= 0; s -= 8) v.push_back((unsigned char)(x >> s)); }
void png_chunk(std::vector<unsigned char>& out, const char* type, const std::vector<unsigned char>& data) {
    put32(out, (uint32_t)data.size());
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    put32(out, (uint32_t)crc32(0L, out.data() + start, (uInt)(out.size() - start)));
}
bool write_png_file(const std::string& path, const uint8_t* rgba, int w, int h) {   // src/imwrite.cpp:14-78 (RGBA8, compression 0)
    std::vector<unsigned char> raw((size_t)h * (w * 4 + 1));
    for (int y = 0; y < h; ++y) {
        raw[(size_t)y * (w * 4 + 1)] = 0;
        memcpy(&raw[(size_t)y * (w * 4 + 1) + 1], rgba + (size_t)y * w * 4, (size_t)w * 4);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<unsigned char> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 1) != Z_OK) return false;
    comp.resize(clen);
    std::vector<unsigned char> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<unsigned char> ihdr;
    put32(ihdr, (uint32_t)w); put32(ihdr, (uint32_t)h);
    ihdr.insert(ihdr.end(), {8, 6, 0, 0, 0});
    png_chunk(out, "IHDR", ihdr);
    png_chunk(out, "IDAT", comp);
    png_chunk(out, "IEND", {});
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(out.data()), (std::streamsize)out.size());
    return bool(f);
}

uint64_t fnv64(const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    uint64_t h = 0xcbf29ce484222325ULL;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ULL; }
    return h;
}

struct Job {
    std::string tree_path, out_dir, ts_module, readback;
    std::vector<Mat43> trans;
    std::vector<std::string> basenames;
    RenderOptions options;
    int width, height;
    float fx, fy;
    bool llff, write_buffer, write_float, graph, tile_split, band_readback;
    int warmup, pipe, writers;
};

struct ShardStats {
    float ms[3] = {0.f, 0.f, 0.f};   // mean render / net / filter ms per frame (stage events; serial protocol only)
    int frames = 0;
    double wall_s = 0.0;             // wall clock of the timed loop (first launch -> last result on the host)
    bool staged = false;             // ms[] valid
};

void write_outputs(const Job& job, size_t i, const float* aux, const uint8_t* rgba8, const float* img) {
    if (job.write_buffer && aux) {   // main_headless.cpp:512-523
        std::ofstream out(job.out_dir + "/buf_" + job.basenames[i] + ".bin", std::ios::out | std::ios::binary);
        out.write(reinterpret_cast<const char*>(aux), (std::streamsize)((size_t)RenderContext::CHANNELS * job.width * job.height * sizeof(float)));
    } else if (rgba8) {              // main_headless.cpp:524-541
        write_png_file(job.out_dir + "/" + job.basenames[i] + ".png", rgba8, job.width, job.height);
    }
    if (job.write_float && img) {
        std::ofstream out(job.out_dir + "/img_" + job.basenames[i] + ".bin", std::ios::out | std::ios::binary);
        out.write(reinterpret_cast<const char*>(img), (std::streamsize)((size_t)4 * job.width * job.height * sizeof(float)));
    }
}

// a few threads that turn finished frames into files while the GPU renders the next ones
class WriterPool {
   public:
    explicit WriterPool(int n) {
        for (int i = 0; i < n; ++i) th_.emplace_back([this] { loop(); });
    }
    ~WriterPool() {
        { std::lock_guard<std::mutex> l(mu_); stop_ = true; }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    void submit(std::function<void()> f) {
        { std::lock_guard<std::mutex> l(mu_); q_.push_back(std::move(f)); }
        cv_.notify_one();
    }
   private:
    void loop() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                f = std::move(q_.front());
                q_.pop_front();
            }
            f();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    std::vector<std::thread> th_;
    bool stop_ = false;
};

// one frame slot of the pipeline: its own context + stream (+ graph) and pinned result buffers
struct Slot {
    RenderContext ctx;
    void* stream = nullptr;
    rto_frame* frame = nullptr;
    uint8_t* h_rgba8 = nullptr;
    float* h_img = nullptr;
    float* h_aux = nullptr;
    long pending = -1;                 // global index of the frame in flight on this slot
    std::mutex mu;                     // writer hand-off: `writing` is true while a writer thread reads the pinned buffers
    std::condition_variable cv;
    bool writing = false;
    ~Slot() {
        rto_frame_destroy(frame);
        rto_host_free(h_rgba8); rto_host_free(h_img); rto_host_free(h_aux);
        ctx.freeResource();
        rto_stream_destroy(stream);
    }
};

// --pipe N: N frames in flight.  Frame i runs on slot i % N: wait until the slot's previous frame (and its file writer) are
// done, then enqueue render -> denoise -> read-backs for pose i without any further host synchronisation.
void run_shard_pipelined(const Job& job, N3Tree& tree, Denoiser& denoiser, size_t begin, size_t end, ShardStats& st) {
    const int W = job.width, H = job.height;
    const size_t px = (size_t)W * H;
    const bool files = job.out_dir.size() > 0;
    const bool want_aux = (files && job.write_buffer) || job.readback == "aux";
    const bool want_rgba8 = (files && !job.write_buffer) || job.readback == "rgba8";
    const bool want_img = (files && job.write_float) || job.readback == "float";
    const rto_render_options opt = job.options.pod();
    std::vector<std::unique_ptr<Slot>> slots;
    for (int k = 0; k < job.pipe; ++k) {
        auto s = std::make_unique<Slot>();
        s->ctx.offscreen = true;
        s->ctx.update(W, H);
        rto_check(rto_stream_create(&s->stream), "stream");
        if (want_rgba8) rto_check(rto_host_alloc(reinterpret_cast<void**>(&s->h_rgba8), px * 4), "pinned rgba8");
        if (want_img) rto_check(rto_host_alloc(reinterpret_cast<void**>(&s->h_img), px * 16), "pinned image");
        if (want_aux) rto_check(rto_host_alloc(reinterpret_cast<void**>(&s->h_aux), px * 32), "pinned aux");
        if (job.graph) {
            rto_frame_desc d{};
            d.tree = tree.device; d.net = job.options.denoise ? denoiser.handle() : nullptr; d.opt = opt;
            d.fx = job.fx; d.fy = job.fy;
            d.host_rgba8 = s->h_rgba8; d.host_image = s->h_img; d.host_aux = s->h_aux;
            rto_check(rto_frame_create(&s->frame, s->ctx.handle, &d), "rto_frame_create");
        }
        slots.push_back(std::move(s));
    }
    rto_camera cam{};
    cam.width = W; cam.height = H; cam.fx = job.fx; cam.fy = job.fy;
    auto issue = [&](Slot& s, const Mat43& pose, int64_t warm, int64_t frame) {
        rto_check(rto_context_rng_set_frame(s.ctx.handle, warm, frame), "rng");   // == `frame` advances after the warm-up
        if (s.frame) {
            rto_check(rto_frame_launch(s.frame, pose.m, s.stream), "rto_frame_launch");
            return;
        }
        memcpy(cam.c2w, pose.m, sizeof cam.c2w);
        rto_check(rto_render(s.ctx.handle, tree.device, &cam, &opt, s.stream), "launch_renderer");
        if (job.options.denoise) rto_check(rto_denoise(s.ctx.handle, denoiser.handle(), s.stream), "denoise");
        if (s.h_rgba8) rto_check(rto_context_read_image_rgba8(s.ctx.handle, s.h_rgba8, s.stream), "read rgba8");
        if (s.h_img) rto_check(rto_context_read_image(s.ctx.handle, s.h_img, s.stream), "read image");
        if (s.h_aux) rto_check(rto_context_read_aux(s.ctx.handle, s.h_aux, s.stream), "read aux");
    };
    // warm up on pose 0 (main_headless.cpp:469-479): frame w of the warm-up uses the rng state after w advances
    for (int w = 0; w < job.warmup; ++w) issue(*slots[w % job.pipe], job.trans[0], 0, w);
    for (auto& s : slots) rto_check(rto_synchronize(s->stream), "sync");

    std::unique_ptr<WriterPool> pool;
    if (files) pool = std::make_unique<WriterPool>(job.writers < 1 ? 1 : job.writers);
    auto retire = [&](Slot& s) {   // the slot's frame is complete on the host: hand it to a writer
        rto_check(rto_synchronize(s.stream), "sync");
        if (s.pending < 0) return;
        const size_t i = (size_t)s.pending;
        s.pending = -1;
        if (!files) return;
        { std::lock_guard<std::mutex> l(s.mu); s.writing = true; }
        Slot* sp = &s;
        pool->submit([&job, sp, i] {
            write_outputs(job, i, sp->h_aux, sp->h_rgba8, sp->h_img);
            { std::lock_guard<std::mutex> l(sp->mu); sp->writing = false; }
            sp->cv.notify_all();
        });
    };
    const auto t0 = std::chrono::steady_clock::now();
    if (job.graph && end > begin) {
        // the whole loop is ONE library call (rto_frame_sequence): frame i on slot i % N, the slot's stream waited for before
        // reuse; the callback hands a finished frame's pinned buffers to a writer and returns when they are free again
        static_assert(sizeof(Mat43) == 12 * sizeof(float), "poses are passed as a dense [n][12] array");
        std::vector<rto_frame*> frames;
        std::vector<void*> streams;
        for (auto& s : slots) { frames.push_back(s->frame); streams.push_back(s->stream); }
        struct Retire {
            decltype(retire)* fn;
            std::vector<std::unique_ptr<Slot>>* slots;
        } ctx{&retire, &slots};
        auto cb = [](void* user, int64_t frame_index, int slot) {
            Retire& r = *static_cast<Retire*>(user);
            Slot& s = *(*r.slots)[(size_t)slot];
            s.pending = (long)frame_index;
            (*r.fn)(s);
            std::unique_lock<std::mutex> l(s.mu);
            s.cv.wait(l, [&] { return !s.writing; });   // pinned buffers free again
        };
        rto_check(rto_frame_sequence(frames.data(), streams.data(), job.pipe, job.trans[0].m, (int64_t)job.trans.size(), job.warmup,
                                     (int64_t)begin, (int64_t)(end - begin), 1, cb, &ctx), "rto_frame_sequence");
    } else {
        for (size_t i = begin; i < end; ++i) {
            Slot& s = *slots[(i - begin) % job.pipe];
            retire(s);
            { std::unique_lock<std::mutex> l(s.mu); s.cv.wait(l, [&] { return !s.writing; }); }   // pinned buffers free again
            issue(s, job.trans[i], job.warmup, (int64_t)i);
            s.pending = (long)i;
        }
        for (auto& s : slots) retire(*s);
    }
    st.wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();   // results on the host
    for (auto& s : slots) { std::unique_lock<std::mutex> l(s->mu); s->cv.wait(l, [&] { return !s->writing; }); }
    pool.reset();
    st.frames = (int)(end - begin);
    st.staged = false;
}

// reusable spinning barrier for the per-frame rendezvous of the tile-split threads (a frame lasts a few hundred microseconds:
// a sleeping barrier would cost more than the work it orders)
class SpinBarrier {
   public:
    explicit SpinBarrier(int n) : n_(n) {}
    void wait() {
        const int gen = gen_.load(std::memory_order_acquire);
        if (count_.fetch_add(1, std::memory_order_acq_rel) == n_ - 1) {
            count_.store(0, std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_release);
        } else {
            while (gen_.load(std::memory_order_acquire) == gen) {}
        }
    }
   private:
    const int n_;
    std::atomic<int> count_{0}, gen_{0};
};

// Band balancer of the tile split: the bands through the object cost more than the top and bottom ones, and the frame is as
// slow as its slowest band.  After every frame each band reports its device time; the cost per row is taken as constant
// inside a band, the cumulative cost is cut into equal parts and the boundaries move half way to those cuts (damping: poses
// change from frame to frame).  Any partition of the rows gives the same pixels, so this is purely a scheduling decision.
std::vector<int> rebalance_bands(const std::vector<int>& bounds, const std::vector<float>& ms, int H, int min_rows = 16) {
    const int n = (int)ms.size();
    std::vector<double> cum(n + 1, 0.0);
    for (int g = 0; g < n; ++g) cum[g + 1] = cum[g] + std::max(1e-6, (double)ms[g]);
    std::vector<int> out(bounds);
    for (int k = 1; k < n; ++k) {
        const double target = cum[n] * k / n;
        int g = 0;
        while (g + 1 < n && cum[g + 1] < target) ++g;
        const double frac = (target - cum[g]) / (cum[g + 1] - cum[g]);
        const double y = bounds[g] + frac * (bounds[g + 1] - bounds[g]);
        out[k] = (int)std::lround(0.5 * bounds[k] + 0.5 * y);
    }
    for (int k = 1; k < n; ++k) out[k] = std::max(out[k], out[k - 1] + min_rows);
    for (int k = n - 1; k >= 1; --k) out[k] = std::min(out[k], out[k + 1] - min_rows);
    out[0] = 0; out[n] = H;
    return out;
}

// --tile_split: single-frame latency mode.  One host thread per GPU; for every frame each thread renders the
// rows of its band plus the denoiser's halo (2 rows for the two 3x3 convolutions + L for the largest filter level), runs the
// GuidanceNet and the filter on its band, and the filter's epilogue stores the band directly into GPU 0's image (float4 and
// RGBA8) through a peer mapping — no gather, no halo exchange.  GPU 0's stream waits on one event per band, then copies the
// assembled frame to the host.  Latency = rendezvous of the threads -> frame in host memory.
void run_tile_split(const Job& job, const std::vector<int>& devices) {
    const int n = (int)devices.size(), W = job.width, H = job.height;
    const size_t px = (size_t)W * H;
    const size_t frames = job.trans.size();
    SpinBarrier bar(n);
    std::vector<void*> events(n, nullptr);
    std::vector<std::string> errors(n);
    std::atomic<bool> failed{false};
    float* dst_img = nullptr;          // GPU 0's image / RGBA8 copy: the destination of every band
    unsigned char* dst_img8 = nullptr;
    // --band_readback: the bands never meet on a GPU; every thread copies its own rows into the shared pinned host frame
    const bool direct = job.band_readback;
    uint8_t* host8 = nullptr;          // the first thread's pinned frame(s), written by every thread in that mode
    float* host_img = nullptr;
    int levels = 4;
    std::vector<double> lat_ms, dev_ms;
    std::vector<int> bounds(n + 1);                  // band g = rows [bounds[g], bounds[g+1]); equal heights to start with
    for (int g = 0; g <= n; ++g) bounds[g] = (int)((int64_t)H * g / n);
    std::vector<float> band_ms(n, 1.f);
    const bool balance = n > 1 && !getenv("RTO_TILE_SPLIT_STATIC");
    const rto_render_options opt = job.options.pod();
    auto worker = [&](int g) {
        try {
            rto_check(rto_set_device(devices[g]), "cudaSetDevice");
            N3Tree tree(job.tree_path);
            if (!tree.is_data_loaded()) throw std::runtime_error("tree not loaded");
            if (job.llff) { tree.use_ndc = true; tree.ndc_width = (float)W; tree.ndc_height = (float)H; tree.ndc_focal = job.fx; }
            tree.sync_ndc();
            Denoiser denoiser(job.ts_module);
            RenderContext ctx;
            ctx.update(W, H);
            void* stream = nullptr;
            rto_check(rto_stream_create(&stream), "stream");
            rto_check(rto_event_create(&events[g]), "event");
            void *ev_t0 = nullptr, *ev_t1 = nullptr;   // device time of this band (for the balancer)
            rto_check(rto_event_create_timed(&ev_t0), "event");
            rto_check(rto_event_create_timed(&ev_t1), "event");
            uint8_t* h8 = nullptr;
            float *himg = nullptr, *haux = nullptr;
            if (g == 0) {
                dst_img = rto_context_image(ctx.handle);
                dst_img8 = rto_context_image_rgba8(ctx.handle);
                levels = denoiser.levels();
                rto_check(rto_host_alloc(reinterpret_cast<void**>(&h8), px * 4), "pinned rgba8");
                if (job.write_float) rto_check(rto_host_alloc(reinterpret_cast<void**>(&himg), px * 16), "pinned image");
                host8 = h8; host_img = himg;
            }
            if (direct && !rto_context_image_rgba8(ctx.handle)) throw std::runtime_error(rto_last_error());   // the producing kernels write RGBA8 too
            bar.wait();   // destination pointers published
            if (g != 0 && !direct) {
                rto_check(rto_peer_enable(devices[0]), "peer access to the first GPU");
                rto_check(rto_context_set_image_target(ctx.handle, dst_img, dst_img8), "image target");
            }
            const int halo = job.options.denoise ? 2 + levels : 0;
            rto_camera cam{};
            cam.width = W; cam.height = H; cam.fx = job.fx; cam.fy = job.fy;
            auto one = [&](const Mat43& pose, int64_t warm, int64_t frame, bool timed, size_t out_index) {
                bar.wait();
                const auto t0 = std::chrono::steady_clock::now();
                const int b0 = bounds[g], b1 = bounds[g + 1];
                const int r0 = std::max(0, b0 - halo), r1 = std::min(H, b1 + halo);
                memcpy(cam.c2w, pose.m, sizeof cam.c2w);
                rto_check(rto_context_rng_set_frame(ctx.handle, warm, frame), "rng");
                rto_check(rto_event_record(ev_t0, stream), "event");
                rto_check(rto_render_rect(ctx.handle, tree.device, &cam, &opt, 0, r0, W, r1, stream), "render band");
                if (job.options.denoise) rto_check(rto_denoise_rows(ctx.handle, denoiser.handle(), b0, b1, stream), "denoise band");
                rto_check(rto_event_record(ev_t1, stream), "event");
                if (direct) {
                    // this band's rows go to the host over this GPU's own PCIe link; the frame is complete when every thread is here
                    rto_check(rto_context_read_rows_rgba8(ctx.handle, host8, b0, b1, stream), "read band rgba8");
                    if (host_img) rto_check(rto_context_read_image_rows(ctx.handle, host_img, b0, b1, stream), "read band image");
                    rto_check(rto_synchronize(stream), "sync");
                    bar.wait();
                    if (g == 0) {
                        if (timed) lat_ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
                        if (timed && job.out_dir.size()) write_outputs(job, out_index, nullptr, h8, himg);
                    }
                    rto_check(rto_event_elapsed_ms(ev_t0, ev_t1, &band_ms[g]), "band time");
                    bar.wait();
                    if (g == 0) {
                        if (timed) dev_ms.push_back(*std::max_element(band_ms.begin(), band_ms.end()));
                        if (balance) bounds = rebalance_bands(bounds, band_ms, H);
                    }
                    return;
                }
                rto_check(rto_event_record(events[g], stream), "event");
                bar.wait();   // every band's completion event has been recorded
                if (g == 0) {
                    for (int k = 1; k < n; ++k) rto_check(rto_stream_wait_event(stream, events[k]), "wait band");
                    rto_check(rto_context_mark_image_written(ctx.handle, 1), "mark");
                    rto_check(rto_context_read_image_rgba8(ctx.handle, h8, stream), "read rgba8");
                    if (himg) rto_check(rto_context_read_image(ctx.handle, himg, stream), "read image");
                    rto_check(rto_synchronize(stream), "sync");
                    if (timed) lat_ms.push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
                    if (timed && job.out_dir.size()) write_outputs(job, out_index, nullptr, h8, himg);
                } else {
                    rto_check(rto_synchronize(stream), "sync");
                }
                // outside the timed window: band times -> next frame's boundaries
                rto_check(rto_event_elapsed_ms(ev_t0, ev_t1, &band_ms[g]), "band time");
                bar.wait();
                if (g == 0) {
                    if (timed) dev_ms.push_back(*std::max_element(band_ms.begin(), band_ms.end()));
                    if (balance) bounds = rebalance_bands(bounds, band_ms, H);
                }
            };
            for (int w = 0; w < job.warmup; ++w) one(job.trans[0], 0, w, false, 0);
            for (size_t i = 0; i < frames; ++i) one(job.trans[i], job.warmup, (int64_t)i, true, i);
            bar.wait();
            if (g != 0 && !direct) rto_context_set_image_target(ctx.handle, nullptr, nullptr);
            rto_host_free(h8); rto_host_free(himg); rto_host_free(haux);
            rto_event_destroy(events[g]);
            rto_event_destroy(ev_t0); rto_event_destroy(ev_t1);
            rto_stream_destroy(stream);
        } catch (const std::exception& e) {
            errors[g] = e.what();
            failed = true;
            fprintf(stderr, "tile split, GPU slot %d: %s\n", g, e.what());
            std::exit(134);   // the other threads are parked in the barrier: leave as a whole
        }
    };
    std::vector<std::thread> th;
    for (int g = 0; g < n; ++g) th.emplace_back(worker, g);
    for (auto& t : th) t.join();
    if (lat_ms.empty()) return;
    std::vector<double> s = lat_ms;
    std::sort(s.begin(), s.end());
    double mean = 0;
    for (double v : lat_ms) mean += v;
    mean /= (double)lat_ms.size();
    printf("tile split: %d bands over GPUs [", n);
    for (int g = 0; g < n; ++g) printf("%s%d", g ? "," : "", devices[g]);
    printf("], %dx%d, halo %d rows, %s\n", W, H, job.options.denoise ? 2 + levels : 0,
           direct ? "every GPU copies its own band to the host frame" : "peer-direct stores into the first GPU");
    printf("latency: median %.6f ms, mean %.6f ms, min %.6f ms per frame (rendezvous -> RGBA8 frame on the host, %zu frames)\n",
           s[s.size() / 2], mean, s.front(), lat_ms.size());
    std::vector<double> d = dev_ms;
    std::sort(d.begin(), d.end());
    printf("slowest band, device time: median %.6f ms (bands %s; last boundaries:", d[d.size() / 2], balance ? "balanced by the previous frame's band times" : "of equal height");
    for (int g = 0; g <= n; ++g) printf(" %d", bounds[g]);
    printf(")\n");
    printf("FPS:    %.10f   (1000 / median latency)\n", 1000.0 / s[s.size() / 2]);
}

// one GPU: frames [begin, end) of the job
void run_shard(const Job& job, int device, size_t begin, size_t end, ShardStats& st, bool verbose) {
    if (device >= 0) rto_check(rto_set_device(device), "cudaSetDevice");
    N3Tree tree(job.tree_path);
    if (!tree.is_data_loaded()) std::exit(1);
    if (job.llff) {   // main_headless.cpp:400-405
        tree.use_ndc = true;
        tree.ndc_width = (float)job.width;
        tree.ndc_height = (float)job.height;
        tree.ndc_focal = job.fx;
    }
    tree.sync_ndc();
    // created unconditionally, like the reference (main_headless.cpp:455-456): an empty --ts_module throws
    std::unique_ptr<Denoiser> denoiser = std::make_unique<Denoiser>(job.ts_module);
    if (job.pipe > 1 || job.readback.size()) {
        run_shard_pipelined(job, tree, *denoiser, begin, end, st);
        if (verbose) {
            printf("pipeline: %d frames in flight, %s\n", job.pipe, job.graph ? "one CUDA-graph launch per frame" : "separate launches");
            printf("all:    %.10f ms per frame (wall clock, results on the host)\n", 1e3 * st.wall_s / st.frames);
            printf("FPS:    %.10f\n", st.frames / st.wall_s);
        }
        return;
    }
    // ---- the reference's protocol: one context, one stream, Timer::record (host sync) after every frame
    Camera camera(job.width, job.height, job.fx, job.fy);
    std::vector<float> buf, img;
    std::vector<uint8_t> u8;
    if (job.out_dir.size()) {
        buf.resize((size_t)RenderContext::CHANNELS * job.width * job.height);
        u8.resize((size_t)4 * job.width * job.height);
        if (job.write_float) img.resize((size_t)4 * job.width * job.height);
    }
    void* stream = nullptr;   // the reference creates a blocking stream (cudaStreamDefault); the legacy stream is equivalent here
    RenderContext ctx;
    ctx.offscreen = true;
    ctx.update(job.width, job.height);
    const RenderOptions& options = job.options;

    memcpy(camera.transform, job.trans[0].m, sizeof camera.transform);
    camera._update(false);
    for (int i = 0; i < job.warmup; ++i) {   // warm up, main_headless.cpp:469-479
        launch_renderer(tree, camera, options, ctx, stream, true);
        if (options.denoise) denoiser->denoise(camera, ctx, stream);
        ctx.rng.advance();
    }
    // frame-sharded: the rng is a pure function of the global frame index, identical to the single-GPU sequence
    rto_check(rto_context_rng_set_frame(ctx.handle, job.warmup, (int64_t)begin), "rng");
    ctx.timer().reset(stream);
    const auto t0 = std::chrono::steady_clock::now();
    for (size_t i = begin; i < end; ++i) {
        memcpy(camera.transform, job.trans[i].m, sizeof camera.transform);
        camera._update(false);
        ctx.timer().render_start();
        launch_renderer(tree, camera, options, ctx, stream, true);
        ctx.timer().render_stop();
        if (options.denoise) denoiser->denoise(camera, ctx, stream);
        ctx.timer().record(options.denoise);
        ctx.rng.advance();
        if (!job.out_dir.size()) continue;
        if (job.write_buffer) {
            rto_check(rto_context_read_aux(ctx.handle, buf.data(), stream), "read aux");
        } else {   // (uint8_t)(v * 255) on the device, as the reference does on the host
            rto_check(rto_context_read_image_rgba8(ctx.handle, u8.data(), stream), "read image");
        }
        if (job.write_float) rto_check(rto_context_read_image(ctx.handle, img.data(), stream), "read image");
        rto_check(rto_synchronize(stream), "sync");
        write_outputs(job, i, buf.data(), u8.data(), img.data());
    }
    rto_check(rto_synchronize(stream), "sync");
    st.wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    rto_check(rto_timer_report(ctx.handle, st.ms, &st.frames), "timer");
    st.staged = true;
    if (verbose) {
        ctx.timer().report();
        // wall-clock FPS beside the stage-sum FPS (includes the host gaps and, with -o, the file writes)
        printf("wall:   %.10f ms per frame (wall-clock FPS %.10f)\n", 1e3 * st.wall_s / std::max(1, st.frames), st.frames / st.wall_s);
    }
}

}  // namespace

static int run_main(int argc, char* argv[]) {
    Args args = parse_args(argc, argv);
    if (args.has("help")) { print_help(); return 0; }
    std::string tree_path = args.str("file");
    std::vector<std::string> rest = args.positional;
    if (tree_path.empty() && !rest.empty()) { tree_path = rest[0]; rest.erase(rest.begin()); }
    if (tree_path.empty() || rest.size() != 1) {
        fprintf(stderr, "usage: volrend_headless <tree.npz> <poses> [options]   (--help for the list)\n");
        return 1;
    }
    const fs::path poses_path = rest[0];
    const int device_id = args.i("gpu", -1);

    int width = args.i("width", 800), height = args.i("height", 800);
    float fx = args.f("fx", -1.f);
    if (fx < 0) fx = 1111.11f;
    float fy = args.f("fy", -1.f);
    if (fy < 0) fy = fx;

    Job job;
    std::vector<Mat43>& trans = job.trans;
    std::vector<std::string>& basenames = job.basenames;
    const std::string dataset_type = args.str("dataset", "blender");
    if (dataset_type == "blender") {   // main_headless.cpp:255-272
        const rtohost::Json poses = rtohost::Json::parse(read_text(poses_path.string()));
        const float camera_angle_x = (float)poses.at("camera_angle_x").as_number();
        fx = fy = 0.5f * width / tanf(0.5f * camera_angle_x);
        const rtohost::Json& frames = poses.at("frames");
        for (size_t i = 0; i < frames.size(); ++i) {
            const rtohost::Json& m = frames[i].at("transform_matrix");
            Mat43 t;
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 4; ++c) t.m[c * 3 + r] = (float)m[r][c].as_number();
            trans.push_back(t);
            basenames.push_back("r_" + std::to_string(i));
        }
    } else if (dataset_type == "tt") {   // main_headless.cpp:273-297
        width = 1920;
        height = 1080;
        {
            std::ifstream ifs((poses_path / ".." / "intrinsics.txt").string());
            if (!ifs) { fprintf(stderr, "ERROR: intrin '%s' does not exist\n", (poses_path / ".." / "intrinsics.txt").string().c_str()); return 1; }
            float g;
            ifs >> fx >> g >> g >> g;
            ifs >> g >> fy;
        }
        std::vector<fs::path> files;
        for (const auto& e : fs::directory_iterator(poses_path)) files.push_back(e.path());
        std::sort(files.begin(), files.end());
        for (const fs::path& p : files) {
            const int cnt = read_transform_matrices(p.string(), trans);
            const std::string fname = remove_ext(p.filename().string());
            if (cnt == 1) basenames.push_back(fname);
            else
                for (int i = 0; i < cnt; ++i) {
                    std::string tmp = std::to_string(i);
                    while (tmp.size() < 6) tmp = "0" + tmp;
                    basenames.push_back(fname + "_" + tmp);
                }
        }
    } else if (dataset_type == "llff") {   // main_headless.cpp:298-370
        const rtohost::NpyArray poses = rtohost::npy_load(poses_path.string());
        const size_t pose_cnt = poses.shape[0], per = poses.shape[1];
        auto get = [&](size_t set, size_t idx) -> float {
            const size_t k = set * per + idx;
            return poses.word_size == 4 ? poses.data<float>()[k] : (float)poses.data<double>()[k];
        };
        constexpr int factor = 4;
        width = (int)(get(0, 9) / factor);
        height = (int)(get(0, 4) / factor);
        fx = fy = get(0, 14) / factor;
        float bds_min = 1e9f;
        for (size_t p = 0; p < pose_cnt; ++p) bds_min = std::min(bds_min, get(p, 15));
        for (size_t p = 0; p < pose_cnt; ++p) {
            float t[12];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 4; ++j) t[j * 3 + i] = get(p, (size_t)i * 5 + j);
            Mat43 r;   // temp * cam_trans: col0 <- col1, col1 <- -col0
            for (int i = 0; i < 3; ++i) { r.m[i] = t[3 + i]; r.m[3 + i] = -t[i]; r.m[6 + i] = t[6 + i]; r.m[9 + i] = t[9 + i]; }
            const float scale = 1.0f / (bds_min * 0.75f);
            for (int i = 0; i < 3; ++i) r.m[9 + i] *= scale;
            trans.push_back(r);
        }
        std::string images_dirname = "images";
        if (factor > 1) images_dirname += "_" + std::to_string(factor);
        const fs::path images_path = poses_path.parent_path() / images_dirname;
        for (const auto& e : fs::directory_iterator(images_path)) basenames.push_back(remove_ext(e.path().filename().string()));
        std::sort(basenames.begin(), basenames.end());
    } else {
        fprintf(stderr, "ERROR: unknown dataset type '%s'\n", dataset_type.c_str());
        return 1;
    }

    // camera convention (main_headless.cpp:372-390)
    if (dataset_type == "tt" || args.has("reverse_yz")) {
        puts("INFO: Use OpenCV camera convention\n");
        for (Mat43& t : trans)
            for (int i = 3; i < 9; ++i) t.m[i] = -t.m[i];   // * diag(1,-1,-1,1): flip the up and back columns
    } else if (dataset_type == "llff") {
        puts("INFO: Use LLFF camera convention\n");
        recenter_poses(trans);
    } else {
        puts("INFO: Use NeRF camera convention\n");
    }
    if (trans.empty()) { fputs("WARNING: No camera poses specified, quitting\n", stderr); return 1; }
    while (basenames.size() < trans.size()) basenames.push_back("frame_" + std::to_string(basenames.size()));

    {   // --scale (main_headless.cpp:407-418)
        const float scale = args.f("scale", 1.f);
        if (scale != 1.f) {
            const int ow = width, oh = height;
            width = (int)(width * scale);
            height = (int)(height * scale);
            fx *= (float)width / ow;
            fy *= (float)height / oh;
        }
    }
    {
        const int max_imgs = args.i("max_imgs", 0);
        if (max_imgs > 0 && trans.size() > (size_t)max_imgs) { trans.resize(max_imgs); basenames.resize(max_imgs); }
    }

    job.tree_path = tree_path;
    job.out_dir = args.str("write_images");
    job.ts_module = args.str("ts_module");
    job.width = width; job.height = height; job.fx = fx; job.fy = fy;
    job.llff = dataset_type == "llff";
    job.write_buffer = args.has("write_buffer");
    job.write_float = args.has("write_float");
    job.warmup = args.i("warmup", 100);
    job.pipe = std::max(1, args.i("pipe", 1));
    job.graph = !args.has("no_graph");
    job.writers = args.i("writers", 4);
    job.readback = args.str("readback");
    job.tile_split = args.has("tile_split");
    job.band_readback = args.has("band_readback");
    if (job.tile_split && job.write_buffer) {
        fprintf(stderr, "ERROR: --tile_split assembles the final image only; --write_buffer needs the frame-sharded modes\n");
        return 1;
    }
    if (job.readback.size() && job.readback != "rgba8" && job.readback != "float" && job.readback != "aux") {
        fprintf(stderr, "ERROR: --readback must be rgba8, float or aux\n");
        return 1;
    }
    if (job.out_dir.size()) fs::create_directories(job.out_dir);

    // render options (main_headless.cpp:459-467; opts.cpp:44-66)
    const std::string options_path = args.str("options");
    if (!options_path.empty()) {
        job.options = RenderOptions::from_json(rtohost::Json::parse(read_text(options_path)));
    } else {
        job.options.background_brightness = args.f("bg", 1.0f);
        job.options.step_size = args.f("step_size", 1e-4f);
        job.options.stop_thresh = args.f("stop_thresh", 1e-2f);
        job.options.sigma_thresh = args.f("sigma_thresh", 1e-2f);
    }

    if (args.has("dump_poses")) {   // all camera transforms after the loader + convention, raw float32 [n][12] (parity tests)
        std::ofstream pf(args.str("dump_poses"), std::ios::binary);
        for (const Mat43& t : trans) pf.write(reinterpret_cast<const char*>(t.m), sizeof t.m);
    }
    if (args.has("dry_run")) {
        rtohost::npz_t z = rtohost::npz_load(tree_path);
        N3Tree::HostArrays h;
        N3Tree::decode_npz(z, h, /*decode_on_host=*/true);
        printf("{\"poses\": %zu, \"width\": %d, \"height\": %d, \"fx\": %.9g, \"fy\": %.9g, \"spp\": %d, \"denoise\": %s, "
               "\"step_size\": %.9g, \"sigma_thresh\": %.9g, \"background\": %.9g, \"capacity\": %d, \"data_dim\": %d, "
               "\"data_format\": \"%s\", \"child_fnv\": \"%016llx\", \"data_fnv\": \"%016llx\", \"pose0\": [",
               trans.size(), width, height, fx, fy, job.options.spp, job.options.denoise ? "true" : "false", job.options.step_size,
               job.options.sigma_thresh, job.options.background_brightness, h.capacity, h.data_dim, h.data_format.to_string().c_str(),
               (unsigned long long)fnv64(h.child, h.n_child * 4), (unsigned long long)fnv64(h.data, h.n_child * (size_t)h.data_dim * 2));
        for (int i = 0; i < 12; ++i) printf("%s%.9g", i ? ", " : "", trans[0].m[i]);
        printf("], \"pose_last\": [");
        for (int i = 0; i < 12; ++i) printf("%s%.9g", i ? ", " : "", trans.back().m[i]);
        printf("], \"basename0\": \"%s\"}\n", basenames[0].c_str());
        return 0;
    }

    int num_gpus = args.i("num_gpus", 1);
    if (num_gpus < 1) num_gpus = 1;
    std::vector<int> devices;
    if (args.has("gpu_list")) {   // one shard per entry; an id may repeat (two host threads driving one GPU)
        std::stringstream ss(args.str("gpu_list"));
        std::string tok;
        while (std::getline(ss, tok, ',')) if (tok.size()) devices.push_back(atoi(tok.c_str()));
        if (devices.empty()) { fprintf(stderr, "ERROR: empty --gpu_list\n"); return 1; }
        num_gpus = (int)devices.size();
    } else {
        for (int g = 0; g < num_gpus; ++g) devices.push_back(num_gpus == 1 ? device_id : g);
    }
    try {
        if (job.tile_split) {
            run_tile_split(job, devices);
        } else if (num_gpus == 1) {
            ShardStats st;
            run_shard(job, devices[0], 0, trans.size(), st, true);
        } else {
            // frame sharding: contiguous slices of the pose list, one host thread and one replica per shard, no collective
            std::vector<std::thread> th;
            std::vector<ShardStats> st(num_gpus);
            const size_t n = trans.size();
            const auto t0 = std::chrono::steady_clock::now();
            for (int g = 0; g < num_gpus; ++g) {
                const size_t b = n * g / num_gpus, e = n * (g + 1) / num_gpus;
                th.emplace_back([&, g, b, e]() { run_shard(job, devices[g], b, e, st[g], false); });
            }
            for (auto& t : th) t.join();
            const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            float agg[3] = {0, 0, 0};
            int tot = 0;
            double slowest = 0.0;
            bool staged = true;
            for (int g = 0; g < num_gpus; ++g) {
                for (int k = 0; k < 3; ++k) agg[k] += st[g].ms[k] * st[g].frames;
                tot += st[g].frames;
                slowest = std::max(slowest, st[g].wall_s);
                staged = staged && st[g].staged;
            }
            if (staged) {
                float all = 0.f;
                const char* names[3] = {"render", "torch", "filter"};
                for (int k = 0; k < 3; ++k) { printf("%s: %.10f ms per frame\n", names[k], agg[k] / tot); all += agg[k] / tot; }
                printf("all:    %.10f ms per frame\n", all);
                printf("FPS:    %.10f   (per GPU; %d shards, %d frames, wall %.3f s incl. load + warm-up)\n", 1000.f / all, num_gpus, tot, wall);
                printf("aggregate FPS: %.10f\n", num_gpus * 1000.f / all);
            } else {
                printf("pipeline: %d frames in flight per shard, %s\n", job.pipe, job.graph ? "one CUDA-graph launch per frame" : "separate launches");
                printf("all:    %.10f ms per frame (wall clock of the slowest shard, results on the host)\n", 1e3 * slowest / tot);
            }
            printf("aggregate wall FPS: %.10f   (%d frames / slowest shard's timed loop %.4f s)\n", tot / slowest, tot, slowest);
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "terminate called after throwing an instance of 'std::runtime_error'\n  what():  %s\n", e.what());
        return 134;
    }
    return 0;
}

// Logic errors are C++ exceptions in the reference and end the process through std::terminate (exit status 134); here they
// are reported the same way, without the core dump.
int main(int argc, char* argv[]) {
    try {
        return run_main(argc, argv);
    } catch (const std::exception& e) {
        fprintf(stderr, "terminate called after throwing an instance of 'std::runtime_error'\n  what():  %s\n", e.what());
        return 134;
    }
}
