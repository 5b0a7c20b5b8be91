// TEST INFRASTRUCTURE — not part of the product.
//
// A small driver of OUR OWN that links against the UNMODIFIED reference objects (built by
// oracle/build_ref.sh from /root/reference in place) and calls the reference's public C++ API
//   volrend::N3Tree(path)                       renderer/include/volrend/n3tree.hpp:24-31
//   volrend::launch_renderer(...)               renderer/include/volrend/cuda/renderer_kernel.hpp:11-16
//   volrend::Denoiser(ts).denoise(cam,ctx,s)    renderer/include/volrend/denoiser/denoiser.hpp:11-21
// with exactly the frame protocol of renderer/main_headless.cpp:441-506 (blocking stream, N warm-up
// frames on pose 0 each followed by ctx.rng.advance(), then one advance per pose).  Unlike the reference
// CLI (whose PNG writer is a no-op without libpng, src/imwrite.cpp:80-85) it dumps, per frame, the aux
// buffer [8][H][W] fp32 and the final float4 image [H][W][4] so the GPU parity tests can compare both.
//
// usage: ref_driver tree.npz poses.bin ts_module W H fx fy spp denoise(0/1) out_dir|- [nframes] [warmup] [bg]
//   poses.bin = float32 [n][12], each pose the column-major 4x3 c2w (right, up, back, centre) that
//   Camera::_update uploads (src/camera.cpp:47-76).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "volrend/common.hpp"
#include "volrend/denoiser/denoiser.hpp"
#include "volrend/n3tree.hpp"
// clang-format off
#include "volrend/cuda/common.cuh"
#include "volrend/cuda/renderer_kernel.hpp"
// clang-format on

using namespace volrend;

static void dump(const std::string& path, const void* p, size_t n) {
    std::ofstream f(path, std::ios::out | std::ios::binary);
    f.write((const char*)p, n);
}

int main(int argc, char** argv) {
    if (argc < 11) {
        fprintf(stderr, "usage: %s tree.npz poses.bin ts W H fx fy spp denoise out_dir|- [nframes] [warmup] [bg]\n", argv[0]);
        return 2;
    }
    const std::string tree_path = argv[1], poses_path = argv[2], ts_path = argv[3];
    const int W = atoi(argv[4]), H = atoi(argv[5]);
    const float fx = atof(argv[6]), fy = atof(argv[7]);
    const int spp = atoi(argv[8]);
    const bool denoise = atoi(argv[9]) != 0;
    const std::string out_dir = argv[10];
    int nframes = argc > 11 ? atoi(argv[11]) : 0;
    const int warmup = argc > 12 ? atoi(argv[12]) : 100;
    const float bg = argc > 13 ? atof(argv[13]) : 1.0f;
    const bool do_dump = out_dir != "-";

    std::vector<float> poses;
    {
        std::ifstream f(poses_path, std::ios::binary | std::ios::ate);
        if (!f) { fprintf(stderr, "cannot open %s\n", poses_path.c_str()); return 1; }
        size_t n = f.tellg();
        f.seekg(0);
        poses.resize(n / 4);
        f.read((char*)poses.data(), n);
    }
    const int npose = (int)(poses.size() / 12);
    if (nframes <= 0 || nframes > npose) nframes = npose;

    N3Tree tree(tree_path);
    Camera camera(W, H, fx, fy);

    cudaArray_t array;
    cudaStream_t stream;
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float4>();
    cuda(MallocArray(&array, &desc, W, H));
    cuda(StreamCreateWithFlags(&stream, cudaStreamDefault));

    RenderContext ctx;
    ctx.offscreen = true;
    ctx.update(array, nullptr, W, H);

    std::unique_ptr<Denoiser> denoiser = std::make_unique<Denoiser>(ts_path);

    RenderOptions options;
    options.background_brightness = bg;
    options.spp = spp;
    options.denoise = denoise;

    auto set_pose = [&](int i) {
        const float* p = &poses[12 * (size_t)i];
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 3; ++r) camera.transform[c][r] = p[c * 3 + r];
        camera._update(false);
    };

    set_pose(0);
    for (int i = 0; i < warmup; ++i) {
        launch_renderer(tree, camera, options, ctx, stream, true);
        if (options.denoise) denoiser->denoise(camera, ctx, stream);
        ctx.rng.advance();
    }
    ctx.timer().reset(stream);

    std::vector<float> aux((size_t)8 * W * H), img((size_t)4 * W * H);
    for (int i = 0; i < nframes; ++i) {
        set_pose(i);
        ctx.timer().render_start();
        launch_renderer(tree, camera, options, ctx, stream, true);
        ctx.timer().render_stop();
        if (options.denoise) denoiser->denoise(camera, ctx, stream);
        ctx.timer().record(options.denoise);
        ctx.rng.advance();
        if (do_dump) {
            cuda(Memcpy(aux.data(), ctx.aux_buffer, aux.size() * 4, cudaMemcpyDeviceToHost));
            dump(out_dir + "/aux_" + std::to_string(i) + ".bin", aux.data(), aux.size() * 4);
            cuda(Memcpy2DFromArray(img.data(), sizeof(float4) * W, array, 0, 0, sizeof(float4) * W, H,
                                   cudaMemcpyDeviceToHost));
            dump(out_dir + "/img_" + std::to_string(i) + ".bin", img.data(), img.size() * 4);
        }
    }
    ctx.timer().report();
    ctx.freeResource();
    cuda(FreeArray(array));
    cuda(StreamDestroy(stream));
    return 0;
}
