// TEST INFRASTRUCTURE — not part of the product.
//
// Pins the host-side llff pose recentring against the reference's OWN code: this translation unit includes the
// reference's renderer/main_headless.cpp where it lies (found through -I /root/reference/renderer; nothing is copied),
// with its main() renamed, so that the functions of its anonymous namespace — _recenter_poses, _poses_avg, _viewmatrix
// (main_headless.cpp:107-189), which call glm::inverse / glm mat4*mat4 — can be called directly.
//
// usage: ref_pose_dump in.bin out.bin     in/out = float32 [n][12], column-major 4x3 poses (glm::mat4x3 memory order),
//                                          in = the poses as they are right before main_headless.cpp:385-387 recentres them.
// Built by oracle/build_ref.sh into oracle/_ref/ref_pose_dump (links the same reference objects as the reference CLI).
#define main ref_main_headless_renamed
#include "main_headless.cpp"
#undef main

#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>

int main(int argc, char** argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]);
        return 2;
    }
    std::ifstream in(argv[1], std::ios::binary);
    std::vector<char> raw((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    const size_t n = raw.size() / (12 * sizeof(float));
    std::vector<glm::mat4x3> trans(n);
    for (size_t i = 0; i < n; ++i) memcpy(&trans[i][0][0], raw.data() + i * 12 * sizeof(float), 12 * sizeof(float));
    _recenter_poses(trans);
    std::ofstream out(argv[2], std::ios::binary);
    for (size_t i = 0; i < n; ++i) out.write(reinterpret_cast<const char*>(&trans[i][0][0]), 12 * sizeof(float));
    return 0;
}
