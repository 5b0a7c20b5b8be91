// rto_tree.h — interface between the C-ABI layer (rto_api.cu) and the device tree loader (rto_tree.cu).  Not installed.
#pragma once
#include <cuda_fp16.h>

#include <cstdint>
#include <string>

namespace rto {

// Host arrays of one tree.npz, dense (`data_f16` != nullptr) or quantised (renderer/src/n3tree.cpp:279-340)
struct TreeSource {
    const int32_t* child = nullptr;      // [capacity][8] relative node offsets, 0 = leaf
    int64_t capacity = 0;
    int data_dim = 0;
    const void* data_f16 = nullptr;      // dense: fp16 [capacity*8][data_dim], sigma last
    const void* quant_colors = nullptr;  // quantised: fp16 [n_q][65536][3] codebooks
    const uint16_t* quant_map = nullptr; //            u16 [n_q][capacity*8] codebook indices
    const void* sigma_f16 = nullptr;     //            fp16 [capacity*8]
    const void* retained_f16 = nullptr;  //            fp16 [n_ret][capacity*8][3] uncompressed leading basis functions
    int n_q = 0, n_ret = 0;
};

struct TreeBuilt {   // device planes, owned by the caller after a successful build
    uint32_t* nodes = nullptr;
    __half* payload = nullptr;
    uint32_t* grid_top = nullptr;
    uint32_t* grid_bricks = nullptr;
    uint8_t* grid_bricks8 = nullptr;   // byte plane of the bricks (rto_ray.cuh GridDev)
    uint32_t* grid_leaf_top = nullptr;     // leaf-id planes (rto_ray.cuh GridDev::leaf_top / leaf_bricks); nullptr: not built
    uint32_t* grid_leaf_bricks = nullptr;
    uint32_t* grid_top_m = nullptr;        // march table of the fused-index marcher (rto_ray.cuh FusedIdx); nullptr: not built
    int grid_K = 0;
    int64_t n_bricks = 0;
    int64_t n_leaves = 0;
    int max_depth = 0;
    int stride = 0;   // payload halfs per entry
};

// Upload + build every plane on the current device.  Returns an rto_status; on failure `err` holds the message and
// nothing stays allocated.  `launches` is incremented by the number of kernels launched.
int build_tree_device(const TreeSource& src, TreeBuilt& out, std::string& err, int64_t* launches);
void tree_built_free(TreeBuilt& b);

}  // namespace rto
