// TEST INFRASTRUCTURE — not part of the product.
//
// Minimal stand-in for the reference's renderer/include/volrend/n3tree.hpp, placed EARLIER on the include
// path when oracle/ref_cpu_shim.cpp host-compiles the reference's device headers.  The real header drags in
// cnpy + the CUDA allocation code; internal::TreeSpec(tree, /*cpu=*/true) (data_spec.hpp:39-52) only needs
// the members below, so this stand-in lets the reference's own trace_ray / query_single_from_root /
// maybe_precalc_basis run on host arrays we hand it.  It restates no algorithm.
#pragma once
#include <array>
#include <cstdint>
#include <cuda_fp16.h>
#include "volrend/common.hpp"
#include "volrend/data_format.hpp"

namespace volrend {
struct HostArrayView {
    const void* ptr = nullptr;
    template <class T> const T* data() const { return reinterpret_cast<const T*>(ptr); }
};
struct N3Tree {
    int N = 2;
    int data_dim = 0;
    DataFormat data_format;
    int capacity = 0;
    std::array<float, 3> scale;
    std::array<float, 3> offset;
    bool use_ndc = false;
    float ndc_width = 0, ndc_height = 0, ndc_focal = 0;
    mutable struct {
        __half* data = nullptr;
        int32_t* child = nullptr;
        float* offset = nullptr;
        float* scale = nullptr;
        float* extra = nullptr;
    } device;
    HostArrayView data_, child_, extra_;
};
}  // namespace volrend
