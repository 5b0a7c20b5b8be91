// npz.hpp — minimal .npy / .npz reader for the host side (replaces renderer/3rdparty/cnpy: npz_load / npy_load,
// cnpy.cpp:230-369).  Handles what numpy.savez / savez_compressed and svox write: zip local headers + central
// directory, zip64 extensions, stored (method 0) and deflate (method 8, via zlib) members, .npy format 1.0/2.0/3.0.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace rtohost {

struct NpyArray {
    std::vector<size_t> shape;
    size_t word_size = 0;      // bytes per element (for '<U3': 4 * 3 = 12)
    char kind = '?';           // numpy kind character: f, i, u, b, U, S
    bool fortran_order = false;
    std::vector<unsigned char> bytes;

    size_t num_vals() const {
        size_t n = 1;
        for (size_t s : shape) n *= s;
        return n;
    }
    template <class T>
    const T* data() const { return reinterpret_cast<const T*>(bytes.data()); }
    // '<U..' string as ASCII (the reference takes every 4th byte: n3tree.cpp:231-239)
    std::string as_string() const;
    // scalar readers that accept either width (the reference hard-codes int64 / double / float per key)
    double scalar_as_double() const;
};

using npz_t = std::map<std::string, NpyArray>;

NpyArray parse_npy(const unsigned char* buf, size_t len);
NpyArray npy_load(const std::string& path);
npz_t npz_load(const std::string& path);
// the same from a memory image of the file (cnpy::npz_load_mem, used by N3Tree::open_mem); `what` names it in error messages
npz_t npz_load_mem(const unsigned char* buf, size_t len, const std::string& what = "<memory>");

}  // namespace rtohost
