// npz.cpp — see npz.hpp
#include "npz.hpp"

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>

namespace rtohost {

namespace {
uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint64_t rd64(const unsigned char* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

std::vector<unsigned char> read_file(const std::string& path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + path);
    const std::streamsize n = f.tellg();
    f.seekg(0);
    std::vector<unsigned char> buf((size_t)n);
    if (n > 0 && !f.read(reinterpret_cast<char*>(buf.data()), n)) throw std::runtime_error("cannot read " + path);
    return buf;
}
}  // namespace

std::string NpyArray::as_string() const {
    std::string s;
    if (kind == 'U') {
        for (size_t i = 0; i + 3 < bytes.size(); i += 4)
            if (bytes[i]) s.push_back((char)bytes[i]);
    } else {
        for (unsigned char c : bytes)
            if (c) s.push_back((char)c);
    }
    return s;
}

double NpyArray::scalar_as_double() const {
    if (bytes.size() < word_size || word_size == 0) throw std::runtime_error("empty npy scalar");
    if (kind == 'f') {
        if (word_size == 8) return *data<double>();
        if (word_size == 4) return *data<float>();
    } else if (kind == 'i') {
        if (word_size == 8) return (double)*data<int64_t>();
        if (word_size == 4) return (double)*data<int32_t>();
    } else if (kind == 'u') {
        if (word_size == 8) return (double)*data<uint64_t>();
        if (word_size == 4) return (double)*data<uint32_t>();
    }
    throw std::runtime_error("unsupported npy scalar type");
}

NpyArray parse_npy(const unsigned char* buf, size_t len) {
    if (len < 10 || memcmp(buf, "\x93NUMPY", 6) != 0) throw std::runtime_error("not an npy stream");
    const int major = buf[6];
    size_t hlen, hoff;
    if (major == 1) { hlen = rd16(buf + 8); hoff = 10; }
    else { hlen = rd32(buf + 8); hoff = 12; }
    if (hoff + hlen > len) throw std::runtime_error("truncated npy header");
    const std::string h(reinterpret_cast<const char*>(buf + hoff), hlen);
    NpyArray a;
    // 'descr': '<f2'
    const size_t npos = std::string::npos;
    size_t p = h.find("'descr'");
    if (p == npos) throw std::runtime_error("npy header: no descr");
    const size_t colon = h.find(':', p);
    p = colon == npos ? npos : h.find('\'', colon);
    const size_t q = p == npos ? npos : h.find('\'', p + 1);
    if (q == npos) throw std::runtime_error("npy header: bad descr");
    const std::string descr = h.substr(p + 1, q - p - 1);
    if (descr.size() < 2) throw std::runtime_error("npy header: bad descr");
    size_t k = 0;
    if (descr[0] == '<' || descr[0] == '|' || descr[0] == '=') k = 1;
    else if (descr[0] == '>') throw std::runtime_error("big-endian npy not supported");
    a.kind = descr[k];
    const size_t width = (size_t)atoi(descr.c_str() + k + 1);
    a.word_size = a.kind == 'U' ? 4 * width : width;
    p = h.find("'fortran_order'");
    if (p != npos) {
        const size_t c2 = h.find(':', p);
        a.fortran_order = c2 != npos && c2 + 6 <= h.size() && h.compare(c2 + 2, 4, "True") == 0;
    }
    p = h.find("'shape'");
    if (p == npos) throw std::runtime_error("npy header: no shape");
    p = h.find('(', p);
    const size_t e = p == npos ? npos : h.find(')', p);
    if (e == npos) throw std::runtime_error("npy header: bad shape");
    const std::string sh = h.substr(p + 1, e - p - 1);
    size_t i = 0;
    while (i < sh.size()) {
        while (i < sh.size() && (sh[i] == ' ' || sh[i] == ',')) ++i;
        if (i >= sh.size()) break;
        a.shape.push_back((size_t)strtoull(sh.c_str() + i, nullptr, 10));
        while (i < sh.size() && sh[i] != ',') ++i;
    }
    // element count and byte size with overflow checks (a crafted shape must not wrap around the bounds test below)
    size_t nbytes = a.word_size;
    for (size_t d : a.shape) {
        if (d != 0 && nbytes > SIZE_MAX / d) throw std::runtime_error("npy header: shape overflows");
        nbytes *= d;
    }
    if (nbytes > len || hoff + hlen > len - nbytes) throw std::runtime_error("truncated npy payload");
    a.bytes.assign(buf + hoff + hlen, buf + hoff + hlen + nbytes);
    return a;
}

NpyArray npy_load(const std::string& path) {
    const auto buf = read_file(path);
    return parse_npy(buf.data(), buf.size());
}

npz_t npz_load(const std::string& path) {
    const auto buf = read_file(path);
    return npz_load_mem(buf.data(), buf.size(), path);
}

npz_t npz_load_mem(const unsigned char* buf, size_t n, const std::string& path) {
    if (n < 22) throw std::runtime_error("npz too small: " + path);
    // end-of-central-directory record (search backwards), zip64 locator if present
    size_t eocd = std::string::npos;
    for (size_t i = n - 22;; --i) {
        if (rd32(&buf[i]) == 0x06054b50u) { eocd = i; break; }
        if (i == 0 || n - i > 22 + 65535) break;
    }
    if (eocd == std::string::npos) throw std::runtime_error("npz: no end-of-central-directory in " + path);
    uint64_t n_entries = rd16(&buf[eocd + 10]);
    uint64_t cd_off = rd32(&buf[eocd + 16]);
    if (eocd >= 20 && rd32(&buf[eocd - 20]) == 0x07064b50u) {  // zip64 EOCD locator
        const uint64_t z64 = rd64(&buf[eocd - 20 + 8]);
        if (z64 <= n && 56 <= n - z64 && rd32(&buf[z64]) == 0x06064b50u) {
            n_entries = rd64(&buf[z64 + 32]);
            cd_off = rd64(&buf[z64 + 48]);
        }
    }
    npz_t out;
    if (cd_off > n) throw std::runtime_error("npz: bad central directory offset");
    size_t p = (size_t)cd_off;
    for (uint64_t e = 0; e < n_entries; ++e) {
        if (p > n || 46 > n - p || rd32(&buf[p]) != 0x02014b50u) throw std::runtime_error("npz: bad central directory");
        const uint16_t method = rd16(&buf[p + 10]);
        uint64_t csize = rd32(&buf[p + 20]), usize = rd32(&buf[p + 24]);
        const uint16_t nlen = rd16(&buf[p + 28]), xlen = rd16(&buf[p + 30]), clen = rd16(&buf[p + 32]);
        uint64_t lho = rd32(&buf[p + 42]);
        // name, extra field and comment must lie inside the buffer before any of them is read
        if ((size_t)nlen + xlen + clen > n - p - 46) throw std::runtime_error("npz: central directory entry runs past the end");
        std::string name(reinterpret_cast<const char*>(&buf[p + 46]), nlen);
        // zip64 extra field: 8-byte values, present only for the 32-bit fields that hold 0xffffffff, in this order
        size_t x = p + 46 + nlen;
        const size_t xend = x + xlen;
        while (x + 4 <= xend) {
            const uint16_t id = rd16(&buf[x]), sz = rd16(&buf[x + 2]);
            if ((size_t)sz > xend - x - 4) throw std::runtime_error("npz: extra field runs past its record");
            if (id == 0x0001) {
                size_t y = x + 4;
                const size_t yend = x + 4 + sz;
                auto take64 = [&](uint64_t& v) {
                    if (8 > yend - y) throw std::runtime_error("npz: short zip64 extra field");
                    v = rd64(&buf[y]);
                    y += 8;
                };
                if (usize == 0xffffffffu) take64(usize);
                if (csize == 0xffffffffu) take64(csize);
                if (lho == 0xffffffffu) take64(lho);
            }
            x += 4 + sz;
        }
        p = xend + clen;
        if (lho > n || 30 > n - lho || rd32(&buf[lho]) != 0x04034b50u) throw std::runtime_error("npz: bad local header");
        const uint64_t doff64 = lho + 30 + rd16(&buf[lho + 26]) + rd16(&buf[lho + 28]);
        if (doff64 > n || csize > n - doff64) throw std::runtime_error("npz: truncated member " + name);
        const size_t doff = (size_t)doff64;
        if (name.size() > 4 && name.substr(name.size() - 4) == ".npy") name.resize(name.size() - 4);
        if (method == 0) {
            out[name] = parse_npy(&buf[doff], (size_t)csize);
        } else if (method == 8) {
            if (usize > ((uint64_t)1 << 40)) throw std::runtime_error("npz: implausible member size for " + name);
            std::vector<unsigned char> raw((size_t)usize);
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -MAX_WBITS) != Z_OK) throw std::runtime_error("zlib init failed");
            size_t in_done = 0, out_done = 0;
            int rc = Z_OK;
            while (rc != Z_STREAM_END) {   // feed in < 4 GiB slices (zlib counters are 32-bit)
                const size_t in_chunk = std::min<size_t>((size_t)csize - in_done, (size_t)1 << 30);
                const size_t out_chunk = std::min<size_t>(raw.size() - out_done, (size_t)1 << 30);
                zs.next_in = const_cast<unsigned char*>(&buf[doff + in_done]);
                zs.avail_in = (uInt)in_chunk;
                zs.next_out = raw.data() + out_done;
                zs.avail_out = (uInt)out_chunk;
                rc = inflate(&zs, Z_NO_FLUSH);
                in_done += in_chunk - zs.avail_in;
                out_done += out_chunk - zs.avail_out;
                if (rc != Z_OK && rc != Z_STREAM_END) { inflateEnd(&zs); throw std::runtime_error("npz: inflate failed for " + name); }
                if (rc == Z_OK && in_chunk == 0 && out_chunk == 0) break;
            }
            inflateEnd(&zs);
            if (out_done != raw.size()) throw std::runtime_error("npz: short inflate for " + name);
            out[name] = parse_npy(raw.data(), raw.size());
        } else {
            throw std::runtime_error("npz: unsupported compression method in " + name);
        }
    }
    return out;
}

}  // namespace rtohost
