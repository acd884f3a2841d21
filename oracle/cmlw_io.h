// CMLW container: a flat list of named n-d arrays (little endian, C order).
// TEST INFRASTRUCTURE (used by oracle/ref_driver.cpp); the Python twin is libcml_b200/cmlw.py.
//   "CMLW0001" | u32 n | n x { u32 name_len | name | u32 dtype | u32 ndim | u64 dims[ndim] | data }
//   dtype: 0=f32 1=f64 2=i32 3=u8 4=i64
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace cmlw {

enum DType : uint32_t { F32 = 0, F64 = 1, I32 = 2, U8 = 3, I64 = 4 };
static inline size_t dtype_size(uint32_t d) { static const size_t s[5] = {4, 8, 4, 1, 8}; return s[d]; }

struct Array {
    uint32_t dtype = F32;
    std::vector<uint64_t> dims;
    std::vector<uint8_t> data;
    size_t count() const { size_t n = 1; for (auto d : dims) n *= d; return n; }
    template <typename T> const T *as() const { return reinterpret_cast<const T *>(data.data()); }
    template <typename T> T *as() { return reinterpret_cast<T *>(data.data()); }
};

struct File {
    std::vector<std::string> order;
    std::map<std::string, Array> arrays;

    template <typename T> static uint32_t code();

    template <typename T> void put(const std::string &name, const T *p, std::vector<uint64_t> dims) {
        Array a; a.dtype = code<T>(); a.dims = dims;
        a.data.resize(a.count() * sizeof(T));
        if (a.count()) memcpy(a.data.data(), p, a.data.size());
        if (!arrays.count(name)) order.push_back(name);
        arrays[name] = std::move(a);
    }
    template <typename T> void put(const std::string &name, const std::vector<T> &v, std::vector<uint64_t> dims) {
        size_t n = 1; for (auto d : dims) n *= d;
        if (n != v.size()) { fprintf(stderr, "cmlw: dims mismatch for %s (%zu vs %zu)\n", name.c_str(), n, v.size()); abort(); }
        put<T>(name, v.data(), dims);
    }
    template <typename T> void put1(const std::string &name, const std::vector<T> &v) { put<T>(name, v.data(), {(uint64_t) v.size()}); }
    template <typename T> void scalar(const std::string &name, T v) { put<T>(name, &v, {1}); }

    bool has(const std::string &name) const { return arrays.count(name) > 0; }
    const Array &get(const std::string &name) const {
        auto it = arrays.find(name);
        if (it == arrays.end()) { fprintf(stderr, "cmlw: missing array %s\n", name.c_str()); abort(); }
        return it->second;
    }

    bool save(const std::string &path) const {
        FILE *f = fopen(path.c_str(), "wb");
        if (!f) return false;
        fwrite("CMLW0001", 1, 8, f);
        uint32_t n = (uint32_t) order.size();
        fwrite(&n, 4, 1, f);
        for (auto &name : order) {
            const Array &a = arrays.at(name);
            uint32_t nl = (uint32_t) name.size(), nd = (uint32_t) a.dims.size();
            fwrite(&nl, 4, 1, f); fwrite(name.data(), 1, nl, f);
            fwrite(&a.dtype, 4, 1, f); fwrite(&nd, 4, 1, f);
            fwrite(a.dims.data(), 8, nd, f);
            if (!a.data.empty()) fwrite(a.data.data(), 1, a.data.size(), f);
        }
        fclose(f);
        return true;
    }

    bool load(const std::string &path) {
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) return false;
        char magic[8];
        if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "CMLW0001", 8)) { fclose(f); return false; }
        uint32_t n = 0;
        if (fread(&n, 4, 1, f) != 1) { fclose(f); return false; }
        for (uint32_t i = 0; i < n; i++) {
            uint32_t nl, nd; Array a;
            if (fread(&nl, 4, 1, f) != 1) { fclose(f); return false; }
            std::string name(nl, '\0');
            if (fread(name.data(), 1, nl, f) != nl) { fclose(f); return false; }
            if (fread(&a.dtype, 4, 1, f) != 1 || fread(&nd, 4, 1, f) != 1) { fclose(f); return false; }
            a.dims.resize(nd);
            if (nd && fread(a.dims.data(), 8, nd, f) != nd) { fclose(f); return false; }
            a.data.resize(a.count() * dtype_size(a.dtype));
            if (!a.data.empty() && fread(a.data.data(), 1, a.data.size(), f) != a.data.size()) { fclose(f); return false; }
            order.push_back(name);
            arrays[name] = std::move(a);
        }
        fclose(f);
        return true;
    }
};

template <> inline uint32_t File::code<float>() { return F32; }
template <> inline uint32_t File::code<double>() { return F64; }
template <> inline uint32_t File::code<int32_t>() { return I32; }
template <> inline uint32_t File::code<uint8_t>() { return U8; }
template <> inline uint32_t File::code<int64_t>() { return I64; }

}  // namespace cmlw
