// BCSD: tiny named-array container for scene / state / dump files (shared by the
// headless driver, the oracle, the reference harness and pytest).
//
//   file   := "BCSD1\0\0\0" record*
//   record := u32 name_len, name bytes, u32 dtype, u64 count, payload
//   dtype  := 0 f32 | 1 i32 | 2 u32 | 3 f64 | 4 i64
//
// Python twin: tests/bcsd.py
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace bcsd {

enum DType : uint32_t { F32 = 0, I32 = 1, U32 = 2, F64 = 3, I64 = 4 };

inline size_t dtype_size(uint32_t d) { return (d == F64 || d == I64) ? 8 : 4; }

struct Array {
    uint32_t dtype = F32;
    std::vector<unsigned char> bytes;
    size_t count() const { return bytes.size() / dtype_size(dtype); }
    template <class T> const T* as() const { return reinterpret_cast<const T*>(bytes.data()); }
    template <class T> T* as() { return reinterpret_cast<T*>(bytes.data()); }
};

class Writer {
    FILE* f_ = nullptr;
public:
    explicit Writer(const std::string& path) {
        f_ = std::fopen(path.c_str(), "wb");
        if (!f_) throw std::runtime_error("bcsd: cannot open for write: " + path);
        const char magic[8] = {'B', 'C', 'S', 'D', '1', 0, 0, 0};
        std::fwrite(magic, 1, 8, f_);
    }
    ~Writer() { if (f_) std::fclose(f_); }
    Writer(const Writer&) = delete;
    Writer& operator=(const Writer&) = delete;

    void put_raw(const std::string& name, uint32_t dtype, const void* data, uint64_t count) {
        uint32_t nl = (uint32_t)name.size();
        std::fwrite(&nl, 4, 1, f_);
        std::fwrite(name.data(), 1, nl, f_);
        std::fwrite(&dtype, 4, 1, f_);
        std::fwrite(&count, 8, 1, f_);
        if (count) std::fwrite(data, dtype_size(dtype), count, f_);
    }
    void put(const std::string& n, const float* p, uint64_t c) { put_raw(n, F32, p, c); }
    void put(const std::string& n, const int32_t* p, uint64_t c) { put_raw(n, I32, p, c); }
    void put(const std::string& n, const uint32_t* p, uint64_t c) { put_raw(n, U32, p, c); }
    void put(const std::string& n, const double* p, uint64_t c) { put_raw(n, F64, p, c); }
    void put(const std::string& n, const int64_t* p, uint64_t c) { put_raw(n, I64, p, c); }
    template <class T> void put(const std::string& n, const std::vector<T>& v) { put(n, v.data(), v.size()); }
    void put_scalar(const std::string& n, int32_t v) { put(n, &v, 1); }
    void put_scalar(const std::string& n, float v) { put(n, &v, 1); }
};

inline std::map<std::string, Array> read_all(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("bcsd: cannot open for read: " + path);
    char magic[8];
    if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "BCSD1", 5) != 0) {
        std::fclose(f);
        throw std::runtime_error("bcsd: bad magic in " + path);
    }
    std::map<std::string, Array> out;
    for (;;) {
        uint32_t nl;
        if (std::fread(&nl, 4, 1, f) != 1) break;
        std::string name(nl, '\0');
        if (std::fread(name.data(), 1, nl, f) != nl) throw std::runtime_error("bcsd: truncated name");
        Array a;
        uint64_t count;
        if (std::fread(&a.dtype, 4, 1, f) != 1 || std::fread(&count, 8, 1, f) != 1)
            throw std::runtime_error("bcsd: truncated header");
        a.bytes.resize(count * dtype_size(a.dtype));
        if (count && std::fread(a.bytes.data(), 1, a.bytes.size(), f) != a.bytes.size())
            throw std::runtime_error("bcsd: truncated payload for " + name);
        out.emplace(std::move(name), std::move(a));
    }
    std::fclose(f);
    return out;
}

template <class T>
inline std::vector<T> get_vec(const std::map<std::string, Array>& m, const std::string& name) {
    auto it = m.find(name);
    if (it == m.end()) throw std::runtime_error("bcsd: missing array " + name);
    if (dtype_size(it->second.dtype) != sizeof(T)) throw std::runtime_error("bcsd: dtype size mismatch " + name);
    const T* p = it->second.as<T>();
    return std::vector<T>(p, p + it->second.count());
}

inline bool has(const std::map<std::string, Array>& m, const std::string& name) { return m.count(name) != 0; }

}  // namespace bcsd
