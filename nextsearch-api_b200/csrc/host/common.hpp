// Shared host helpers: error slot, little-endian raw IO, counter-based RNG.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <new>
#include <string>
#include <vector>

namespace nsb {

// thread-local error text surfaced through ns_last_error()
void set_error(const std::string& msg);
const char* last_error();

// No exception crosses the C ABI: every exported entry point that can allocate runs its body through this barrier
// (std::bad_alloc -> `nomem`, anything else -> `other`; the text goes to ns_last_error()).
template <class F>
int abi_guard(const char* what, int nomem, int other, F&& body) noexcept {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        try { set_error(std::string(what) + ": out of host memory"); } catch (...) {}
        return nomem;
    } catch (const std::exception& ex) {
        try { set_error(std::string(what) + ": " + ex.what()); } catch (...) {}
        return other;
    } catch (...) {
        try { set_error(std::string(what) + ": unknown exception"); } catch (...) {}
        return other;
    }
}

// ---- raw little-endian IO (format of include/indexio.hpp:8-29 in the reference:
// u32/u64/f32 as host bytes, strings as u32 length + bytes) ----
struct Reader {
    const uint8_t* p = nullptr;
    const uint8_t* end = nullptr;
    bool ok = true;
    Reader(const uint8_t* b, size_t n) : p(b), end(b + n) {}
    bool need(size_t n) {
        if ((size_t)(end - p) < n) { ok = false; return false; }
        return true;
    }
    uint32_t u32() { uint32_t v = 0; if (need(4)) { std::memcpy(&v, p, 4); p += 4; } return v; }
    uint64_t u64() { uint64_t v = 0; if (need(8)) { std::memcpy(&v, p, 8); p += 8; } return v; }
    float f32() { float v = 0; if (need(4)) { std::memcpy(&v, p, 4); p += 4; } return v; }
    std::string str() {
        uint32_t n = u32();
        std::string s;
        if (ok && need(n)) { s.assign((const char*)p, n); p += n; }
        return s;
    }
    void skip_str() { uint32_t n = u32(); if (ok && need(n)) p += n; }
};

struct Writer {
    std::vector<uint8_t> buf;
    void u32(uint32_t v) { put(&v, 4); }
    void u64(uint64_t v) { put(&v, 8); }
    void f32(float v) { put(&v, 4); }
    void str(const std::string& s) { u32((uint32_t)s.size()); put(s.data(), s.size()); }
    void put(const void* d, size_t n) {
        const uint8_t* b = (const uint8_t*)d;
        buf.insert(buf.end(), b, b + n);
    }
};

bool read_file(const std::string& path, std::vector<uint8_t>& out);
bool write_file(const std::string& path, const void* data, size_t n);
bool file_exists(const std::string& path);
bool is_dir(const std::string& path);
bool make_dirs(const std::string& path);

// ---- counter-based RNG: every draw is a pure function of (seed, a, b) so that
// segments / docs / tokens can be generated independently and in parallel ----
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static inline uint64_t hash3(uint64_t seed, uint64_t a, uint64_t b) {
    return splitmix64(splitmix64(splitmix64(seed) ^ a) ^ b);
}
static inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }

}  // namespace nsb
