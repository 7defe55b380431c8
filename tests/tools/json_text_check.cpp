// Test tool: compares nsb::jsontext (the product's JSON text emitter) with nlohmann::json::dump()
// (the library the reference serialises with) on f32 scores widened to double and on strings.
// Built by tests/test_json_text.py against the nlohmann copy the image carries; prints "bad=0" on success.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

#include "nlohmann/json.hpp"
#include "json_text.hpp"

int main(int argc, char** argv) {
    const long n = argc > 1 ? std::atol(argv[1]) : 2000000;
    std::mt19937_64 rng(20260101);
    long bad = 0, checked = 0;
    auto check_double = [&](double v) {
        nlohmann::json j = v;
        std::string a = j.dump(), b;
        nsb::jsontext::append_double(b, v);
        checked++;
        if (a != b) {
            if (bad < 20) std::printf("MISMATCH %a nlohmann=%s ours=%s\n", v, a.c_str(), b.c_str());
            bad++;
        }
    };
    for (long i = 0; i < n; i++) {
        uint32_t bits = (uint32_t)rng();
        if (i % 4 != 0) bits = (bits & 0x007FFFFFu) | ((96u + (uint32_t)(rng() % 48)) << 23);  // BM25-like magnitudes
        float f;
        std::memcpy(&f, &bits, 4);
        check_double((double)f);  // r["score"] = h.s widens the float (src/api_engine.cpp:511)
    }
    for (long i = 0; i < n / 4; i++) {  // arbitrary doubles too
        uint64_t b = rng();
        double d;
        std::memcpy(&d, &b, 8);
        check_double(d);
    }
    const double special[] = {0.0, -0.0, 1.0, 1e-5, 9.999e-5, 1e-4, 1e15, 1e16, 123456789012345.0, 1234567890123456.0,
                              0.1, 0.5, 1e21, 1e22, 5e-324, 1.7976931348623157e308, 2.2250738585072014e-308, 100.0, 1e2};
    for (double v : special) { check_double(v); check_double(-v); }
    // strings: every byte value in context, escapes, valid multi-byte sequences
    for (int c = 1; c < 0x80; c++) {
        std::string s = std::string("a") + (char)c + "z";
        nlohmann::json j = s;
        std::string a = j.dump(), b;
        bool ok = nsb::jsontext::append_string(b, s);
        checked++;
        if (!ok || a != b) { std::printf("MISMATCH string byte %d\n", c); bad++; }
    }
    const char* utf8_ok[] = {"\xC3\xA9", "\xE2\x82\xAC", "\xF0\x9F\x98\x80", "\xED\x9F\xBF", "\xEE\x80\x80", "\xF4\x8F\xBF\xBF"};
    for (const char* u : utf8_ok) {
        nlohmann::json j = std::string(u);
        std::string a = j.dump(), b;
        if (!nsb::jsontext::append_string(b, std::string(u)) || a != b) { std::printf("MISMATCH utf8 ok\n"); bad++; }
        checked++;
    }
    const char* utf8_bad[] = {"\x80", "\xC0\xAF", "\xC3", "\xE0\x80\x80", "\xED\xA0\x80", "\xF4\x90\x80\x80", "\xF8\x88\x80\x80\x80", "a\xFFz"};
    for (const char* u : utf8_bad) {
        bool threw = false;
        try { nlohmann::json j = std::string(u); (void)j.dump(); } catch (const nlohmann::json::type_error&) { threw = true; }
        std::string b;
        const bool ok = nsb::jsontext::append_string(b, std::string(u));
        checked++;
        if (threw == ok) { std::printf("MISMATCH utf8 bad\n"); bad++; }
    }
    std::printf("checked=%ld bad=%ld\n", checked, bad);
    return bad ? 1 : 0;
}
