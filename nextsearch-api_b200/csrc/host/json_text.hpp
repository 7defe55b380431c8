// JSON text emission that is byte-compatible with what the reference returns over HTTP:
// Engine::search builds an nlohmann::json object and api_server writes j.dump()
// (src/api_engine.cpp:400-404,508-536, src/api_server.cpp:134-136).  nlohmann/json is an un-vendored
// third-party dependency of the reference (third_party/nlohmann/json.hpp, .gitignore'd; 3.11.x), so its
// published serialisation rules are restated here:
//   * object keys in lexicographic order (std::map), no whitespace;
//   * strings: \" \\ \b \f \n \r \t, other bytes < 0x20 as \u00xx (lower-case hex), everything else
//     verbatim; input must be valid UTF-8 — dump() throws type_error.316 otherwise (the HTTP handler
//     turns that into a 500), reported here as `false`;
//   * floating point numbers: Grisu2 (Loitsch, "Printing floating-point numbers quickly and accurately
//     with integers", PLDI 2010) with alpha = -60, gamma = -32 and a table of cached powers 10^k every
//     8th k — NOT always the shortest/closest digit string, so std::to_chars cannot stand in for it —
//     then fixed notation for decimal exponents in (-4, 15], scientific otherwise, ".0" appended to
//     integral values; exponents carry a sign and at least two digits.
// tests/test_json_text.py compares this formatter with the image's nlohmann 3.11.3 on tens of millions
// of f32 scores widened to double (r["score"] = h.s, src/api_engine.cpp:511).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

namespace nsb {
namespace jsontext {

struct Fp {  // f * 2^e
    uint64_t f;
    int e;
};

inline Fp fp_mul(Fp x, Fp y) {
    // upper 64 bits of the 128-bit product, rounded half up
    const unsigned __int128 p = (unsigned __int128)x.f * y.f + ((unsigned __int128)1 << 63);
    return Fp{(uint64_t)(p >> 64), x.e + y.e + 64};
}

inline Fp fp_normalize(Fp x) {
    while ((x.f >> 63) == 0) {
        x.f <<= 1;
        x.e--;
    }
    return x;
}

struct Pow10 {
    uint64_t f;
    int e;
    int k;
};

inline const Pow10& cached_pow10(int binary_exponent) {
    static const Pow10 table[] = {
#include "pow10_table.inc"
    };
    // smallest k with alpha <= e_c + e + 64: k = ceil((alpha - e - 1) * log10(2))
    const int f = -60 - binary_exponent - 1;
    const int k = (f * 78913) / (1 << 18) + (f > 0 ? 1 : 0);
    const int index = (300 + k + 7) / 8;
    return table[index];
}

// Grisu2 for a finite, strictly positive double: digits into buf (no NUL), value = digits * 10^dec_exp
inline void grisu2(double value, char* buf, int& len, int& dec_exp) {
    uint64_t bits;
    std::memcpy(&bits, &value, 8);
    const uint64_t frac = bits & ((1ull << 52) - 1);
    const int biased = (int)(bits >> 52) & 0x7FF;
    const bool denormal = biased == 0;
    const Fp v = denormal ? Fp{frac, 1 - 1075} : Fp{frac | (1ull << 52), biased - 1075};
    // neighbours' midpoints; the lower one is closer when v is a power of two (except the smallest normal)
    const bool lower_closer = frac == 0 && biased > 1;
    const Fp plus = fp_normalize(Fp{2 * v.f + 1, v.e - 1});
    Fp minus = lower_closer ? Fp{4 * v.f - 1, v.e - 2} : Fp{2 * v.f - 1, v.e - 1};
    minus.f <<= (minus.e - plus.e);
    minus.e = plus.e;
    const Fp w = fp_normalize(v);

    const Pow10& c = cached_pow10(plus.e);
    const Fp cp{c.f, c.e};
    const Fp W = fp_mul(w, cp);
    Fp lo = fp_mul(minus, cp), hi = fp_mul(plus, cp);
    lo.f += 1;  // shrink the interval by one unit on each side: every number inside rounds to v
    hi.f -= 1;
    dec_exp = -c.k;

    // digit generation for hi, stopping as soon as the remainder fits into the interval
    uint64_t delta = hi.f - lo.f, dist = hi.f - W.f;
    const int sh = -hi.e;  // 32..60
    const uint64_t one = 1ull << sh;
    uint32_t p1 = (uint32_t)(hi.f >> sh);
    uint64_t p2 = hi.f & (one - 1);
    auto round_weed = [&](uint64_t rest, uint64_t unit) {
        // move the last digit down while that brings the number closer to w and keeps it in the interval
        while (rest < dist && delta - rest >= unit && (rest + unit < dist || dist - rest > rest + unit - dist)) {
            buf[len - 1]--;
            rest += unit;
        }
    };
    uint32_t pow10 = 1;
    int n = 1;
    while (n < 10 && p1 >= pow10 * 10u) {
        pow10 *= 10u;
        n++;
    }
    len = 0;
    while (n > 0) {
        const uint32_t d = p1 / pow10;
        p1 -= d * pow10;
        buf[len++] = (char)('0' + d);
        n--;
        const uint64_t rest = ((uint64_t)p1 << sh) + p2;
        if (rest <= delta) {
            dec_exp += n;
            round_weed(rest, (uint64_t)pow10 << sh);
            return;
        }
        pow10 /= 10u;
    }
    int m = 0;
    for (;;) {
        p2 *= 10;
        const uint64_t d = p2 >> sh;
        p2 &= one - 1;
        buf[len++] = (char)('0' + d);
        m++;
        delta *= 10;
        dist *= 10;
        if (p2 <= delta) break;
    }
    dec_exp -= m;
    round_weed(p2, one);
}

// Appends the JSON text of a double.  NaN / infinities serialise as null.
inline void append_double(std::string& out, double v) {
    if (!(v == v) || v - v != 0.0) {
        out += "null";
        return;
    }
    uint64_t bits;
    std::memcpy(&bits, &v, 8);
    if (bits >> 63) {
        out.push_back('-');
        v = -v;
    }
    if (v == 0.0) {
        out += "0.0";
        return;
    }
    char d[32];
    int len = 0, dec = 0;
    grisu2(v, d, len, dec);
    const int n = len + dec;  // position of the decimal point relative to the first digit
    if (len <= n && n <= 15) {
        out.append(d, (size_t)len);
        out.append((size_t)(n - len), '0');
        out += ".0";
    } else if (0 < n && n <= 15) {
        out.append(d, (size_t)n);
        out.push_back('.');
        out.append(d + n, (size_t)(len - n));
    } else if (-4 < n && n <= 0) {
        out += "0.";
        out.append((size_t)(-n), '0');
        out.append(d, (size_t)len);
    } else {
        out.push_back(d[0]);
        if (len > 1) {
            out.push_back('.');
            out.append(d + 1, (size_t)(len - 1));
        }
        out.push_back('e');
        int e = n - 1;
        out.push_back(e < 0 ? '-' : '+');
        if (e < 0) e = -e;
        if (e < 10) out.push_back('0');
        out += std::to_string(e);
    }
}

// Well-formed UTF-8 per Unicode Table 3-7 (no overlong forms, no surrogates, <= U+10FFFF): what
// nlohmann's serializer accepts under its default (strict) error handler.
inline bool valid_utf8(const char* s, size_t n) {
    const unsigned char* p = (const unsigned char*)s;
    size_t i = 0;
    while (i < n) {
        const unsigned char c = p[i];
        if (c < 0x80) {
            i++;
            continue;
        }
        size_t need;
        unsigned char lo = 0x80, hi = 0xBF;
        if (c >= 0xC2 && c <= 0xDF) need = 1;
        else if (c == 0xE0) { need = 2; lo = 0xA0; }
        else if (c >= 0xE1 && c <= 0xEC) need = 2;
        else if (c == 0xED) { need = 2; hi = 0x9F; }
        else if (c >= 0xEE && c <= 0xEF) need = 2;
        else if (c == 0xF0) { need = 3; lo = 0x90; }
        else if (c >= 0xF1 && c <= 0xF3) need = 3;
        else if (c == 0xF4) { need = 3; hi = 0x8F; }
        else return false;
        if (i + need >= n) return false;  // truncated sequence
        if (p[i + 1] < lo || p[i + 1] > hi) return false;
        for (size_t j = 2; j <= need; j++)
            if (p[i + j] < 0x80 || p[i + j] > 0xBF) return false;
        i += need + 1;
    }
    return true;
}

// Appends "…" with nlohmann's escapes; returns false (out untouched past the opening state) on invalid UTF-8.
inline bool append_string(std::string& out, const char* s, size_t n) {
    if (!valid_utf8(s, n)) return false;
    out.push_back('"');
    for (size_t i = 0; i < n; i++) {
        const unsigned char c = (unsigned char)s[i];
        switch (c) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\b': out += "\\b"; break;
            case '\f': out += "\\f"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            default:
                if (c < 0x20) {
                    static const char hex[] = "0123456789abcdef";
                    out += "\\u00";
                    out.push_back(hex[c >> 4]);
                    out.push_back(hex[c & 15]);
                } else {
                    out.push_back((char)c);
                }
        }
    }
    out.push_back('"');
    return true;
}
inline bool append_string(std::string& out, const std::string& s) { return append_string(out, s.data(), s.size()); }

}  // namespace jsontext
}  // namespace nsb
