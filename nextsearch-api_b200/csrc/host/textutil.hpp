// Query text front end: same token stream as the reference's include/textutil.hpp:13-37 plus the
// filter of src/api_engine.cpp:391-397 (drop len<2 and stopwords, keep order and duplicates).
#pragma once
#include <cstring>
#include <string>
#include <vector>

namespace nsb {

// ASCII [A-Za-z0-9]+ runs, lower-cased.  The reference calls std::isalnum/std::tolower in the
// "C" locale, where only ASCII letters and digits qualify; bytes >= 0x80 separate tokens.
inline void tokenize(const char* text, std::vector<std::string>& out) {
    out.clear();
    std::string cur;
    for (const unsigned char* p = (const unsigned char*)text; *p; ++p) {
        unsigned char c = *p;
        bool digit = c >= '0' && c <= '9';
        bool lower = c >= 'a' && c <= 'z';
        bool upper = c >= 'A' && c <= 'Z';
        if (digit || lower || upper) {
            cur.push_back(upper ? (char)(c - 'A' + 'a') : (char)c);
        } else if (!cur.empty()) {
            out.push_back(cur);
            cur.clear();
        }
    }
    if (!cur.empty()) out.push_back(cur);
}

inline bool is_stopword(const std::string& t) {
    // the 24 words of include/textutil.hpp:32-35; all have length 1..4
    static const char* const sw[] = {"the", "a",  "an",  "and",  "or", "of",   "to", "in",   "for",  "on",   "with", "by",
                                     "as",  "is", "are", "was",  "were", "be", "been", "it", "this", "that", "from", "at"};
    if (t.size() > 4) return false;
    for (const char* w : sw)
        if (t == w) return true;
    return false;
}

// Allocation-free form for the batched front end: the kept tokens (lower-cased) are appended to `buf`
// back to back; `spans` receives (offset, length) of each.  Same token stream as query_terms.
struct TokSpan {
    uint32_t off, len;
};
inline void query_term_spans(const char* query, std::string& buf, std::vector<TokSpan>& spans) {
    buf.clear();
    spans.clear();
    size_t start = 0;
    auto close = [&]() {
        const size_t n = buf.size() - start;
        bool keep = n >= 2;
        if (keep && n <= 4) {
            static const char* const sw[] = {"the", "an",  "and",  "or", "of",   "to", "in",   "for",  "on",   "with", "by", "as",
                                             "is",  "are", "was",  "were", "be", "been", "it", "this", "that", "from", "at"};
            for (const char* w : sw)
                if (std::strlen(w) == n && std::memcmp(w, buf.data() + start, n) == 0) {
                    keep = false;
                    break;
                }
        }
        if (keep) {
            spans.push_back(TokSpan{(uint32_t)start, (uint32_t)n});
            start = buf.size();
        } else {
            buf.resize(start);
        }
    };
    for (const unsigned char* p = (const unsigned char*)query; *p; ++p) {
        const unsigned char c = *p;
        const bool digit = c >= '0' && c <= '9', lower = c >= 'a' && c <= 'z', upper = c >= 'A' && c <= 'Z';
        if (digit || lower || upper) buf.push_back(upper ? (char)(c - 'A' + 'a') : (char)c);
        else if (buf.size() > start) close();
    }
    if (buf.size() > start) close();
}

inline void query_terms(const char* query, std::vector<std::string>& out) {
    static thread_local std::vector<std::string> toks;  // reused: no allocation per query in steady state
    tokenize(query, toks);
    out.clear();
    for (auto& t : toks) {
        if (t.size() < 2) continue;
        if (is_stopword(t)) continue;
        out.push_back(std::move(t));
    }
}

}  // namespace nsb
