// Reader for NextSearch's on-disk segment format (the bytes written by the reference's
// include/segment_writer.hpp:65-168 / src/lexicon.cpp and read by src/api_segment.cpp:45-136).
// Unlike the reference, postings are read fully into memory (they go to HBM), and the
// lexicon rows are kept in an array so that a row index can cross the C ABI.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace nsb {

struct LexRow {
    uint32_t termId = 0;   // LexEntry.termId   (include/api_types.hpp:24)
    uint32_t df = 0;       // LexEntry.df       (used for IDF and the df==0 skip)
    uint32_t count = 0;    // LexEntry.count    (number of postings streamed)
    uint32_t barrel = 0;   // LexEntry.barrelId
    uint64_t begin = 0;    // first posting in the concatenated posting array (= barrel_base + offset/8)
    float idf = 0.0f;      // bm25_idf(N, df) — src/api_engine.cpp:45-47, host logf
};

// 64-bit FNV-1a of a term's bytes: computed once per query token, reused for every segment's table.
inline uint64_t term_hash(const char* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) h = (h ^ (unsigned char)p[i]) * 1099511628211ull;
    return h;
}

// Flat open-addressing view of HostSegment::lex for the query front end (one cache line per probe,
// hash computed by the caller).  Keys point into the map's own nodes, which never move.
struct TermTable {
    struct Slot {
        uint64_t h = 0;
        const std::string* key = nullptr;  // nullptr = empty
        uint32_t row = 0;
        uint32_t pad = 0;
    };
    std::vector<Slot> slots;
    uint64_t mask = 0;

    void build(const std::unordered_map<std::string, uint32_t>& lex) {
        size_t cap = 16;
        while (cap < lex.size() * 2 + 1) cap <<= 1;
        slots.assign(cap, Slot{});
        mask = cap - 1;
        for (auto& kv : lex) {
            const uint64_t h = term_hash(kv.first.data(), kv.first.size());
            size_t i = (size_t)(h & mask);
            while (slots[i].key) i = (i + 1) & mask;
            slots[i].h = h;
            slots[i].key = &kv.first;
            slots[i].row = kv.second;
        }
    }
    // row of the term, or -1
    int64_t find(const char* p, size_t n, uint64_t h) const {
        if (slots.empty()) return -1;
        for (size_t i = (size_t)(h & mask);; i = (i + 1) & mask) {
            const Slot& s = slots[i];
            if (!s.key) return -1;
            if (s.h == h && s.key->size() == n && std::memcmp(s.key->data(), p, n) == 0) return (int64_t)s.row;
        }
    }
};

struct HostSegment {
    std::string dir;
    uint32_t N = 0;        // stats.bin
    float avgdl = 0.0f;    // stats.bin, used verbatim
    std::vector<uint32_t> doc_len;   // docs.bin
    std::vector<char> uid_chars;     // concatenated cord_uid bytes
    std::vector<uint64_t> uid_off;   // [n+1]
    std::vector<LexRow> rows;
    std::unordered_map<std::string, uint32_t> lex;  // term -> row (first occurrence wins, like emplace)
    TermTable table;                                // the same mapping, flat (built by load_segment)
    std::vector<uint64_t> postings;  // interleaved {u32 docId, u32 tf} = file bytes
    bool use_barrels = false;
    uint32_t barrel_count = 0, terms_per_barrel = 0;

    std::string cord_uid(uint32_t doc) const {
        if ((size_t)doc + 1 >= uid_off.size()) return std::string();
        return std::string(uid_chars.data() + uid_off[doc], uid_chars.data() + uid_off[doc + 1]);
    }
    void drop_postings() { std::vector<uint64_t>().swap(postings); }
};

// bm25_idf of src/api_engine.cpp:45-47: u32 subtraction first, then float ops, then logf.
float bm25_idf(uint32_t N, uint32_t df);

// src/api_segment.cpp:105-136.  Returns false (error text set) if a file is missing or truncated.
// nthreads > 1 reads barrels in parallel.
bool load_segment(const std::string& segdir, HostSegment& s, int nthreads = 1);

// src/api_segment.cpp:14-42
std::vector<std::string> load_manifest(const std::string& manifest_path);
bool save_manifest(const std::string& manifest_path, const std::vector<std::string>& segs);
std::string seg_name(uint32_t id);
// manifest, else sorted scan of <index_dir>/segments/seg_*  (src/api_engine.cpp:57-70)
std::vector<std::string> discover_segments(const std::string& index_dir);

std::string barrel_suffix(uint32_t b);  // "%03u" — include/barrels.hpp:50-54

}  // namespace nsb
