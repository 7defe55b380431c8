// Reader for NextSearch's on-disk segment format (the bytes written by the reference's
// include/segment_writer.hpp:65-168 / src/lexicon.cpp and read by src/api_segment.cpp:45-136).
// Unlike the reference, postings are read fully into memory (they go to HBM), and the
// lexicon rows are kept in an array so that a row index can cross the C ABI.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
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

// Where load_segment puts the posting bytes.  The default keeps them in HostSegment::postings; the engine
// supplies a sink backed by pinned memory whose filled() starts the host->device copy of every finished
// barrel while the other barrels are still being read (ns_upload_*).
struct PostingSink {
    virtual ~PostingSink() = default;
    // called once: room for P postings (8 bytes each) or nullptr on failure
    virtual uint8_t* begin(uint64_t P) = 0;
    // postings [first, first + count) are in place; may be called from several reader threads
    virtual void filled(uint64_t first, uint64_t count) = 0;
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
// nthreads > 1 reads barrels in parallel.  sink != nullptr receives the postings instead of s.postings.
bool load_segment(const std::string& segdir, HostSegment& s, int nthreads = 1, PostingSink* sink = nullptr);

// One dictionary over the lexicons of all segments an engine owns: a query token is hashed and probed ONCE,
// and its (row, idf) in every owned segment sits in one contiguous run — the front end's cost per token no
// longer grows with the number of segments (the reference probes one unordered_map per segment,
// src/api_engine.cpp:454).  Open addressing, 64-bit FNV-1a computed by the caller, keys compared in full.
struct TermDict {
    static constexpr uint32_t kAbsent = 0xFFFFFFFFu;
    struct Entry {
        uint32_t row;  // kAbsent: the term is not in that segment, or its df is 0 (src/api_engine.cpp:455,458)
        float idf;     // bm25_idf(N, df) of that segment
        uint32_t count;  // LexEntry.count: postings of the row
    };
    struct Slot {  // 32 bytes: a probe touches one cache line, keys of up to 12 bytes are compared in place
        uint64_t h = 0;
        uint32_t gid = kAbsent;  // kAbsent = empty slot
        uint32_t key_off = 0;    // into keys[] (longer keys)
        uint32_t key_len = 0;
        char inl[12] = {0};
    };
    std::vector<uint32_t> owned;   // global segment index of column j, ascending
    std::vector<Slot> slots;
    uint64_t mask = 0;
    std::vector<char> keys;
    std::vector<Entry> table;      // [gid * owned.size() + j]
    uint32_t nterms = 0;

    // segs[i] may be null (not owned)
    void build(const std::vector<std::unique_ptr<HostSegment>>& segs);
    static bool key_equal(const Slot& s, const char* keys, const char* p, size_t n) {
        if (s.key_len != n) return false;
        return std::memcmp(n <= sizeof(s.inl) ? s.inl : keys + s.key_off, p, n) == 0;
    }
    void prefetch(uint64_t h) const {
        if (!slots.empty()) __builtin_prefetch(&slots[(size_t)(h & mask)]);
    }
    // global term id or -1
    int64_t find(const char* p, size_t n, uint64_t h) const {
        if (slots.empty()) return -1;
        for (size_t i = (size_t)(h & mask);; i = (i + 1) & mask) {
            const Slot& s = slots[i];
            if (s.gid == kAbsent) return -1;
            if (s.h == h && key_equal(s, keys.data(), p, n)) return (int64_t)s.gid;
        }
    }
    const Entry* row_of(uint32_t gid) const { return table.data() + (size_t)gid * owned.size(); }
};

// src/api_segment.cpp:14-42
std::vector<std::string> load_manifest(const std::string& manifest_path);
bool save_manifest(const std::string& manifest_path, const std::vector<std::string>& segs);
std::string seg_name(uint32_t id);
// manifest, else sorted scan of <index_dir>/segments/seg_*  (src/api_engine.cpp:57-70)
std::vector<std::string> discover_segments(const std::string& index_dir);

std::string barrel_suffix(uint32_t b);  // "%03u" — include/barrels.hpp:50-54

}  // namespace nsb
