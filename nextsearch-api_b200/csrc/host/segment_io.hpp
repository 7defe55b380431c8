// Reader for NextSearch's on-disk segment format (the bytes written by the reference's
// include/segment_writer.hpp:65-168 / src/lexicon.cpp and read by src/api_segment.cpp:45-136).
// Unlike the reference, postings are read fully into memory (they go to HBM), and the
// lexicon rows are kept in an array so that a row index can cross the C ABI.
#pragma once
#include <cstdint>
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
