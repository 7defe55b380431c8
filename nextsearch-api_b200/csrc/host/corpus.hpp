// Deterministic synthetic CORD-19-shaped corpus + a streaming writer for the reference's
// barrelized segment format.  This replaces include/segment_writer.hpp:23-169 of the reference
// for corpora too large for its in-memory maps; tests check it byte-for-byte against the
// reference's own SegmentWriter on small corpora (tests/test_oracle_golden.py: sha256 of all 133 files per segment).
//
// Corpus definition (recorded in BASELINE.md / DESIGN.md):
//   term of rank r (1..V) is the string "t<r>";  p(r) ∝ 1/(r+q)^s
//   doc g has length L = len_lo + hash3(seed^0xD0C5EED, g, 0) % (len_hi-len_lo)
//   token i of doc g has rank = inverse-CDF( u01(hash3(seed, g, i+1)) )
//   a doc's term_freqs are listed in ascending rank;  tf = multiplicity;  doc_len = L
//   cord_uid = "uid<g>", title = "", json_relpath = ""
//   termIds are interned in first-seen order, as SegmentWriter::intern_term does
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace nsb {

struct CorpusSpec {
    uint64_t seed = 20260101ULL;
    uint32_t vocab = 400000;
    double zipf_s = 1.0;
    double zipf_q = 25.0;
    uint32_t len_lo = 100;
    uint32_t len_hi = 250;
};

class ZipfSampler {
   public:
    explicit ZipfSampler(const CorpusSpec& spec);
    // rank in 1..V for u in [0,1)
    uint32_t rank(double u) const;

   private:
    std::vector<double> cdf_;
    std::vector<uint32_t> guide_;
    static constexpr uint32_t kGuide = 1u << 16;
};

struct GenSegment {
    uint64_t doc_base = 0;
    uint32_t N = 0;
    float avgdl = 0.0f;
    std::vector<uint32_t> doc_len;
    uint32_t T = 0;
    std::vector<uint32_t> term_rank;   // [T] termId -> zipf rank
    std::vector<uint64_t> term_off;    // [T+1] posting offsets, termId order
    std::vector<uint64_t> postings;    // {u32 doc, u32 tf} interleaved
    // forward index in doc order: (rank, tf) ascending rank
    std::vector<uint64_t> fwd_off;     // [N+1]
    std::vector<uint64_t> fwd;         // rank | tf<<32
};

void generate_segment(const CorpusSpec& spec, uint64_t doc_base, uint32_t ndocs, int nthreads, GenSegment& out);

// Writes stats.bin, docs.bin, barrels.bin, lexicon_bNNN.bin, inverted_bNNN.bin (+ forward.bin,
// terms.bin when write_forward).
bool write_segment_files(const GenSegment& g, const std::string& segdir, bool write_forward);

// Flat dump consumed by oracle/ref_driver.cpp: u32 ndocs; per doc {str uid, u32 doc_len, u32 n,
// n x {str term, u32 tf}}.
bool write_corpus_dump(const GenSegment& g, const std::string& path);

std::vector<std::string> make_queries(const CorpusSpec& spec, uint64_t query_seed, uint32_t nq, uint32_t min_terms,
                                      uint32_t max_terms, uint32_t head_ranks);

}  // namespace nsb
