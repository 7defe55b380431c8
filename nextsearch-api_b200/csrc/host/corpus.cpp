#include "corpus.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <thread>

#include "common.hpp"
#include "segment_io.hpp"

namespace nsb {

ZipfSampler::ZipfSampler(const CorpusSpec& spec) {
    const uint32_t V = std::max<uint32_t>(1, spec.vocab);
    cdf_.resize(V);
    double total = 0.0;
    for (uint32_t r = 1; r <= V; r++) {
        double x = (double)r + spec.zipf_q;
        double w = (spec.zipf_s == 1.0) ? 1.0 / x : std::pow(x, -spec.zipf_s);
        total += w;
        cdf_[r - 1] = total;
    }
    for (uint32_t i = 0; i < V; i++) cdf_[i] /= total;
    cdf_[V - 1] = 1.0;
    guide_.resize((size_t)kGuide + 1);
    for (uint32_t b = 0; b <= kGuide; b++) {
        double x = (double)b / (double)kGuide;
        guide_[b] = (uint32_t)(std::upper_bound(cdf_.begin(), cdf_.end(), x) - cdf_.begin());
    }
}

uint32_t ZipfSampler::rank(double u) const {
    uint32_t b = (uint32_t)(u * (double)kGuide);
    if (b >= kGuide) b = kGuide - 1;
    auto lo = cdf_.begin() + guide_[b];
    auto hi = cdf_.begin() + guide_[b + 1];
    uint32_t idx = (uint32_t)(std::upper_bound(lo, hi, u) - cdf_.begin());
    uint32_t V = (uint32_t)cdf_.size();
    return (idx >= V ? V - 1 : idx) + 1;
}

static inline uint32_t doc_length(const CorpusSpec& spec, uint64_t g) {
    uint32_t span = spec.len_hi > spec.len_lo ? spec.len_hi - spec.len_lo : 1;
    return spec.len_lo + (uint32_t)(hash3(spec.seed ^ 0xD0C5EEDULL, g, 0) % span);
}

void generate_segment(const CorpusSpec& spec, uint64_t doc_base, uint32_t ndocs, int nthreads, GenSegment& out) {
    out = GenSegment{};
    out.doc_base = doc_base;
    out.N = ndocs;
    const uint32_t V = std::max<uint32_t>(1, spec.vocab);
    ZipfSampler zipf(spec);
    int nt = std::max(1, nthreads);
    if ((uint32_t)nt > std::max<uint32_t>(1, ndocs)) nt = (int)std::max<uint32_t>(1, ndocs);

    out.doc_len.resize(ndocs);
    std::vector<uint32_t> fwd_cnt(ndocs, 0);
    std::vector<std::vector<uint64_t>> tfwd(nt);
    std::vector<std::vector<uint32_t>> tcnt(nt), tfirst(nt);
    auto lo_of = [&](int t) { return (uint32_t)((uint64_t)ndocs * t / nt); };

    auto pass1 = [&](int t) {
        uint32_t d0 = lo_of(t), d1 = lo_of(t + 1);
        auto& cnt = tcnt[t];
        auto& first = tfirst[t];
        auto& fw = tfwd[t];
        cnt.assign(V, 0);
        first.assign(V, UINT32_MAX);
        fw.reserve((size_t)(d1 - d0) * 140);
        std::vector<uint32_t> toks;
        for (uint32_t d = d0; d < d1; d++) {
            uint64_t g = doc_base + d;
            uint32_t L = doc_length(spec, g);
            out.doc_len[d] = L;
            toks.resize(L);
            for (uint32_t i = 0; i < L; i++) toks[i] = zipf.rank(u01(hash3(spec.seed, g, (uint64_t)i + 1)));
            std::sort(toks.begin(), toks.end());
            uint32_t n = 0;
            for (uint32_t i = 0; i < L;) {
                uint32_t j = i;
                while (j < L && toks[j] == toks[i]) j++;
                uint32_t r = toks[i];
                fw.push_back((uint64_t)r | ((uint64_t)(j - i) << 32));
                if (cnt[r - 1]++ == 0) first[r - 1] = d;
                n++;
                i = j;
            }
            fwd_cnt[d] = n;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(pass1, t);
        pass1(0);
        for (auto& x : th) x.join();
    }

    uint64_t total_len = 0;
    for (uint32_t d = 0; d < ndocs; d++) total_len += out.doc_len[d];
    // include/segment_writer.hpp:68
    out.avgdl = ndocs == 0 ? 0.0f : (float)total_len / (float)ndocs;

    out.fwd_off.resize((size_t)ndocs + 1);
    out.fwd_off[0] = 0;
    for (uint32_t d = 0; d < ndocs; d++) out.fwd_off[d + 1] = out.fwd_off[d] + fwd_cnt[d];
    out.fwd.resize(out.fwd_off[ndocs]);
    for (int t = 0; t < nt; t++) {
        std::copy(tfwd[t].begin(), tfwd[t].end(), out.fwd.begin() + out.fwd_off[lo_of(t)]);
        std::vector<uint64_t>().swap(tfwd[t]);
    }

    // termIds in first-seen order: (first doc, rank) ascending
    struct Seen { uint32_t first, rank, df; };
    std::vector<Seen> seen;
    for (uint32_t r = 0; r < V; r++) {
        uint32_t df = 0, first = UINT32_MAX;
        for (int t = 0; t < nt; t++) {
            if (tcnt[t][r]) {
                if (first == UINT32_MAX) first = tfirst[t][r];
                df += tcnt[t][r];
            }
        }
        if (df) seen.push_back({first, r + 1, df});
    }
    std::sort(seen.begin(), seen.end(), [](const Seen& a, const Seen& b) {
        return a.first != b.first ? a.first < b.first : a.rank < b.rank;
    });
    out.T = (uint32_t)seen.size();
    out.term_rank.resize(out.T);
    out.term_off.resize((size_t)out.T + 1);
    out.term_off[0] = 0;
    std::vector<uint32_t> tid_of_rank(V, UINT32_MAX);
    for (uint32_t i = 0; i < out.T; i++) {
        out.term_rank[i] = seen[i].rank;
        out.term_off[i + 1] = out.term_off[i] + seen[i].df;
        tid_of_rank[seen[i].rank - 1] = i;
    }
    out.postings.resize(out.term_off[out.T]);

    // per-thread write cursors so that every list comes out in ascending docId
    std::vector<std::vector<uint64_t>> cur(nt);
    for (int t = 0; t < nt; t++) cur[t].resize(V);
    for (uint32_t r = 0; r < V; r++) {
        uint32_t tid = tid_of_rank[r];
        if (tid == UINT32_MAX) continue;
        uint64_t c = out.term_off[tid];
        for (int t = 0; t < nt; t++) {
            cur[t][r] = c;
            c += tcnt[t][r];
        }
    }
    auto pass2 = [&](int t) {
        uint32_t d0 = lo_of(t), d1 = lo_of(t + 1);
        auto& c = cur[t];
        for (uint32_t d = d0; d < d1; d++) {
            for (uint64_t i = out.fwd_off[d]; i < out.fwd_off[d + 1]; i++) {
                uint32_t r = (uint32_t)out.fwd[i];
                uint32_t tf = (uint32_t)(out.fwd[i] >> 32);
                out.postings[c[r - 1]++] = (uint64_t)d | ((uint64_t)tf << 32);
            }
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < nt; t++) th.emplace_back(pass2, t);
        pass2(0);
        for (auto& x : th) x.join();
    }
}

static std::string term_string(uint32_t rank) { return "t" + std::to_string(rank); }

bool write_segment_files(const GenSegment& g, const std::string& segdir, bool write_forward) {
    if (!make_dirs(segdir)) { set_error("cannot create " + segdir); return false; }
    {
        Writer w;  // stats.bin — include/segment_writer.hpp:71-75
        w.u32(g.N);
        w.f32(g.avgdl);
        if (!write_file(segdir + "/stats.bin", w.buf.data(), w.buf.size())) { set_error("write stats.bin"); return false; }
    }
    {
        Writer w;  // docs.bin — :78-87
        w.buf.reserve((size_t)g.N * 28 + 4);
        w.u32(g.N);
        for (uint32_t d = 0; d < g.N; d++) {
            w.str("uid" + std::to_string(g.doc_base + d));
            w.str(std::string());
            w.str(std::string());
            w.u32(g.doc_len[d]);
        }
        if (!write_file(segdir + "/docs.bin", w.buf.data(), w.buf.size())) { set_error("write docs.bin"); return false; }
    }
    std::vector<uint32_t> tid_of_rank;
    if (write_forward) {
        uint32_t maxr = 0;
        for (uint32_t r : g.term_rank) maxr = std::max(maxr, r);
        tid_of_rank.assign((size_t)maxr + 1, 0);
        for (uint32_t i = 0; i < g.T; i++) tid_of_rank[g.term_rank[i]] = i;
        Writer w;  // forward.bin — :91-101 (per doc sorted by termId)
        w.buf.reserve(g.fwd.size() * 8 + (size_t)g.N * 4 + 4);
        w.u32(g.N);
        std::vector<std::pair<uint32_t, uint32_t>> row;
        for (uint32_t d = 0; d < g.N; d++) {
            row.clear();
            for (uint64_t i = g.fwd_off[d]; i < g.fwd_off[d + 1]; i++)
                row.push_back({tid_of_rank[(uint32_t)g.fwd[i]], (uint32_t)(g.fwd[i] >> 32)});
            std::sort(row.begin(), row.end());
            w.u32((uint32_t)row.size());
            for (auto& p : row) { w.u32(p.first); w.u32(p.second); }
        }
        if (!write_file(segdir + "/forward.bin", w.buf.data(), w.buf.size())) { set_error("write forward.bin"); return false; }
        Writer t;  // terms.bin — :104-108
        t.u32(g.T);
        for (uint32_t i = 0; i < g.T; i++) t.str(term_string(g.term_rank[i]));
        if (!write_file(segdir + "/terms.bin", t.buf.data(), t.buf.size())) { set_error("write terms.bin"); return false; }
    }
    // barrels — :114-166, include/barrels.hpp:26-47
    const uint32_t B = 64;
    uint32_t tpb = (g.T + B - 1) / B;
    if (tpb == 0) tpb = 1;
    {
        Writer w;
        w.u32(B);
        w.u32(tpb);
        if (!write_file(segdir + "/barrels.bin", w.buf.data(), w.buf.size())) { set_error("write barrels.bin"); return false; }
    }
    for (uint32_t b = 0; b < B; b++) {
        // barrel b holds termIds [b*tpb, (b+1)*tpb), the last barrel everything beyond
        uint64_t t0 = std::min<uint64_t>((uint64_t)b * tpb, g.T);
        uint64_t t1 = (b == B - 1) ? g.T : std::min<uint64_t>((uint64_t)(b + 1) * tpb, g.T);
        Writer lx;
        lx.u32((uint32_t)(t1 - t0));
        uint64_t base = g.term_off[t0];
        for (uint64_t t = t0; t < t1; t++) {
            uint32_t df = (uint32_t)(g.term_off[t + 1] - g.term_off[t]);
            lx.str(term_string(g.term_rank[t]));
            lx.u32((uint32_t)t);
            lx.u32(df);
            lx.u64((g.term_off[t] - base) * 8);
            lx.u32(df);
        }
        std::string sfx = barrel_suffix(b);
        if (!write_file(segdir + "/lexicon_b" + sfx + ".bin", lx.buf.data(), lx.buf.size())) { set_error("write lexicon barrel"); return false; }
        if (!write_file(segdir + "/inverted_b" + sfx + ".bin", g.postings.data() + base, (g.term_off[t1] - base) * 8)) {
            set_error("write inverted barrel");
            return false;
        }
    }
    return true;
}

bool write_corpus_dump(const GenSegment& g, const std::string& path) {
    Writer w;
    w.u32(g.N);
    for (uint32_t d = 0; d < g.N; d++) {
        w.str("uid" + std::to_string(g.doc_base + d));
        w.u32(g.doc_len[d]);
        w.u32((uint32_t)(g.fwd_off[d + 1] - g.fwd_off[d]));
        for (uint64_t i = g.fwd_off[d]; i < g.fwd_off[d + 1]; i++) {
            w.str(term_string((uint32_t)g.fwd[i]));
            w.u32((uint32_t)(g.fwd[i] >> 32));
        }
    }
    return write_file(path, w.buf.data(), w.buf.size());
}

std::vector<std::string> make_queries(const CorpusSpec& spec, uint64_t query_seed, uint32_t nq, uint32_t min_terms,
                                      uint32_t max_terms, uint32_t head_ranks) {
    ZipfSampler zipf(spec);
    std::vector<std::string> out;
    out.reserve(nq);
    if (min_terms < 1) min_terms = 1;
    if (max_terms < min_terms) max_terms = min_terms;
    for (uint32_t i = 0; i < nq; i++) {
        uint32_t n = min_terms + (uint32_t)(hash3(query_seed, i, 0) % (max_terms - min_terms + 1));
        std::string q;
        for (uint32_t j = 0; j < n; j++) {
            uint64_t h = hash3(query_seed, i, (uint64_t)j + 1);
            uint32_t r = (j == 0 && head_ranks > 0) ? 1 + (uint32_t)(h % head_ranks) : zipf.rank(u01(h));
            if (j) q.push_back(' ');
            q += term_string(r);
        }
        out.push_back(std::move(q));
    }
    return out;
}

}  // namespace nsb
