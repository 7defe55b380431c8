// Host mirror of cord19::Engine for the search path (reference: include/api_engine.hpp:23-91,
// src/api_engine.cpp:50-90 and :369-542), plus the corpus tooling entry points.
// Tokenise / stop-filter / lexicon lookup / IDF stay on the host exactly as in the reference;
// the per-posting work goes through ns_search_batch.
#include <algorithm>
#include <cstdlib>
#include <atomic>
#include <charconv>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/nextsearch_b200.h"
#include "common.hpp"
#include "corpus.hpp"
#include "segment_io.hpp"
#include "textutil.hpp"
#include "workpool.hpp"

using namespace nsb;

struct ns_engine {
    std::string index_dir;
    int device = -1;
    int rank = 0, world = 1;
    ns_index* idx = nullptr;
    mutable std::shared_mutex mu;  // reload takes it exclusively; searches share it
    std::vector<std::string> seg_names;
    // segs[i] is loaded only when this engine owns segment i (i % world == rank)
    std::vector<std::unique_ptr<HostSegment>> segs;
    bool owns(size_t i) const { return (int)(i % (size_t)world) == rank; }
    std::unique_ptr<WorkPool> pool;  // query front-end workers, created on first batch
    std::once_flag pool_once;
};

namespace {

int hw_threads() {
    unsigned n = std::thread::hardware_concurrency();
    return n == 0 ? 1 : (int)n;
}

void json_escape(const std::string& s, std::string& out) {
    for (unsigned char c : s) {
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
                    char buf[8];
                    std::snprintf(buf, sizeof(buf), "\\u%04x", c);
                    out += buf;
                } else {
                    out.push_back((char)c);
                }
        }
    }
}

// shortest round-trip decimal of the f32 score widened to double (r["score"] = h.s,
// src/api_engine.cpp:511), with ".0" appended to integral values like nlohmann::json::dump
void json_double(double v, std::string& out) {
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof(buf), v);
    std::string s(buf, res.ptr);
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
    out += s;
}

// Resolve one query against the owned segments.  Emits terms ordered by (segment asc, query order).
// Returns whether the reference would compute "found" (base_terms and segments non-empty).
bool resolve_one(const ns_engine* e, const char* query, std::vector<ns_qterm>& out) {
    static thread_local std::string buf;
    static thread_local std::vector<TokSpan> spans;
    static thread_local std::vector<uint64_t> hashes;
    query_term_spans(query, buf, spans);
    if (spans.empty() || e->seg_names.empty()) return false;  // src/api_engine.cpp:407
    hashes.resize(spans.size());
    for (size_t i = 0; i < spans.size(); i++) hashes[i] = term_hash(buf.data() + spans[i].off, spans[i].len);  // once per token
    for (size_t si = 0; si < e->segs.size(); si++) {
        const HostSegment* seg = e->segs[si].get();
        if (!seg) continue;
        for (size_t i = 0; i < spans.size(); i++) {  // qweight 1.0f: src/api_engine.cpp:420
            const int64_t row = seg->table.find(buf.data() + spans[i].off, spans[i].len, hashes[i]);
            if (row < 0) continue;  // :454-455
            const LexRow& r = seg->rows[(size_t)row];
            if (r.df == 0) continue;  // :458
            out.push_back(ns_qterm{(uint32_t)si, (uint32_t)row, r.idf, 1.0f});
        }
    }
    return true;
}

}  // namespace

extern "C" int ns_engine_create(const char* index_dir, int device, ns_engine** out) {
    if (!index_dir || !out) { set_error("ns_engine_create: null argument"); return NS_ERR_INVALID; }
    *out = nullptr;
    auto e = std::make_unique<ns_engine>();
    e->index_dir = index_dir;
    e->device = device;
    if (device >= 0) {
        int rc = ns_index_create(device, &e->idx);
        if (rc != NS_OK) return rc;
    }
    *out = e.release();
    return NS_OK;
}

extern "C" void ns_engine_destroy(ns_engine* e) {
    if (!e) return;
    if (e->idx) ns_index_destroy(e->idx);
    delete e;
}

extern "C" int ns_engine_set_shard(ns_engine* e, int rank, int world) {
    if (!e || world < 1 || rank < 0 || rank >= world) { set_error("ns_engine_set_shard: bad rank/world"); return NS_ERR_INVALID; }
    std::unique_lock<std::shared_mutex> lk(e->mu);
    e->rank = rank;
    e->world = world;
    return NS_OK;
}

extern "C" int ns_engine_reload(ns_engine* e) {
    if (!e) { set_error("ns_engine_reload: null"); return NS_ERR_INVALID; }
    std::unique_lock<std::shared_mutex> lk(e->mu);
    std::vector<std::string> names = discover_segments(e->index_dir);
    if (names.empty()) { set_error("no segments under " + e->index_dir); return NS_ERR_IO; }  // :73
    std::vector<std::unique_ptr<HostSegment>> loaded(names.size());
    const int nt = hw_threads();
    for (size_t i = 0; i < names.size(); i++) {
        if ((int)(i % (size_t)e->world) != e->rank) continue;
        auto seg = std::make_unique<HostSegment>();
        if (!load_segment(e->index_dir + "/segments/" + names[i], *seg, nt)) {
            if (e->idx) ns_index_abort(e->idx);
            return NS_ERR_IO;  // reference: reload() returns false, old segments stay (:82-85)
        }
        if (e->idx) {
            std::vector<uint64_t> begin(seg->rows.size());
            std::vector<uint32_t> count(seg->rows.size());
            for (size_t r = 0; r < seg->rows.size(); r++) {
                begin[r] = seg->rows[r].begin;
                count[r] = seg->rows[r].count;
            }
            int rc = ns_index_add_segment(e->idx, (uint32_t)i, (uint32_t)seg->doc_len.size(), seg->avgdl,
                                          seg->doc_len.data(), (uint32_t)seg->rows.size(), begin.data(), count.data(),
                                          seg->postings.data(), seg->postings.size());
            if (rc != NS_OK) {
                ns_index_abort(e->idx);
                return rc;
            }
            seg->drop_postings();  // they live in HBM now
        }
        loaded[i] = std::move(seg);
    }
    if (e->idx) {
        int rc = ns_index_commit(e->idx);
        if (rc != NS_OK) {
            ns_index_abort(e->idx);
            return rc;
        }
    }
    e->seg_names = std::move(names);
    e->segs = std::move(loaded);
    return NS_OK;
}

extern "C" int ns_engine_num_segments(const ns_engine* e) {
    if (!e) return 0;
    std::shared_lock<std::shared_mutex> lk(e->mu);
    return (int)e->seg_names.size();
}

extern "C" int ns_engine_segment_name(const ns_engine* e, int i, char* buf, size_t cap) {
    if (!e || !buf) return -1;
    std::shared_lock<std::shared_mutex> lk(e->mu);
    if (i < 0 || (size_t)i >= e->seg_names.size()) return -1;
    const std::string& s = e->seg_names[i];
    if (s.size() + 1 > cap) return -1;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

extern "C" int ns_engine_segment_stats(const ns_engine* e, int i, uint32_t* N, float* avgdl, uint32_t* T, uint64_t* P) {
    if (!e) return NS_ERR_INVALID;
    std::shared_lock<std::shared_mutex> lk(e->mu);
    if (i < 0 || (size_t)i >= e->segs.size() || !e->segs[i]) { set_error("segment not loaded by this engine"); return NS_ERR_INVALID; }
    const HostSegment& s = *e->segs[i];
    if (N) *N = s.N;
    if (avgdl) *avgdl = s.avgdl;
    if (T) *T = (uint32_t)s.rows.size();
    if (P) {
        uint64_t p = 0;
        for (auto& r : s.rows) p += r.count;
        *P = p;
    }
    return NS_OK;
}

extern "C" int ns_engine_term_stats(const ns_engine* e, int i, const char* term, uint32_t* df, uint32_t* count) {
    if (df) *df = 0;
    if (count) *count = 0;
    if (!e || !term) return NS_ERR_INVALID;
    std::shared_lock<std::shared_mutex> lk(e->mu);
    if (i < 0 || (size_t)i >= e->segs.size() || !e->segs[i]) { set_error("segment not loaded by this engine"); return NS_ERR_INVALID; }
    auto it = e->segs[i]->lex.find(term);
    if (it == e->segs[i]->lex.end()) return NS_OK;
    if (df) *df = e->segs[i]->rows[it->second].df;
    if (count) *count = e->segs[i]->rows[it->second].count;
    return NS_OK;
}

extern "C" ns_index* ns_engine_index(ns_engine* e) { return e ? e->idx : nullptr; }

extern "C" int ns_engine_cord_uid(const ns_engine* e, uint32_t seg, uint32_t doc, char* buf, size_t cap) {
    if (!e || !buf) return -1;
    std::shared_lock<std::shared_mutex> lk(e->mu);
    if (seg >= e->segs.size() || !e->segs[seg]) return -1;
    std::string s = e->segs[seg]->cord_uid(doc);
    if (s.size() + 1 > cap) return -1;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

namespace {

// One pass over the batch: tokenise, filter, look every term up in every owned segment's lexicon.
// The queries are independent, so the batch is cut into contiguous ranges, one per host thread
// (the reference serialises whole searches on Engine::mtx, src/api_engine.cpp:372).
struct Resolved {
    std::vector<uint64_t> q_off;   // [Q+1]
    std::vector<ns_qterm> terms;
    std::vector<uint8_t> has;      // [Q]
};

template <class GetQuery>
void resolve_all(ns_engine* e, uint32_t Q, GetQuery get, Resolved& out) {
    std::call_once(e->pool_once, [&] { e->pool.reset(new WorkPool(std::max(0, std::min(hw_threads(), 8) - 1))); });
    static const int max_nt = std::getenv("NSB200_HOST_THREADS") ? std::atoi(std::getenv("NSB200_HOST_THREADS")) : 1 << 20;
    // a sharded engine is one of `world` processes on the same box: take a 1/world share of the cores
    const int share = std::max(1, hw_threads() / std::max(1, e->world));
    const int nt = std::max(1, std::min(std::min(std::min(e->pool->workers() + 1, max_nt), share), (int)(Q / 256) + 1));
    std::vector<std::vector<ns_qterm>> per((size_t)nt);
    std::vector<uint32_t> cnt(Q, 0);
    out.has.assign(Q, 0);
    auto lo_of = [&](int t) { return (uint32_t)((uint64_t)Q * t / nt); };
    auto work = [&](int t) {
        per[t].reserve((size_t)(lo_of(t + 1) - lo_of(t)) * 4);
        for (uint32_t q = lo_of(t); q < lo_of(t + 1); q++) {
            const size_t before = per[t].size();
            out.has[q] = resolve_one(e, get(q), per[t]) ? 1 : 0;
            cnt[q] = (uint32_t)(per[t].size() - before);
        }
    };
    e->pool->run(nt, work);
    out.q_off.resize((size_t)Q + 1);
    uint64_t total = 0;
    out.q_off[0] = 0;
    for (uint32_t q = 0; q < Q; q++) {
        total += cnt[q];
        out.q_off[q + 1] = total;
    }
    out.terms.resize(std::max<uint64_t>(1, total));
    uint64_t at = 0;
    for (int t = 0; t < nt; t++) {
        if (!per[t].empty()) std::memcpy(out.terms.data() + at, per[t].data(), per[t].size() * sizeof(ns_qterm));
        at += per[t].size();
    }
}

// start of each NUL-terminated string in a packed buffer; false if fewer than Q strings fit
bool split_packed(const char* z, size_t nbytes, uint32_t Q, std::vector<const char*>& starts) {
    starts.resize(Q);
    size_t at = 0;
    for (uint32_t q = 0; q < Q; q++) {
        if (at >= nbytes) return false;
        starts[q] = z + at;
        const void* nul = std::memchr(z + at, 0, nbytes - at);
        if (!nul) return false;
        at = (size_t)((const char*)nul - z) + 1;
    }
    return true;
}

}  // namespace

extern "C" int ns_engine_resolve_batch(ns_engine* e, uint32_t Q, const char* const* queries, uint64_t* q_off,
                                       ns_qterm* terms, uint64_t terms_cap, uint64_t* n_terms, uint8_t* has_terms) {
    if (!e || (Q && !queries) || !q_off || !n_terms) { set_error("ns_engine_resolve_batch: null argument"); return NS_ERR_INVALID; }
    std::shared_lock<std::shared_mutex> lk(e->mu);
    Resolved r;
    resolve_all(e, Q, [&](uint32_t q) { return queries[q]; }, r);
    const uint64_t total = r.q_off[Q];
    std::memcpy(q_off, r.q_off.data(), ((size_t)Q + 1) * sizeof(uint64_t));
    *n_terms = total;
    if (has_terms && Q) std::memcpy(has_terms, r.has.data(), Q);
    if (!terms) return NS_OK;
    if (terms_cap < total) { set_error("ns_engine_resolve_batch: terms buffer too small"); return NS_ERR_INVALID; }
    if (total) std::memcpy(terms, r.terms.data(), total * sizeof(ns_qterm));
    return NS_OK;
}

extern "C" int ns_engine_resolve_batch_packed(ns_engine* e, uint32_t Q, const char* zqueries, size_t nbytes,
                                              uint64_t* q_off, ns_qterm* terms, uint64_t terms_cap, uint64_t* n_terms,
                                              uint8_t* has_terms) {
    if (!e || (Q && !zqueries) || !q_off || !n_terms) { set_error("ns_engine_resolve_batch_packed: null argument"); return NS_ERR_INVALID; }
    std::vector<const char*> starts;
    if (!split_packed(zqueries, nbytes, Q, starts)) {
        set_error("ns_engine_resolve_batch_packed: buffer holds fewer than Q NUL-terminated strings");
        return NS_ERR_INVALID;
    }
    std::shared_lock<std::shared_mutex> lk(e->mu);
    Resolved r;
    resolve_all(e, Q, [&](uint32_t q) { return starts[q]; }, r);
    const uint64_t total = r.q_off[Q];
    std::memcpy(q_off, r.q_off.data(), ((size_t)Q + 1) * sizeof(uint64_t));
    *n_terms = total;
    if (has_terms && Q) std::memcpy(has_terms, r.has.data(), Q);
    if (!terms) return NS_OK;
    if (terms_cap < total) { set_error("ns_engine_resolve_batch_packed: terms buffer too small"); return NS_ERR_INVALID; }
    if (total) std::memcpy(terms, r.terms.data(), total * sizeof(ns_qterm));
    return NS_OK;
}

extern "C" int ns_engine_search_batch(ns_engine* e, uint32_t Q, const char* const* queries, int k, ns_hit* out_hits,
                                      uint32_t* out_nhits, uint64_t* out_found, uint8_t* has_found) {
    if (!e || (Q && !queries)) { set_error("ns_engine_search_batch: null argument"); return NS_ERR_INVALID; }
    if (!e->idx) { set_error("engine was created without a CUDA device; there is no CPU search path"); return NS_ERR_STATE; }
    Resolved r;
    {
        std::shared_lock<std::shared_mutex> lk(e->mu);
        resolve_all(e, Q, [&](uint32_t q) { return queries[q]; }, r);
    }
    if (has_found && Q) std::memcpy(has_found, r.has.data(), Q);
    return ns_search_batch(e->idx, Q, k, r.q_off.data(), r.terms.data(), out_hits, out_nhits, out_found);
}

extern "C" int ns_engine_search_batch_packed(ns_engine* e, uint32_t Q, const char* zqueries, size_t nbytes, int k,
                                             ns_hit* out_hits, uint32_t* out_nhits, uint64_t* out_found,
                                             uint8_t* has_found) {
    if (!e || (Q && !zqueries)) { set_error("ns_engine_search_batch_packed: null argument"); return NS_ERR_INVALID; }
    if (!e->idx) { set_error("engine was created without a CUDA device; there is no CPU search path"); return NS_ERR_STATE; }
    std::vector<const char*> starts;
    if (!split_packed(zqueries, nbytes, Q, starts)) {
        set_error("ns_engine_search_batch_packed: buffer holds fewer than Q NUL-terminated strings");
        return NS_ERR_INVALID;
    }
    Resolved r;
    static const bool trace = std::getenv("NSB200_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    {
        std::shared_lock<std::shared_mutex> lk(e->mu);
        resolve_all(e, Q, [&](uint32_t q) { return starts[q]; }, r);
    }
    if (trace)
        std::fprintf(stderr, "[nsb200] Q=%u resolve %.3f ms (%llu terms)\n", Q,
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(),
                     (unsigned long long)r.q_off[Q]);
    if (has_found && Q) std::memcpy(has_found, r.has.data(), Q);
    return ns_search_batch(e->idx, Q, k, r.q_off.data(), r.terms.data(), out_hits, out_nhits, out_found);
}

extern "C" int ns_engine_search_json(ns_engine* e, const char* query, int k, char* buf, size_t cap, size_t* needed) {
    if (!e || !query) { set_error("ns_engine_search_json: null argument"); return NS_ERR_INVALID; }
    const int K = std::max(1, std::min(k, NS_MAX_K));  // src/api_engine.cpp:377
    std::vector<ns_hit> hits((size_t)K);
    uint32_t nh = 0;
    uint64_t found = 0;
    uint8_t has = 0;
    const char* qs[1] = {query};
    int rc = ns_engine_search_batch(e, 1, qs, K, hits.data(), &nh, &found, &has);
    if (rc != NS_OK) return rc;
    std::shared_lock<std::shared_mutex> lk(e->mu);
    // nlohmann::json objects dump with keys in lexicographic order
    std::string out = "{";
    if (has) out += "\"found\":" + std::to_string(found) + ",";
    out += "\"k\":" + std::to_string(K) + ",\"query\":\"";
    json_escape(query, out);
    out += "\",\"results\":[";
    for (uint32_t i = 0; i < nh; i++) {
        const ns_hit& h = hits[i];
        if (i) out += ",";
        out += "{\"cord_uid\":\"";
        if (h.seg < e->segs.size() && e->segs[h.seg]) json_escape(e->segs[h.seg]->cord_uid(h.doc), out);
        out += "\",\"docId\":" + std::to_string(h.doc) + ",\"score\":";
        json_double((double)h.score, out);
        out += ",\"segment\":\"";
        if (h.seg < e->seg_names.size()) json_escape(e->seg_names[h.seg], out);
        out += "\"}";
    }
    out += "],\"segments\":" + std::to_string(e->seg_names.size()) + "}";
    if (needed) *needed = out.size();
    if (buf && cap) {
        size_t n = std::min(cap - 1, out.size());
        std::memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return NS_OK;
}

// ------------------------------------------------------------------------------------------

extern "C" int ns_text_query_terms(const char* query, char* buf, size_t cap) {
    if (!query || !buf) return -1;
    std::vector<std::string> terms;
    query_terms(query, terms);
    size_t at = 0;
    for (auto& t : terms) {
        if (at + t.size() + 1 > cap) return -1;
        std::memcpy(buf + at, t.c_str(), t.size() + 1);
        at += t.size() + 1;
    }
    return (int)terms.size();
}

static CorpusSpec to_spec(const ns_corpus_spec* s) {
    CorpusSpec c;
    c.seed = s->seed;
    c.vocab = s->vocab;
    c.zipf_s = s->zipf_s;
    c.zipf_q = s->zipf_q;
    c.len_lo = s->len_lo;
    c.len_hi = s->len_hi;
    return c;
}

extern "C" int ns_corpus_write_segment(const ns_corpus_spec* spec, uint64_t doc_base, uint32_t ndocs, const char* segdir,
                                       int write_forward, const char* dump_path, int nthreads) {
    if (!spec || !segdir) { set_error("ns_corpus_write_segment: null argument"); return NS_ERR_INVALID; }
    GenSegment g;
    generate_segment(to_spec(spec), doc_base, ndocs, nthreads > 0 ? nthreads : hw_threads(), g);
    if (!write_segment_files(g, segdir, write_forward != 0)) return NS_ERR_IO;
    if (dump_path && !write_corpus_dump(g, dump_path)) { set_error("cannot write corpus dump"); return NS_ERR_IO; }
    return NS_OK;
}

extern "C" int ns_corpus_write_manifest(const char* index_dir, uint32_t nseg, const char* const* names) {
    if (!index_dir || (nseg && !names)) return NS_ERR_INVALID;
    if (!make_dirs(index_dir)) { set_error("cannot create index dir"); return NS_ERR_IO; }
    std::vector<std::string> v;
    for (uint32_t i = 0; i < nseg; i++) v.push_back(names[i]);
    return save_manifest(std::string(index_dir) + "/manifest.bin", v) ? NS_OK : NS_ERR_IO;
}

extern "C" int ns_corpus_make_queries(const ns_corpus_spec* spec, uint64_t query_seed, uint32_t nq, uint32_t min_terms,
                                      uint32_t max_terms, uint32_t head_ranks, char* buf, size_t cap, size_t* needed) {
    if (!spec) return NS_ERR_INVALID;
    auto qs = make_queries(to_spec(spec), query_seed, nq, min_terms, max_terms, head_ranks);
    size_t total = 0;
    for (auto& q : qs) total += q.size() + 1;
    if (needed) *needed = total;
    if (!buf) return NS_OK;
    if (cap < total) { set_error("ns_corpus_make_queries: buffer too small"); return NS_ERR_INVALID; }
    size_t at = 0;
    for (auto& q : qs) {
        std::memcpy(buf + at, q.c_str(), q.size() + 1);
        at += q.size() + 1;
    }
    return NS_OK;
}
