// Host mirror of cord19::Engine for the search path (reference: include/api_engine.hpp:23-91,
// src/api_engine.cpp:50-162 and :369-542), plus the corpus tooling entry points.
// Tokenise / stop-filter / lexicon lookup / IDF stay on the host exactly as in the reference;
// the per-posting work goes to the GPU(s) through the device layer (device_api.cu).
//
// Differences in shape, all deliberate:
//   * one engine may span several GPUs of the box (segment j of the engine's share lives on device
//     j % ndev); a batch is tokenised and resolved ONCE, fanned out, scored on every device, the per-device
//     result blobs are stored into device 0's gather buffer by the score kernels themselves (peer memory)
//     and merged there — the reference's loop over segments (src/api_engine.cpp:441-505) spread out;
//   * everything a search needs — segment names, lexicons, cord_uids, metadata, embeddings and the
//     device arrays — forms one immutable GENERATION that a call snapshots once; reload builds a new one
//     and swaps it in (the reference holds Engine::mtx for the whole search, :372);
//   * single queries may be coalesced into batches by a dispatcher (ns_engine_coalescer_start).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/nextsearch_b200.h"
#include "../device_internal.hpp"
#include "common.hpp"
#include "corpus.hpp"
#include "json_text.hpp"
#include "metadata.hpp"
#include "segment_io.hpp"
#include "semantic.hpp"
#include "textutil.hpp"
#include "workpool.hpp"

using namespace nsb;

namespace {

thread_local bool tl_scan_failed = false;  // a device similarity scan of this thread's current query failed

int hw_threads() {
    unsigned n = std::thread::hardware_concurrency();
    return n == 0 ? 1 : (int)n;
}

// One committed state of the engine.  Immutable after publication.
struct Generation {
    std::vector<std::string> seg_names;
    std::vector<std::unique_ptr<HostSegment>> segs;  // loaded only for segments this engine owns
    std::vector<int> seg_dev;                         // device slot of each owned segment, -1 otherwise
    TermDict dict;
    std::vector<std::vector<uint32_t>> dev_cols;      // per device slot: the dict columns (owned segments) it holds, ascending
    std::vector<std::shared_ptr<const void>> dev_state;  // per device slot: the committed device index
    MetaIndex meta;
    SemanticIndex sem;
    std::shared_ptr<ns_semantic> sem_dev;  // the embeddings on device slot 0 (similarity scans of the expansion)
    double load_seconds = 0, upload_seconds = 0;
    uint64_t posting_bytes = 0;
};

struct ReloadStats {
    double total_s = 0, read_s = 0, dict_s = 0;
    uint64_t posting_bytes = 0;
};

// Exchange group of a multi-device engine: one ns_exchange per device, every device publishes into the
// root's (device slot 0) gather buffer; slots = 1 because a call owns its group until it has fetched.
struct XGroup {
    std::vector<ns_exchange*> x;
    uint32_t max_q = 0;
    uint64_t step = 0;
    ~XGroup() {
        for (auto* e : x)
            if (e) ns_exchange_destroy(e);
    }
};

struct Coalescer;

// One host thread per device slot of a multi-device engine; every CUDA call that touches device d outside
// device 0 — prepare, launch, batch teardown — runs on thread d.  A host thread's first CUDA call on a device
// binds it to that device's context, which costs milliseconds; with caller threads (or pool workers) fanning
// out to all devices themselves that price is paid once per (thread, device) PAIR and keeps recurring as new
// threads appear.  With device threads it is paid ndev times, at start-up.
class DeviceWorker {
  public:
    DeviceWorker() : th_([this] { loop(); }) {}
    ~DeviceWorker() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        th_.join();
    }
    void submit(std::function<void()> fn) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            q_.push_back(std::move(fn));
        }
        cv_.notify_one();
    }

  private:
    void loop() {
        for (;;) {
            std::function<void()> fn;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;  // stop requested and nothing left
                fn = std::move(q_.front());
                q_.pop_front();
            }
            fn();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    bool stop_ = false;
    std::thread th_;
};

// counts down to zero; wait() returns then
class Latch {
  public:
    explicit Latch(int n) : n_(n) {}
    void done() {
        std::lock_guard<std::mutex> lk(mu_);
        if (--n_ == 0) cv_.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return n_ == 0; });
    }

  private:
    std::mutex mu_;
    std::condition_variable cv_;
    int n_;
};

}  // namespace

struct ns_engine {
    std::string index_dir;
    std::vector<int> devices;       // CUDA ordinals, empty = host-only engine
    std::vector<ns_index*> idx;     // one per device slot
    std::vector<std::unique_ptr<DeviceWorker>> dev_threads;  // multi-device engines: one per device slot
    int rank = 0, world = 1;        // process-level share: segment i is owned when i % world == rank
    bool trace = false;             // NSB200_TRACE (read at create): per-phase host timings of every search on stderr
    bool keep_raw = false;          // NSB200_KEEP_RAW (read at create): keep {docId, tf} next to the resident scores, so that
                                    // batches through the raw ABI may name a row with a foreign idf
    std::mutex gen_mu;              // guards `gen` (the pointer only)
    std::shared_ptr<const Generation> gen;
    std::mutex reload_mu;           // one reload at a time
    std::unique_ptr<WorkPool> pool;
    std::once_flag pool_once;
    std::mutex xg_mu;
    std::vector<std::unique_ptr<XGroup>> xg_pool;
    std::mutex sc_mu;
    std::vector<std::shared_ptr<void>> sc_pool;  // ResolveScratch objects (type-erased: defined further down)
    std::shared_ptr<Coalescer> coalescer;  // callers copy the pointer under co_mu: a stop cannot free it under them
    std::mutex co_mu;               // guards `coalescer` start/stop
    ReloadStats last_reload;

    std::shared_ptr<const Generation> snapshot() {
        std::lock_guard<std::mutex> lk(gen_mu);
        return gen;
    }
    bool owns(size_t i) const { return (int)(i % (size_t)world) == rank; }
    WorkPool& workers() {
        std::call_once(pool_once, [&] {
            int n = std::min(hw_threads(), 16);
            if (const char* s = std::getenv("NSB200_HOST_THREADS")) n = std::max(1, std::atoi(s));
            // a sharded engine is one of `world` processes on the same box: take a 1/world share of the cores
            n = std::max(1, std::min(n, std::max(1, hw_threads() / std::max(1, world))));
            pool.reset(new WorkPool(std::max(0, n - 1)));
        });
        return *pool;
    }
};

namespace {

// ------------------------------------------------------------------------------------------
// reload
// ------------------------------------------------------------------------------------------

// Pinned staging of one segment's postings: every barrel file that has been read starts its H2D copy.
struct UploadSink : PostingSink {
    ns_index* idx;
    ns_upload* up = nullptr;
    int rc = NS_OK;
    explicit UploadSink(ns_index* i) : idx(i) {}
    uint8_t* begin(uint64_t P) override {
        rc = ns_upload_begin(idx, P, &up);
        return rc == NS_OK ? (uint8_t*)ns_upload_buffer(up) : nullptr;
    }
    void filled(uint64_t first, uint64_t count) override { ns_upload_push(up, first, count); }
    ~UploadSink() override {
        if (up) ns_upload_abort(up);
    }
};

int do_reload(ns_engine* e) {
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    auto g = std::make_shared<Generation>();
    g->seg_names = discover_segments(e->index_dir);
    if (g->seg_names.empty()) { set_error("no segments under " + e->index_dir); return NS_ERR_IO; }  // src/api_engine.cpp:73
    const size_t nseg = g->seg_names.size();
    const int ndev = (int)e->devices.size();
    g->segs.resize(nseg);
    g->seg_dev.assign(nseg, -1);
    g->dev_cols.resize((size_t)std::max(1, ndev));
    auto abort_all = [&]() {
        for (auto* ix : e->idx) ns_index_abort(ix);
    };
    const int nt = hw_threads();
    uint32_t owned_j = 0;
    double read_s = 0;
    for (size_t i = 0; i < nseg; i++) {
        if (!e->owns(i)) continue;
        const int d = ndev ? (int)(owned_j % (uint32_t)ndev) : -1;
        g->seg_dev[i] = d;
        g->dev_cols[(size_t)std::max(0, d)].push_back(owned_j);
        owned_j++;
        auto seg = std::make_unique<HostSegment>();
        const std::string dir = e->index_dir + "/segments/" + g->seg_names[i];
        const auto r0 = clk::now();
        if (d < 0) {
            if (!load_segment(dir, *seg, nt)) return NS_ERR_IO;
            seg->drop_postings();  // a host-only engine resolves queries; it never scores
        } else {
            UploadSink sink(e->idx[(size_t)d]);
            if (!load_segment(dir, *seg, nt, &sink)) {
                abort_all();
                return sink.rc != NS_OK ? sink.rc : NS_ERR_IO;  // reference: reload() returns false, old segments stay (:82-85)
            }
            read_s += std::chrono::duration<double>(clk::now() - r0).count();
            std::vector<uint64_t> begin(seg->rows.size());
            std::vector<uint32_t> count(seg->rows.size());
            std::vector<float> idf(seg->rows.size());
            for (size_t r = 0; r < seg->rows.size(); r++) {
                begin[r] = seg->rows[r].begin;
                count[r] = seg->rows[r].count;
                idf[r] = seg->rows[r].idf;  // bm25_idf(N, df): the resident scores are built with what queries will ask for
            }
            ns_upload* up = sink.up;
            sink.up = nullptr;  // finish frees the ticket
            if (!up) {
                abort_all();
                set_error("segment loader did not open an upload for " + dir);
                return NS_ERR_STATE;
            }
            int rc = ns_upload_finish(up, (uint32_t)i, (uint32_t)seg->doc_len.size(), seg->avgdl, seg->doc_len.data(),
                                      (uint32_t)seg->rows.size(), begin.data(), count.data(), idf.data(),
                                      e->keep_raw ? 0u : NS_SEG_DROP_RAW);
            if (rc != NS_OK) {
                abort_all();
                return rc;
            }
            for (auto& r : seg->rows) g->posting_bytes += 8ull * r.count;
        }
        g->segs[i] = std::move(seg);
    }
    for (auto* ix : e->idx) {
        int rc = ns_index_commit(ix);
        if (rc != NS_OK) {
            abort_all();
            return rc;
        }
    }
    for (auto* ix : e->idx) g->dev_state.push_back(index_live_state(ix));
    const auto t1 = clk::now();
    g->dict.build(g->segs);
    const auto t2 = clk::now();
    // result decoration and semantic expansion data (src/api_engine.cpp:110-152)
    g->meta.load(e->index_dir + "/metadata.csv");
    {
        std::string emb;
        if (const char* p = std::getenv("EMBEDDINGS_PATH")) emb = p;
        else
            for (const char* nm : {"embeddings.vec", "embeddings.txt", "glove.txt", "vectors.txt"})
                if (emb.empty() && file_exists(e->index_dir + "/" + nm)) emb = e->index_dir + "/" + nm;
        if (!emb.empty() && file_exists(emb)) {
            const TermDict& dict = g->dict;
            g->sem.load(emb, [&](const std::string& w) { return dict.find(w.data(), w.size(), term_hash(w.data(), w.size())) >= 0; });
        }
        if (g->sem.enabled && !e->devices.empty() && !std::getenv("NSB200_SEMANTIC_ON_HOST")) {
            // the similarity scans run on the GPU (cosine_scan_kernel); sims are bit-identical to the host loop's
            ns_semantic* h = nullptr;
            if (ns_semantic_upload(e->devices[0], (uint32_t)g->sem.terms.size(), (uint32_t)g->sem.dim, g->sem.vecs.data(), &h) == NS_OK) {
                g->sem_dev.reset(h, [](ns_semantic* p) { ns_semantic_destroy(p); });
                const SemanticIndex* si = &g->sem;
                g->sem.device_scan = [h, si](const float* qvecs, uint32_t M, float min_sim,
                                             std::vector<std::vector<SemanticIndex::Survivor>>& out) {
                    constexpr uint32_t cap = 4096;
                    std::vector<uint32_t> rows((size_t)M * cap), count(M);
                    std::vector<float> sims((size_t)M * cap);
                    out.resize(M);
                    (void)si;
                    uint32_t use = cap;
                    bool ok = ns_semantic_scan(h, M, qvecs, min_sim, use, rows.data(), sims.data(), count.data()) == NS_OK;
                    uint32_t most = 0;
                    for (uint32_t m = 0; ok && m < M; m++) most = std::max(most, count[m]);
                    if (ok && most > use) {  // more survivors than room: scan again, on the device, with room for all
                        use = most;
                        rows.assign((size_t)M * use, 0u);
                        sims.assign((size_t)M * use, 0.0f);
                        ok = ns_semantic_scan(h, M, qvecs, min_sim, use, rows.data(), sims.data(), count.data()) == NS_OK;
                    }
                    if (!ok) {  // no host fallback behind a GPU engine: the search that asked fails with the CUDA error
                        tl_scan_failed = true;
                        for (auto& v : out) v.clear();
                        return;
                    }
                    const uint32_t cap_used = use;
                    for (uint32_t m = 0; m < M; m++) {
                        const uint32_t cap = cap_used;
                        out[m].resize(count[m]);
                        for (uint32_t i = 0; i < count[m]; i++) out[m][i] = SemanticIndex::Survivor{rows[(size_t)m * cap + i], sims[(size_t)m * cap + i]};
                        std::sort(out[m].begin(), out[m].end(), [](const SemanticIndex::Survivor& a, const SemanticIndex::Survivor& b) { return a.row < b.row; });
                    }
                };
            }
        }
    }
    ReloadStats rs;
    rs.total_s = std::chrono::duration<double>(clk::now() - t0).count();
    rs.read_s = read_s;
    rs.dict_s = std::chrono::duration<double>(t2 - t1).count();
    rs.posting_bytes = g->posting_bytes;
    std::lock_guard<std::mutex> lk(e->gen_mu);
    e->last_reload = rs;  // read by ns_engine_reload_stats under the same lock
    e->gen = g;  // in-flight calls keep the generation they started with
    return NS_OK;
}

// ------------------------------------------------------------------------------------------
// front end: query text -> (segment, row, idf, weight) tuples, once per batch
// ------------------------------------------------------------------------------------------

struct QueryTerm {
    uint32_t gid;  // dictionary id
    float w;
};

// Terms of one query in scoring order.  Plain search: the tokens that survive the filter, weight 1.0f
// (src/api_engine.cpp:388-397, 418-421).  With embeddings: SemanticIndex::expand's list (:410-417).
// Returns whether the reference would compute "found" (:407, :424).
bool query_terms_of(const Generation& g, const char* query, std::vector<QueryTerm>& out) {
    static thread_local std::string buf;
    static thread_local std::vector<TokSpan> spans;
    out.clear();
    query_term_spans(query, buf, spans);
    if (spans.empty() || g.seg_names.empty()) return false;
    if (!g.sem.enabled) {
        // hash every token and ask for its dictionary slot first, probe afterwards: the cache misses of one
        // query's tokens overlap instead of following each other
        static thread_local std::vector<uint64_t> hashes;
        hashes.resize(spans.size());
        for (size_t i = 0; i < spans.size(); i++) {
            hashes[i] = term_hash(buf.data() + spans[i].off, spans[i].len);
            g.dict.prefetch(hashes[i]);
        }
        const size_t ncol = g.dict.owned.size();
        for (size_t i = 0; i < spans.size(); i++) {
            const int64_t gid = g.dict.find(buf.data() + spans[i].off, spans[i].len, hashes[i]);
            if (gid < 0) continue;
            __builtin_prefetch(g.dict.table.data() + (size_t)gid * ncol);
            out.push_back(QueryTerm{(uint32_t)gid, 1.0f});
        }
        return true;
    }
    std::vector<std::string> base;
    base.reserve(spans.size());
    for (const TokSpan& s : spans) base.emplace_back(buf.data() + s.off, s.len);
    const auto qw = g.sem.expand(base);
    if (qw.empty()) return false;
    for (const auto& tw : qw) {
        const int64_t gid = g.dict.find(tw.first.data(), tw.first.size(), term_hash(tw.first.data(), tw.first.size()));
        if (gid >= 0) out.push_back(QueryTerm{(uint32_t)gid, tw.second});
    }
    return true;
}

// Emit the query's terms for the dictionary columns `cols` (ascending segments): ordered by
// (segment asc, query-term order), duplicates kept (:391-397, :449).
inline void emit_terms(const Generation& g, const std::vector<QueryTerm>& qt, const std::vector<uint32_t>& cols,
                       std::vector<ns_qterm>& out) {
    const size_t ncol = g.dict.owned.size();
    for (uint32_t j : cols) {
        const uint32_t seg = g.dict.owned[j];
        for (const QueryTerm& t : qt) {
            const TermDict::Entry& en = g.dict.table[(size_t)t.gid * ncol + j];
            if (en.row == TermDict::kAbsent) continue;  // :454-458
            out.push_back(ns_qterm{seg, en.row, en.idf, t.w});
        }
    }
}

struct Resolved {
    std::vector<uint64_t> q_off;  // [Q+1]
    std::vector<ns_qterm> terms;
};

// One pass over the batch on the engine's host threads.  parts[p] receives the terms of the dictionary
// columns colsets[p] (one part per device, or a single part with every owned segment).
template <class TermsOf>
void resolve_all(ns_engine* e, const Generation& g, uint32_t Q, TermsOf terms_of,
                 const std::vector<const std::vector<uint32_t>*>& colsets, std::vector<Resolved>& parts,
                 std::vector<uint8_t>& has) {
    const size_t np = colsets.size();
    WorkPool& pool = e->workers();
    const int nt = std::max(1, std::min(pool.workers() + 1, (int)(Q / 128) + 1));
    std::vector<std::vector<std::vector<ns_qterm>>> per((size_t)nt, std::vector<std::vector<ns_qterm>>(np));
    std::vector<std::vector<uint32_t>> cnt(np, std::vector<uint32_t>(Q, 0));
    has.assign(Q, 0);
    auto lo_of = [&](int t) { return (uint32_t)((uint64_t)Q * t / nt); };
    auto work = [&](int t) {
        std::vector<QueryTerm> qt;
        for (size_t p = 0; p < np; p++) per[t][p].reserve((size_t)(lo_of(t + 1) - lo_of(t)) * 4);
        for (uint32_t q = lo_of(t); q < lo_of(t + 1); q++) {
            has[q] = terms_of(q, qt) ? 1 : 0;
            for (size_t p = 0; p < np; p++) {
                const size_t before = per[t][p].size();
                emit_terms(g, qt, *colsets[p], per[t][p]);
                cnt[p][q] = (uint32_t)(per[t][p].size() - before);
            }
        }
    };
    pool.run(nt, work);
    parts.assign(np, Resolved{});
    for (size_t p = 0; p < np; p++) {
        Resolved& r = parts[p];
        r.q_off.resize((size_t)Q + 1);
        uint64_t total = 0;
        r.q_off[0] = 0;
        for (uint32_t q = 0; q < Q; q++) {
            total += cnt[p][q];
            r.q_off[q + 1] = total;
        }
        r.terms.resize(std::max<uint64_t>(1, total));
        uint64_t at = 0;
        for (int t = 0; t < nt; t++) {
            if (!per[t][p].empty()) std::memcpy(r.terms.data() + at, per[t][p].data(), per[t][p].size() * sizeof(ns_qterm));
            at += per[t][p].size();
        }
    }
}

// The same pass in the device layer's own form (batch_prepare_trusted): per device, the terms as kernel
// records with the slot the segment occupies on that device, the per-query posting totals and the batch flags —
// everything the generic ns_batch_prepare would have to look up again per term.
struct DevResolved {
    std::vector<uint32_t> qoff;        // [Q+1]
    std::vector<PreparedTerm> terms;
    std::vector<uint64_t> weight;      // [Q]
    uint32_t max_in_seg = 0;
    bool unit = true, scan_always = false;
};

// Buffers of one front-end pass, pooled by the engine: their capacity survives from call to call.  (Fresh
// megabyte-sized vectors per call are mmap'ed and unmapped by malloc every time, which serialises concurrent
// callers on the process's address-space lock.)
struct ResolveScratch {
    struct ThreadOut {
        std::vector<PreparedTerm> terms;
        uint32_t max_in_seg = 0;
        bool unit = true, scan_always = false;
    };
    std::vector<DevResolved> parts;
    std::vector<std::vector<ThreadOut>> per;
    std::vector<uint8_t> has;
    std::vector<uint8_t> wide;   // [Q] 1: some (query, segment) has more than 32 terms (needs a wide kernel variant)
    std::atomic<int> failed{0};
};

template <class TermsOf>
void resolve_devices(ns_engine* e, const Generation& g, uint32_t Q, TermsOf terms_of, ResolveScratch& sc) {
    using ThreadOut = ResolveScratch::ThreadOut;
    std::vector<DevResolved>& parts = sc.parts;
    std::vector<uint8_t>& has = sc.has;
    const size_t np = g.dev_cols.size();
    const size_t ncol = g.dict.owned.size();
    WorkPool& pool = e->workers();
    const int nt = std::max(1, std::min(pool.workers() + 1, (int)(Q / 128) + 1));
    std::vector<std::vector<ThreadOut>>& per = sc.per;
    if (per.size() < (size_t)nt) per.resize((size_t)nt);
    for (int t = 0; t < nt; t++) {
        per[t].resize(np);
        for (auto& o : per[t]) {
            o.terms.clear();
            o.max_in_seg = 0;
            o.unit = true;
            o.scan_always = false;
        }
    }
    parts.resize(np);
    for (size_t p = 0; p < np; p++) {
        parts[p].qoff.assign((size_t)Q + 1, 0);  // filled with counts first, turned into offsets below
        parts[p].weight.assign(Q, 0);
        parts[p].max_in_seg = 0;
        parts[p].unit = true;
        parts[p].scan_always = false;
    }
    has.assign(Q, 0);
    sc.wide.assign(Q, 0);
    sc.failed = 0;
    auto lo_of = [&](int t) { return (uint32_t)((uint64_t)Q * t / nt); };
    auto work_body = [&](int t) {
        std::vector<QueryTerm> qt;
        for (size_t p = 0; p < np; p++) per[t][p].terms.reserve((size_t)(lo_of(t + 1) - lo_of(t)) * 4 * std::max<size_t>(1, g.dev_cols[p].size()));
        for (uint32_t q = lo_of(t); q < lo_of(t + 1); q++) {
            tl_scan_failed = false;
            has[q] = terms_of(q, qt) ? 1 : 0;
            if (tl_scan_failed) sc.failed = 1;
            for (size_t p = 0; p < np; p++) {
                ThreadOut& o = per[t][p];
                const size_t before = o.terms.size();
                uint64_t wsum = 0;
                const std::vector<uint32_t>& cols = g.dev_cols[p];
                for (uint32_t slot = 0; slot < (uint32_t)cols.size(); slot++) {
                    const uint32_t j = cols[slot];
                    uint32_t in_seg = 0;
                    for (const QueryTerm& x : qt) {
                        const TermDict::Entry& en = g.dict.table[(size_t)x.gid * ncol + j];
                        if (en.row == TermDict::kAbsent || en.count == 0) continue;  // src/api_engine.cpp:454-458
                        o.terms.push_back(PreparedTerm{slot, en.row, en.idf, x.w, 0u, 0u});
                        wsum += en.count;
                        in_seg++;
                        if (!(x.w >= 0.0f) || !(en.idf >= 0.0f)) o.scan_always = true;
                        if (x.w != 1.0f || !(en.idf >= 9.094947017729282e-13f && en.idf <= 64.0f)) o.unit = false;
                    }
                    o.max_in_seg = std::max(o.max_in_seg, in_seg);
                    if (in_seg > 32u) sc.wide[q] = 1;  // (a query is resolved by exactly one thread)
                }
                parts[p].qoff[q + 1] = (uint32_t)(o.terms.size() - before);
                parts[p].weight[q] = wsum;
            }
        }
    };
    auto work = [&](int t) {   // pool threads: an exception (out of memory) must not escape them
        try {
            work_body(t);
        } catch (...) {
            sc.failed = 2;
        }
    };
    pool.run(nt, work);
    if (sc.failed == 2) return;  // the caller reports it; the partial buffers are not used
    for (size_t p = 0; p < np; p++) {
        DevResolved& r = parts[p];
        for (uint32_t q = 0; q < Q; q++) r.qoff[q + 1] += r.qoff[q];
        r.terms.resize(std::max<uint32_t>(1, r.qoff[Q]));
        size_t at = 0;
        for (int t = 0; t < nt; t++) {
            const ThreadOut& o = per[t][p];
            if (!o.terms.empty()) std::memcpy(r.terms.data() + at, o.terms.data(), o.terms.size() * sizeof(PreparedTerm));
            at += o.terms.size();
            r.max_in_seg = std::max(r.max_in_seg, o.max_in_seg);
            r.unit = r.unit && o.unit;
            r.scan_always = r.scan_always || o.scan_always;
        }
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

// ------------------------------------------------------------------------------------------
// scoring: one device, or all devices + peer exchange
// ------------------------------------------------------------------------------------------

int acquire_group(ns_engine* e, uint32_t Q, std::unique_ptr<XGroup>& out) {
    {
        std::lock_guard<std::mutex> lk(e->xg_mu);
        for (size_t i = 0; i < e->xg_pool.size(); i++) {
            if (e->xg_pool[i]->max_q >= Q) {
                out = std::move(e->xg_pool[i]);
                e->xg_pool.erase(e->xg_pool.begin() + (long)i);
                return NS_OK;
            }
        }
    }
    auto g = std::make_unique<XGroup>();
    uint32_t cap = 4096;
    while (cap < Q) cap <<= 1;
    g->max_q = cap;
    const uint32_t ndev = (uint32_t)e->devices.size();
    g->x.assign(ndev, nullptr);
    for (uint32_t d = 0; d < ndev; d++) {
        int rc = d == 0 ? ns_exchange_create(e->devices[d], ndev, d, cap, 1, &g->x[d])
                        : exchange_create_publisher(e->devices[d], ndev, d, cap, 1, &g->x[d]);
        if (rc != NS_OK) return rc;
    }
    for (uint32_t d = 0; d < ndev; d++) {  // everybody publishes into the root's gather buffer; the root receives
        int rc = ns_exchange_attach_local(g->x[d], g->x[0]);
        if (rc != NS_OK) return rc;
    }
    out = std::move(g);
    return NS_OK;
}

void release_group(ns_engine* e, std::unique_ptr<XGroup> g) {
    std::lock_guard<std::mutex> lk(e->xg_mu);
    if (e->xg_pool.size() < 64) e->xg_pool.push_back(std::move(g));
}

// Everything after the front end: prepare, launch, (exchange,) fetch of one resolved batch of Q queries.
int run_resolved(ns_engine* e, const Generation& g, const std::vector<DevResolved>& parts, uint32_t Q, int k, ns_hit* out_hits,
                 uint32_t* out_nhits, uint64_t* out_found, double resolve_ms) {
    const size_t ndev = e->idx.size();
    using clk = std::chrono::steady_clock;
    const auto t1 = clk::now();
    auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    // prepare on device slot d: the trusted form; if the device index carries no resident scores
    // (NSB200_NO_RESIDENT, or a segment whose rows share postings) the same terms go through the generic, validating
    // ns_batch_prepare
    auto prepare = [&](size_t d, ns_batch** out) -> int {
        const DevResolved& r = parts[d];
        PreparedBatch pb{r.qoff.data(), r.terms.data(), r.weight.data(), r.max_in_seg, r.unit, r.scan_always};
        int rc = batch_prepare_trusted(e->idx[d], g.dev_state[d], Q, k, pb, out);
        if (rc != NS_ERR_STATE) return rc;
        std::vector<uint64_t> q_off((size_t)Q + 1);
        std::vector<ns_qterm> qt(std::max<size_t>(1, r.qoff[Q]));
        for (uint32_t q = 0; q <= Q; q++) q_off[q] = r.qoff[q];
        for (uint32_t i = 0; i < r.qoff[Q]; i++)
            qt[i] = ns_qterm{g.dict.owned[g.dev_cols[d][r.terms[i].slot]], r.terms[i].row, r.terms[i].idf, r.terms[i].w};
        return batch_prepare_on(e->idx[d], g.dev_state[d], Q, k, q_off.data(), qt.data(), out);
    };

    if (ndev == 1) {
        ns_batch* b = nullptr;
        int rc = prepare(0, &b);
        if (rc != NS_OK) return rc;
        const auto t2 = clk::now();
        rc = ns_batch_launch(b, nullptr);
        if (rc == NS_OK) rc = ns_batch_fetch(b, out_hits, out_nhits, out_found);
        const auto t3 = clk::now();
        ns_batch_destroy(b);
        if (e->trace)
            std::fprintf(stderr, "[nsb200] Q=%u resolve %.3f ms, prepare %.3f, launch+fetch %.3f, destroy %.3f\n", Q, resolve_ms,
                         ms(t1, t2), ms(t2, t3), ms(t3, clk::now()));
        return rc;
    }

    // several devices: prepare + launch on every device in parallel; each score kernel stores its per-query
    // results into the root's gather buffer; the root merges once all score kernels are done.
    std::unique_ptr<XGroup> grp;
    int rc = acquire_group(e, std::max<uint32_t>(1, Q), grp);
    if (rc != NS_OK) return rc;
    const uint64_t step = grp->step++;
    std::vector<ns_batch*> bs(ndev, nullptr);
    std::vector<int> rcs(ndev, NS_OK);
    std::vector<std::string> errs(ndev);
    std::vector<double> t_prep(ndev, 0.0), t_launch(ndev, 0.0);
    auto one = [&](int d) {   // runs on device thread d: nothing may escape it (the caller waits on the latch)
        rcs[d] = abi_guard("device thread", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int {
            const auto a = clk::now();
            int r = prepare((size_t)d, &bs[d]);
            const auto b = clk::now();
            if (r == NS_OK) r = ns_batch_launch_exchange(bs[d], grp->x[d], step, nullptr);
            t_prep[d] = ms(a, b);
            t_launch[d] = ms(b, clk::now());
            return r;
        });
        if (rcs[d] != NS_OK) {
            try { errs[d] = ns_last_error(); } catch (...) {}
        }
    };
    {
        Latch latch((int)ndev);
        for (size_t d = 0; d < ndev; d++)
            e->dev_threads[d]->submit([&, d] {
                one((int)d);
                latch.done();
            });
        latch.wait();
    }
    const auto t2 = clk::now();
    for (size_t d = 0; d < ndev && rc == NS_OK; d++)
        if (rcs[d] != NS_OK) {
            rc = rcs[d];
            set_error(errs[d]);
        }
    if (rc == NS_OK) rc = exchange_root_merge(grp->x[0], bs.data(), (int)ndev, step, Q, k);
    const auto t2b = clk::now();
    if (rc == NS_OK) rc = ns_exchange_fetch(grp->x[0], step, Q, k, out_hits, out_nhits, out_found);
    const auto t3 = clk::now();
    std::string keep = rc != NS_OK ? std::string(ns_last_error()) : std::string();
    // Teardown (waits for that device's kernels, returns the buffers to the device's pool) on the device threads;
    // nothing below depends on it.  The batches hold their generation's device state alive until then.
    // A FAILED call drops its exchange group below, and the other devices' score kernels may still be storing
    // into the root's gather buffer: the teardowns (which wait for those kernels) must have run before the
    // group's memory is freed.
    int live = 0;
    for (size_t d = 0; d < ndev; d++) live += bs[d] ? 1 : 0;
    auto torn_down = std::make_shared<Latch>(live);
    for (size_t d = 0; d < ndev; d++)
        if (bs[d]) {
            ns_batch* b = bs[d];
            e->dev_threads[d]->submit([b, torn_down] {
                ns_batch_destroy(b);
                torn_down->done();
            });
        }
    if (rc != NS_OK && live) torn_down->wait();
    if (e->trace)
    {
        double sp = 0, sl = 0;
        for (size_t d = 0; d < ndev; d++) {
            sp += t_prep[d];
            sl += t_launch[d];
        }
        std::fprintf(stderr,
                     "[nsb200] Q=%u ndev=%zu resolve %.3f ms, prepare+launch %.3f (sum over devices: prepare %.3f, launch %.3f), "
                     "root merge enqueue %.3f, fetch(wait) %.3f, destroy %.3f\n",
                     Q, ndev, resolve_ms, ms(t1, t2), sp, sl, ms(t2, t2b), ms(t2b, t3), ms(t3, clk::now()));
    }
    if (rc == NS_OK) release_group(e, std::move(grp));  // a failed group is dropped: its flags may be in any state
    else set_error(keep);
    return rc;
}

// The queries `idx` of a resolved batch as a batch of their own (same per-device layout).
void subset_parts(const std::vector<DevResolved>& parts, const std::vector<uint32_t>& idx, uint32_t max_in_seg_cap,
                  std::vector<DevResolved>& out) {
    out.assign(parts.size(), DevResolved{});
    for (size_t p = 0; p < parts.size(); p++) {
        const DevResolved& r = parts[p];
        DevResolved& o = out[p];
        o.qoff.assign(idx.size() + 1, 0);
        o.weight.resize(idx.size());
        size_t total = 0;
        for (uint32_t q : idx) total += r.qoff[q + 1] - r.qoff[q];
        o.terms.resize(std::max<size_t>(1, total));
        size_t at = 0;
        for (size_t i = 0; i < idx.size(); i++) {
            const uint32_t q = idx[i];
            const size_t n = r.qoff[q + 1] - r.qoff[q];
            if (n) std::memcpy(o.terms.data() + at, r.terms.data() + r.qoff[q], n * sizeof(PreparedTerm));
            at += n;
            o.qoff[i + 1] = (uint32_t)at;
            o.weight[i] = r.weight[q];
        }
        o.max_in_seg = std::min(r.max_in_seg, max_in_seg_cap);
        o.unit = r.unit;                // flags of the whole batch: conservative for a part of it
        o.scan_always = r.scan_always;
    }
}

template <class TermsOf>
int search_core(ns_engine* e, const std::shared_ptr<const Generation>& gen, uint32_t Q, TermsOf terms_of, int k,
                ns_hit* out_hits, uint32_t* out_nhits, uint64_t* out_found, uint8_t* has_found) {
    if (e->idx.empty()) { set_error("engine was created without a CUDA device; there is no CPU search path"); return NS_ERR_STATE; }
    if (!gen) { set_error("search before a successful reload"); return NS_ERR_STATE; }
    const Generation& g = *gen;
    std::shared_ptr<ResolveScratch> scratch;
    {
        std::lock_guard<std::mutex> lk(e->sc_mu);
        if (!e->sc_pool.empty()) {
            scratch = std::static_pointer_cast<ResolveScratch>(e->sc_pool.back());
            e->sc_pool.pop_back();
        }
    }
    if (!scratch) scratch = std::make_shared<ResolveScratch>();
    struct GiveBack {
        ns_engine* e;
        std::shared_ptr<ResolveScratch>& s;
        ~GiveBack() {
            std::lock_guard<std::mutex> lk(e->sc_mu);
            if (e->sc_pool.size() < 32) e->sc_pool.push_back(std::static_pointer_cast<void>(s));
        }
    } give_back{e, scratch};
    using clk = std::chrono::steady_clock;
    const auto t0 = clk::now();
    resolve_devices(e, g, Q, terms_of, *scratch);
    if (scratch->failed == 2) { set_error("front end: out of host memory"); return NS_ERR_NOMEM; }
    if (scratch->failed) { set_error("semantic expansion: the device similarity scan failed"); return NS_ERR_CUDA; }
    if (has_found && Q) std::memcpy(has_found, scratch->has.data(), Q);
    const double resolve_ms = std::chrono::duration<double, std::milli>(clk::now() - t0).count();

    // A few long queries must not put the whole batch on a wide kernel variant (one 40-term query among 4096 short
    // ones costs the batch +22 %, one 250-term query 2.9 x: profiles/r2_long_query_probe.json): they are scored as a
    // batch of their own, the short ones by the NG = 1 kernel, and the answers are put back in the caller's order.
    uint32_t nwide = 0;
    for (uint32_t q = 0; q < Q; q++) nwide += scratch->wide[q];
    if (Q >= 256 && nwide != 0 && (uint64_t)nwide * 4 <= Q && out_hits && out_nhits && out_found) {
        const uint32_t K = (uint32_t)std::max(1, std::min(k, NS_MAX_K));
        std::vector<uint32_t> idx[2];
        for (uint32_t q = 0; q < Q; q++) idx[scratch->wide[q]].push_back(q);
        for (int cls = 0; cls < 2; cls++) {
            std::vector<DevResolved> sub;
            subset_parts(scratch->parts, idx[cls], cls == 0 ? 32u : 0xFFFFFFFFu, sub);
            const uint32_t n = (uint32_t)idx[cls].size();
            std::vector<ns_hit> h((size_t)n * K);
            std::vector<uint32_t> nh(n);
            std::vector<uint64_t> fo(n);
            int rc = run_resolved(e, g, sub, n, k, h.data(), nh.data(), fo.data(), cls == 0 ? resolve_ms : 0.0);
            if (rc != NS_OK) return rc;
            for (uint32_t i = 0; i < n; i++) {
                const uint32_t q = idx[cls][i];
                std::memcpy(out_hits + (size_t)q * K, h.data() + (size_t)i * K, (size_t)K * sizeof(ns_hit));
                out_nhits[q] = nh[i];
                out_found[q] = fo[i];
            }
        }
        return NS_OK;
    }
    return run_resolved(e, g, scratch->parts, Q, k, out_hits, out_nhits, out_found, resolve_ms);
}

// ------------------------------------------------------------------------------------------
// request coalescing: many threads each with ONE query -> GPU-sized batches
// ------------------------------------------------------------------------------------------

struct Coalescer {
    struct Req {
        const char* query;
        int k;
        ns_hit* hits;
        uint32_t* nhits;
        uint64_t* found;
        uint8_t* has;
        std::shared_ptr<const Generation>* gen_out;
        int rc = NS_OK;
        std::string err;
        bool done = false;
        std::chrono::steady_clock::time_point t_in;
    };
    ns_engine* e;
    uint32_t max_batch, max_wait_us;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<Req*> queue;
    bool stop = false;
    std::vector<std::thread> th;
    std::atomic<uint64_t> n_batches{0}, n_queries{0}, max_seen{0};

    Coalescer(ns_engine* eng, uint32_t mb, uint32_t mw, int dispatchers) : e(eng), max_batch(std::max(1u, mb)), max_wait_us(mw) {
        for (int i = 0; i < std::max(1, dispatchers); i++) th.emplace_back([this] { loop(); });
    }
    ~Coalescer() { shutdown(); }
    // stop accepting, serve what is queued, join the dispatchers (idempotent)
    void shutdown() {
        {
            std::lock_guard<std::mutex> lk(mu);
            if (stop) return;
            stop = true;
        }
        cv_work.notify_all();
        for (auto& t : th) t.join();
    }
    int submit_and_wait(Req& r) {
        r.t_in = std::chrono::steady_clock::now();
        std::unique_lock<std::mutex> lk(mu);
        if (stop) return -1;  // stopped between the caller's look-up and now: the caller searches directly
        queue.push_back(&r);
        if (queue.size() == 1 || queue.size() >= max_batch) cv_work.notify_one();
        cv_done.wait(lk, [&] { return r.done; });
        if (r.rc != NS_OK) set_error(r.err);
        return r.rc;
    }
    void loop();
};

void Coalescer::loop() {
    std::vector<Req*> take;
    for (;;) {
        take.clear();
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_work.wait(lk, [&] { return stop || !queue.empty(); });
            if (queue.empty()) {
                if (stop) return;
                continue;
            }
            // gather: until the batch is full or the oldest request has waited max_wait_us
            const auto deadline = queue.front()->t_in + std::chrono::microseconds(max_wait_us);
            while (!stop && !queue.empty() && queue.size() < max_batch && std::chrono::steady_clock::now() < deadline)
                cv_work.wait_until(lk, deadline);
            if (queue.empty()) continue;  // another dispatcher took the requests while this one was gathering
            // one batch = requests of one k class (k <= 16 and k > 16 run different kernel variants)
            const bool big = queue.front()->k > 16;
            for (auto it = queue.begin(); it != queue.end() && take.size() < max_batch;) {
                if (((*it)->k > 16) == big) {
                    take.push_back(*it);
                    it = queue.erase(it);
                } else {
                    ++it;
                }
            }
            if (!queue.empty()) cv_work.notify_one();  // another dispatcher may start on the rest
        }
        const uint32_t Q = (uint32_t)take.size();
        int K = 1;
        for (Req* r : take) K = std::max(K, std::max(1, std::min(r->k, NS_MAX_K)));
        std::vector<ns_hit> hits((size_t)Q * K);
        std::vector<uint32_t> nh(Q);
        std::vector<uint64_t> fo(Q);
        std::vector<uint8_t> has(Q);
        auto gen = e->snapshot();
        int rc = NS_ERR_STATE;
        std::string err = "search before a successful reload";
        if (gen) {
            const Generation& g = *gen;
            rc = search_core(e, gen, Q, [&](uint32_t q, std::vector<QueryTerm>& qt) { return query_terms_of(g, take[q]->query, qt); },
                             K, hits.data(), nh.data(), fo.data(), has.data());
            if (rc != NS_OK) err = ns_last_error();
            if (rc == NS_ERR_INVALID && Q > 1) {
                // One request the device layer refuses (e.g. more than NS_MAX_TERMS terms) must not fail the strangers
                // it happened to share a batch with: every request is served again on its own and gets its own status.
                for (uint32_t q = 0; q < Q; q++) {
                    Req* r = take[q];
                    const int kr = std::max(1, std::min(r->k, NS_MAX_K));
                    uint32_t n1 = 0;
                    uint64_t f1 = 0;
                    uint8_t h1 = 0;
                    r->rc = search_core(e, gen, 1, [&](uint32_t, std::vector<QueryTerm>& qt) { return query_terms_of(g, r->query, qt); },
                                        kr, hits.data(), &n1, &f1, &h1);
                    if (r->rc != NS_OK) {
                        r->err = ns_last_error();
                        continue;
                    }
                    if (r->hits && n1) std::memcpy(r->hits, hits.data(), (size_t)n1 * sizeof(ns_hit));
                    if (r->nhits) *r->nhits = n1;
                    if (r->found) *r->found = f1;
                    if (r->has) *r->has = h1;
                    if (r->gen_out) *r->gen_out = gen;
                }
                n_batches++;
                n_queries += Q;
                {
                    std::lock_guard<std::mutex> lk(mu);
                    for (Req* r : take) r->done = true;
                }
                cv_done.notify_all();
                continue;
            }
        }
        n_batches++;
        n_queries += Q;
        uint64_t seen = max_seen.load();
        while (Q > seen && !max_seen.compare_exchange_weak(seen, Q)) {}
        for (uint32_t q = 0; q < Q; q++) {
            Req* r = take[q];
            r->rc = rc;
            if (rc == NS_OK) {
                // the first k' entries of a top-K list are the top-k' list (total order): cut to the caller's k
                const uint32_t kr = (uint32_t)std::max(1, std::min(r->k, NS_MAX_K));
                const uint32_t n = std::min(nh[q], kr);
                if (r->hits && n) std::memcpy(r->hits, hits.data() + (size_t)q * K, (size_t)n * sizeof(ns_hit));
                if (r->nhits) *r->nhits = n;
                if (r->found) *r->found = fo[q];
                if (r->has) *r->has = has[q];
                if (r->gen_out) *r->gen_out = gen;
            } else {
                r->err = err;
            }
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            for (Req* r : take) r->done = true;
        }
        cv_done.notify_all();
    }
}

// one query, through the coalescer when it runs; gen_out receives the generation the answer came from
int search_one(ns_engine* e, const char* query, int k, ns_hit* hits, uint32_t* nhits, uint64_t* found, uint8_t* has,
               std::shared_ptr<const Generation>* gen_out) {
    std::shared_ptr<Coalescer> co;
    {
        std::lock_guard<std::mutex> lk(e->co_mu);
        co = e->coalescer;
    }
    if (co) {
        Coalescer::Req r;
        r.query = query;
        r.k = k;
        r.hits = hits;
        r.nhits = nhits;
        r.found = found;
        r.has = has;
        r.gen_out = gen_out;
        const int rc = co->submit_and_wait(r);
        if (rc >= 0) return rc;
    }
    auto gen = e->snapshot();
    if (gen_out) *gen_out = gen;
    if (!gen) { set_error("search before a successful reload"); return NS_ERR_STATE; }
    const Generation& g = *gen;
    return search_core(e, gen, 1, [&](uint32_t, std::vector<QueryTerm>& qt) { return query_terms_of(g, query, qt); }, k, hits,
                       nhits, found, has);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------

static int ns_engine_create_multi_impl(const char* index_dir, int ndev, const int* devices, ns_engine** out) {
    if (!index_dir || !out || ndev < 0 || (ndev > 0 && !devices) || ndev > NS_MAX_PEERS) {
        set_error("ns_engine_create_multi: bad argument");
        return NS_ERR_INVALID;
    }
    *out = nullptr;
    auto e = std::make_unique<ns_engine>();
    e->index_dir = index_dir;
    e->keep_raw = std::getenv("NSB200_KEEP_RAW") != nullptr;
    e->trace = std::getenv("NSB200_TRACE") != nullptr;
    for (int d = 0; d < ndev; d++) {
        // (the same ordinal may be listed more than once: two device slots on one GPU — how the
        //  multi-device path is exercised on a single-GPU box)
        ns_index* ix = nullptr;
        if (ns_index_create(devices[d], &ix) != NS_OK) goto fail_cuda;
        e->idx.push_back(ix);
        e->devices.push_back(devices[d]);
    }
    if (ndev > 1)
        for (int d = 0; d < ndev; d++) e->dev_threads.emplace_back(new DeviceWorker());
    *out = e.release();
    return NS_OK;
fail_cuda:
    for (auto* ix : e->idx) ns_index_destroy(ix);
    return NS_ERR_CUDA;
}

extern "C" int ns_engine_create_multi(const char* index_dir, int ndev, const int* devices, ns_engine** out) {
    return abi_guard("ns_engine_create_multi", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_engine_create_multi_impl(index_dir, ndev, devices, out); });
}

extern "C" int ns_engine_create(const char* index_dir, int device, ns_engine** out) {
    if (device < 0) return ns_engine_create_multi(index_dir, 0, nullptr, out);
    return ns_engine_create_multi(index_dir, 1, &device, out);
}

extern "C" void ns_engine_destroy(ns_engine* e) {
    if (!e) return;
    {
        std::shared_ptr<Coalescer> c;
        {
            std::lock_guard<std::mutex> lk(e->co_mu);
            c = std::move(e->coalescer);
        }
        if (c) c->shutdown();  // drains and joins the dispatchers
    }
    e->dev_threads.clear();  // runs what is queued (batch teardowns), then joins
    e->xg_pool.clear();
    {
        std::lock_guard<std::mutex> lk(e->gen_mu);
        e->gen.reset();
    }
    for (auto* ix : e->idx) ns_index_destroy(ix);
    delete e;
}

extern "C" int ns_engine_num_devices(const ns_engine* e) { return e ? (int)e->devices.size() : 0; }

extern "C" int ns_engine_set_shard(ns_engine* e, int rank, int world) {
    if (!e || world < 1 || rank < 0 || rank >= world) { set_error("ns_engine_set_shard: bad rank/world"); return NS_ERR_INVALID; }
    std::lock_guard<std::mutex> lk(e->reload_mu);
    e->rank = rank;
    e->world = world;
    return NS_OK;
}

static int ns_engine_reload_impl(ns_engine* e) {
    if (!e) { set_error("ns_engine_reload: null"); return NS_ERR_INVALID; }
    std::lock_guard<std::mutex> lk(e->reload_mu);
    return do_reload(e);
}

extern "C" int ns_engine_reload(ns_engine* e) {
    return abi_guard("ns_engine_reload", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_engine_reload_impl(e); });
}

extern "C" int ns_engine_reload_stats(const ns_engine* e, double* total_s, double* read_upload_s, double* dict_s,
                                      uint64_t* posting_bytes, uint64_t* device_bytes) {
    if (!e) return NS_ERR_INVALID;
    ReloadStats rs;
    {
        std::lock_guard<std::mutex> lk(const_cast<ns_engine*>(e)->gen_mu);
        rs = e->last_reload;
    }
    if (total_s) *total_s = rs.total_s;
    if (read_upload_s) *read_upload_s = rs.read_s;
    if (dict_s) *dict_s = rs.dict_s;
    if (posting_bytes) *posting_bytes = rs.posting_bytes;
    if (device_bytes) {
        uint64_t b = 0;
        for (auto* ix : e->idx) b += ns_index_device_bytes(ix);
        *device_bytes = b;
    }
    return NS_OK;
}

extern "C" int ns_engine_num_segments(const ns_engine* e) {
    if (!e) return 0;
    auto g = const_cast<ns_engine*>(e)->snapshot();
    return g ? (int)g->seg_names.size() : 0;
}

static int ns_engine_segment_name_impl(const ns_engine* e, int i, char* buf, size_t cap) {
    if (!e || !buf) return -1;
    auto g = const_cast<ns_engine*>(e)->snapshot();
    if (!g || i < 0 || (size_t)i >= g->seg_names.size()) return -1;
    const std::string& s = g->seg_names[(size_t)i];
    if (s.size() + 1 > cap) return -1;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

extern "C" int ns_engine_segment_name(const ns_engine* e, int i, char* buf, size_t cap) {
    return abi_guard("ns_engine_segment_name", -1, -1, [&]() -> int { return ns_engine_segment_name_impl(e, i, buf, cap); });
}

static int ns_engine_segment_stats_impl(const ns_engine* e, int i, uint32_t* N, float* avgdl, uint32_t* T, uint64_t* P) {
    if (!e) return NS_ERR_INVALID;
    auto g = const_cast<ns_engine*>(e)->snapshot();
    if (!g || i < 0 || (size_t)i >= g->segs.size() || !g->segs[(size_t)i]) { set_error("segment not loaded by this engine"); return NS_ERR_INVALID; }
    const HostSegment& s = *g->segs[(size_t)i];
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

extern "C" int ns_engine_segment_stats(const ns_engine* e, int i, uint32_t* N, float* avgdl, uint32_t* T, uint64_t* P) {
    return abi_guard("ns_engine_segment_stats", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_engine_segment_stats_impl(e, i, N, avgdl, T, P); });
}

static int ns_engine_term_stats_impl(const ns_engine* e, int i, const char* term, uint32_t* df, uint32_t* count) {
    if (df) *df = 0;
    if (count) *count = 0;
    if (!e || !term) return NS_ERR_INVALID;
    auto g = const_cast<ns_engine*>(e)->snapshot();
    if (!g || i < 0 || (size_t)i >= g->segs.size() || !g->segs[(size_t)i]) { set_error("segment not loaded by this engine"); return NS_ERR_INVALID; }
    auto it = g->segs[(size_t)i]->lex.find(term);
    if (it == g->segs[(size_t)i]->lex.end()) return NS_OK;
    if (df) *df = g->segs[(size_t)i]->rows[it->second].df;
    if (count) *count = g->segs[(size_t)i]->rows[it->second].count;
    return NS_OK;
}

extern "C" int ns_engine_term_stats(const ns_engine* e, int i, const char* term, uint32_t* df, uint32_t* count) {
    return abi_guard("ns_engine_term_stats", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_engine_term_stats_impl(e, i, term, df, count); });
}

extern "C" ns_index* ns_engine_index(ns_engine* e) { return e && !e->idx.empty() ? e->idx[0] : nullptr; }
extern "C" ns_index* ns_engine_device_index(ns_engine* e, int slot) {
    return e && slot >= 0 && (size_t)slot < e->idx.size() ? e->idx[(size_t)slot] : nullptr;
}

static int ns_engine_cord_uid_impl(const ns_engine* e, uint32_t seg, uint32_t doc, char* buf, size_t cap) {
    if (!e || !buf) return -1;
    auto g = const_cast<ns_engine*>(e)->snapshot();
    if (!g || seg >= g->segs.size() || !g->segs[seg]) return -1;
    std::string s = g->segs[seg]->cord_uid(doc);
    if (s.size() + 1 > cap) return -1;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

extern "C" int ns_engine_cord_uid(const ns_engine* e, uint32_t seg, uint32_t doc, char* buf, size_t cap) {
    return abi_guard("ns_engine_cord_uid", -1, -1, [&]() -> int { return ns_engine_cord_uid_impl(e, seg, doc, buf, cap); });
}

namespace {

template <class GetQuery>
int resolve_abi(ns_engine* e, uint32_t Q, GetQuery get, uint64_t* q_off, ns_qterm* terms, uint64_t terms_cap,
                uint64_t* n_terms, uint8_t* has_terms) {
    auto gen = e->snapshot();
    if (!gen) { set_error("resolve before a successful reload"); return NS_ERR_STATE; }
    const Generation& g = *gen;
    std::vector<uint32_t> all(g.dict.owned.size());
    for (size_t j = 0; j < all.size(); j++) all[j] = (uint32_t)j;
    std::vector<const std::vector<uint32_t>*> colsets{&all};
    std::vector<Resolved> parts;
    std::vector<uint8_t> has;
    resolve_all(e, g, Q, [&](uint32_t q, std::vector<QueryTerm>& qt) { return query_terms_of(g, get(q), qt); }, colsets, parts, has);
    const Resolved& r = parts[0];
    const uint64_t total = r.q_off[Q];
    std::memcpy(q_off, r.q_off.data(), ((size_t)Q + 1) * sizeof(uint64_t));
    *n_terms = total;
    if (has_terms && Q) std::memcpy(has_terms, has.data(), Q);
    if (!terms) return NS_OK;
    if (terms_cap < total) { set_error("ns_engine_resolve_batch: terms buffer too small"); return NS_ERR_INVALID; }
    if (total) std::memcpy(terms, r.terms.data(), total * sizeof(ns_qterm));
    return NS_OK;
}

}  // namespace

static int ns_engine_resolve_batch_impl(ns_engine* e, uint32_t Q, const char* const* queries, uint64_t* q_off,
                                       ns_qterm* terms, uint64_t terms_cap, uint64_t* n_terms, uint8_t* has_terms) {
    if (!e || (Q && !queries) || !q_off || !n_terms) { set_error("ns_engine_resolve_batch: null argument"); return NS_ERR_INVALID; }
    return resolve_abi(e, Q, [&](uint32_t q) { return queries[q]; }, q_off, terms, terms_cap, n_terms, has_terms);
}

extern "C" int ns_engine_resolve_batch(ns_engine* e, uint32_t Q, const char* const* queries, uint64_t* q_off,
                                       ns_qterm* terms, uint64_t terms_cap, uint64_t* n_terms, uint8_t* has_terms) {
    return abi_guard("ns_engine_resolve_batch", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_engine_resolve_batch_impl(e, Q, queries, q_off, terms, terms_cap, n_terms, has_terms); });
}

static int ns_engine_resolve_batch_packed_impl(ns_engine* e, uint32_t Q, const char* zqueries, size_t nbytes,
                                              uint64_t* q_off, ns_qterm* terms, uint64_t terms_cap, uint64_t* n_terms,
                                              uint8_t* has_terms) {
    if (!e || (Q && !zqueries) || !q_off || !n_terms) { set_error("ns_engine_resolve_batch_packed: null argument"); return NS_ERR_INVALID; }
    std::vector<const char*> starts;
    if (!split_packed(zqueries, nbytes, Q, starts)) {
        set_error("ns_engine_resolve_batch_packed: buffer holds fewer than Q NUL-terminated strings");
        return NS_ERR_INVALID;
    }
    return resolve_abi(e, Q, [&](uint32_t q) { return starts[q]; }, q_off, terms, terms_cap, n_terms, has_terms);
}

extern "C" int ns_engine_resolve_batch_packed(ns_engine* e, uint32_t Q, const char* zqueries, size_t nbytes,
                                              uint64_t* q_off, ns_qterm* terms, uint64_t terms_cap, uint64_t* n_terms,
                                              uint8_t* has_terms) {
    return abi_guard("ns_engine_resolve_batch_packed", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_engine_resolve_batch_packed_impl(e, Q, zqueries, nbytes, q_off, terms, terms_cap, n_terms, has_terms); });
}

static int ns_engine_search_batch_impl(ns_engine* e, uint32_t Q, const char* const* queries, int k, ns_hit* out_hits,
                                      uint32_t* out_nhits, uint64_t* out_found, uint8_t* has_found) {
    if (!e || (Q && !queries)) { set_error("ns_engine_search_batch: null argument"); return NS_ERR_INVALID; }
    auto gen = e->snapshot();
    if (e->idx.empty()) { set_error("engine was created without a CUDA device; there is no CPU search path"); return NS_ERR_STATE; }
    if (!gen) { set_error("search before a successful reload"); return NS_ERR_STATE; }
    const Generation& g = *gen;
    return search_core(e, gen, Q, [&](uint32_t q, std::vector<QueryTerm>& qt) { return query_terms_of(g, queries[q], qt); }, k,
                       out_hits, out_nhits, out_found, has_found);
}

extern "C" int ns_engine_search_batch(ns_engine* e, uint32_t Q, const char* const* queries, int k, ns_hit* out_hits,
                                      uint32_t* out_nhits, uint64_t* out_found, uint8_t* has_found) {
    return abi_guard("ns_engine_search_batch", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_engine_search_batch_impl(e, Q, queries, k, out_hits, out_nhits, out_found, has_found); });
}

static int ns_engine_search_batch_packed_impl(ns_engine* e, uint32_t Q, const char* zqueries, size_t nbytes, int k,
                                             ns_hit* out_hits, uint32_t* out_nhits, uint64_t* out_found,
                                             uint8_t* has_found) {
    if (!e || (Q && !zqueries)) { set_error("ns_engine_search_batch_packed: null argument"); return NS_ERR_INVALID; }
    if (e->idx.empty()) { set_error("engine was created without a CUDA device; there is no CPU search path"); return NS_ERR_STATE; }
    std::vector<const char*> starts;
    if (!split_packed(zqueries, nbytes, Q, starts)) {
        set_error("ns_engine_search_batch_packed: buffer holds fewer than Q NUL-terminated strings");
        return NS_ERR_INVALID;
    }
    auto gen = e->snapshot();
    if (!gen) { set_error("search before a successful reload"); return NS_ERR_STATE; }
    const Generation& g = *gen;
    return search_core(e, gen, Q, [&](uint32_t q, std::vector<QueryTerm>& qt) { return query_terms_of(g, starts[q], qt); }, k,
                       out_hits, out_nhits, out_found, has_found);
}

extern "C" int ns_engine_search_batch_packed(ns_engine* e, uint32_t Q, const char* zqueries, size_t nbytes, int k,
                                             ns_hit* out_hits, uint32_t* out_nhits, uint64_t* out_found,
                                             uint8_t* has_found) {
    return abi_guard("ns_engine_search_batch_packed", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_engine_search_batch_packed_impl(e, Q, zqueries, nbytes, k, out_hits, out_nhits, out_found, has_found); });
}

// Front end + prepare of a batch on a single-device engine, WITHOUT launching: tokenise, dictionary, kernel-form
// records, pinned staging, H2D.  What a caller that drives the launch itself needs (nextsearch-api_b200/dist.py: one
// process per GPU, launches ordered across ranks) — the same work ns_engine_search_batch_packed does before its launch.
static int ns_engine_prepare_batch_packed_impl(ns_engine* e, uint32_t Q, const char* zqueries, size_t nbytes, int k,
                                              ns_batch** out, uint8_t* has_found) {
    if (!e || !out || (Q && !zqueries)) { set_error("ns_engine_prepare_batch_packed: null argument"); return NS_ERR_INVALID; }
    *out = nullptr;
    if (e->idx.size() != 1) { set_error("ns_engine_prepare_batch_packed: needs an engine with exactly one device"); return NS_ERR_STATE; }
    std::vector<const char*> starts;
    if (!split_packed(zqueries, nbytes, Q, starts)) {
        set_error("ns_engine_prepare_batch_packed: buffer holds fewer than Q NUL-terminated strings");
        return NS_ERR_INVALID;
    }
    auto gen = e->snapshot();
    if (!gen) { set_error("search before a successful reload"); return NS_ERR_STATE; }
    const Generation& g = *gen;
    ResolveScratch sc;
    resolve_devices(e, g, Q, [&](uint32_t q, std::vector<QueryTerm>& qt) { return query_terms_of(g, starts[q], qt); }, sc);
    if (sc.failed == 2) { set_error("front end: out of host memory"); return NS_ERR_NOMEM; }
    if (sc.failed) { set_error("semantic expansion: the device similarity scan failed"); return NS_ERR_CUDA; }
    if (has_found && Q) std::memcpy(has_found, sc.has.data(), Q);
    const DevResolved& r = sc.parts[0];
    PreparedBatch pb{r.qoff.data(), r.terms.data(), r.weight.data(), r.max_in_seg, r.unit, r.scan_always};
    int rc = batch_prepare_trusted(e->idx[0], g.dev_state[0], Q, k, pb, out);
    if (rc != NS_ERR_STATE) return rc;
    // no resident scores on the device (NSB200_NO_RESIDENT): the generic, validating prepare
    std::vector<uint64_t> q_off((size_t)Q + 1);
    std::vector<ns_qterm> qt(std::max<size_t>(1, r.qoff[Q]));
    for (uint32_t q = 0; q <= Q; q++) q_off[q] = r.qoff[q];
    for (uint32_t i = 0; i < r.qoff[Q]; i++)
        qt[i] = ns_qterm{g.dict.owned[g.dev_cols[0][r.terms[i].slot]], r.terms[i].row, r.terms[i].idf, r.terms[i].w};
    return batch_prepare_on(e->idx[0], g.dev_state[0], Q, k, q_off.data(), qt.data(), out);
}

extern "C" int ns_engine_prepare_batch_packed(ns_engine* e, uint32_t Q, const char* zqueries, size_t nbytes, int k,
                                              ns_batch** out, uint8_t* has_found) {
    return abi_guard("ns_engine_prepare_batch_packed", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_engine_prepare_batch_packed_impl(e, Q, zqueries, nbytes, k, out, has_found); });
}

// Explicit qterms_w lists (the reference's vector<pair<string, float>> of src/api_engine.cpp:410-421), one per
// query: t_off[Q+1] indexes terms[] / weights[].  No tokenisation, no filter, no expansion: exactly what the
// scoring loop (:426-505) receives.
static int ns_engine_search_terms_batch_impl(ns_engine* e, uint32_t Q, const uint64_t* t_off, const char* const* terms,
                                            const float* weights, int k, ns_hit* out_hits, uint32_t* out_nhits,
                                            uint64_t* out_found, uint8_t* has_found) {
    if (!e || !t_off || (Q && t_off[Q] && (!terms || !weights))) { set_error("ns_engine_search_terms_batch: null argument"); return NS_ERR_INVALID; }
    if (e->idx.empty()) { set_error("engine was created without a CUDA device; there is no CPU search path"); return NS_ERR_STATE; }
    auto gen = e->snapshot();
    if (!gen) { set_error("search before a successful reload"); return NS_ERR_STATE; }
    const Generation& g = *gen;
    auto terms_of = [&](uint32_t q, std::vector<QueryTerm>& qt) {
        qt.clear();
        for (uint64_t i = t_off[q]; i < t_off[q + 1]; i++) {
            const size_t n = std::strlen(terms[i]);
            const int64_t gid = g.dict.find(terms[i], n, term_hash(terms[i], n));
            if (gid >= 0) qt.push_back(QueryTerm{(uint32_t)gid, weights[i]});
        }
        return t_off[q + 1] > t_off[q] && !g.seg_names.empty();  // :407, :424
    };
    return search_core(e, gen, Q, terms_of, k, out_hits, out_nhits, out_found, has_found);
}

extern "C" int ns_engine_search_terms_batch(ns_engine* e, uint32_t Q, const uint64_t* t_off, const char* const* terms,
                                            const float* weights, int k, ns_hit* out_hits, uint32_t* out_nhits,
                                            uint64_t* out_found, uint8_t* has_found) {
    return abi_guard("ns_engine_search_terms_batch", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_engine_search_terms_batch_impl(e, Q, t_off, terms, weights, k, out_hits, out_nhits, out_found, has_found); });
}

// SemanticIndex::expand for one query (src/api_engine.cpp:410-417): the kept tokens are expanded with the
// loaded embeddings.  Terms come back NUL-separated in `buf`, weights in `weights`; returns the count, -1
// if a buffer is too small, 0 with *enabled = 0 when no embeddings are loaded.
static int ns_engine_expand_impl(ns_engine* e, const char* query, char* buf, size_t cap, float* weights, int wcap, int* enabled) {
    if (enabled) *enabled = 0;
    if (!e || !query || !buf) return -1;
    auto gen = e->snapshot();
    if (!gen || !gen->sem.enabled) return 0;
    if (enabled) *enabled = 1;
    std::vector<std::string> base;
    query_terms(query, base);
    if (base.empty()) return 0;
    const auto qw = gen->sem.expand(base);
    size_t at = 0;
    int n = 0;
    for (auto& tw : qw) {
        if (at + tw.first.size() + 1 > cap || n >= wcap) return -1;
        std::memcpy(buf + at, tw.first.c_str(), tw.first.size() + 1);
        at += tw.first.size() + 1;
        if (weights) weights[n] = tw.second;
        n++;
    }
    return n;
}

extern "C" int ns_engine_expand(ns_engine* e, const char* query, char* buf, size_t cap, float* weights, int wcap, int* enabled) {
    return abi_guard("ns_engine_expand", -1, -1, [&]() -> int { return ns_engine_expand_impl(e, query, buf, cap, weights, wcap, enabled); });
}

static int ns_engine_search_one_impl(ns_engine* e, const char* query, int k, ns_hit* out_hits, uint32_t* out_nhits,
                                    uint64_t* out_found, uint8_t* has_found) {
    if (!e || !query) { set_error("ns_engine_search_one: null argument"); return NS_ERR_INVALID; }
    if (e->idx.empty()) { set_error("engine was created without a CUDA device; there is no CPU search path"); return NS_ERR_STATE; }
    return search_one(e, query, k, out_hits, out_nhits, out_found, has_found, nullptr);
}

extern "C" int ns_engine_search_one(ns_engine* e, const char* query, int k, ns_hit* out_hits, uint32_t* out_nhits,
                                    uint64_t* out_found, uint8_t* has_found) {
    return abi_guard("ns_engine_search_one", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_engine_search_one_impl(e, query, k, out_hits, out_nhits, out_found, has_found); });
}

static int ns_engine_coalescer_start_impl(ns_engine* e, uint32_t max_batch, uint32_t max_wait_us, int dispatchers) {
    if (!e) return NS_ERR_INVALID;
    if (e->idx.empty()) { set_error("engine was created without a CUDA device; there is no CPU search path"); return NS_ERR_STATE; }
    std::lock_guard<std::mutex> lk(e->co_mu);
    if (e->coalescer) { set_error("coalescer already running"); return NS_ERR_STATE; }
    e->coalescer.reset(new Coalescer(e, max_batch ? max_batch : 4096, max_wait_us, dispatchers > 0 ? dispatchers : 2));
    return NS_OK;
}

extern "C" int ns_engine_coalescer_start(ns_engine* e, uint32_t max_batch, uint32_t max_wait_us, int dispatchers) {
    return abi_guard("ns_engine_coalescer_start", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_engine_coalescer_start_impl(e, max_batch, max_wait_us, dispatchers); });
}

extern "C" int ns_engine_coalescer_stop(ns_engine* e) {
    if (!e) return NS_ERR_INVALID;
    std::shared_ptr<Coalescer> c;
    {
        std::lock_guard<std::mutex> lk(e->co_mu);
        c = std::move(e->coalescer);
    }
    if (c) c->shutdown();  // serves what is queued, then joins; late submitters are refused, not left hanging
    return NS_OK;
}

extern "C" int ns_engine_coalescer_stats(ns_engine* e, uint64_t* batches, uint64_t* queries, uint64_t* max_batch_seen) {
    if (!e) return NS_ERR_INVALID;
    std::lock_guard<std::mutex> lk(e->co_mu);
    if (batches) *batches = e->coalescer ? e->coalescer->n_batches.load() : 0;
    if (queries) *queries = e->coalescer ? e->coalescer->n_queries.load() : 0;
    if (max_batch_seen) *max_batch_seen = e->coalescer ? e->coalescer->max_seen.load() : 0;
    return NS_OK;
}

// Load generator (bench tooling): `nthreads` host threads each issue `per_thread` blocking ns_engine_search_one
// calls — the traffic shape of the reference's HTTP workers (src/api_server.cpp:117-178) without Python in the
// loop.  Thread t starts at query (t * per_thread) % Q and walks the Q given queries cyclically.
static int ns_engine_load_test_impl(ns_engine* e, uint32_t nthreads, uint32_t per_thread, uint32_t Q, const char* zqueries,
                                   size_t nbytes, int k, double* qps, double* p50_us, double* p99_us) {
    if (!e || !zqueries || Q == 0 || nthreads == 0 || per_thread == 0) { set_error("ns_engine_load_test: bad argument"); return NS_ERR_INVALID; }
    std::vector<const char*> starts;
    if (!split_packed(zqueries, nbytes, Q, starts)) { set_error("ns_engine_load_test: fewer than Q strings"); return NS_ERR_INVALID; }
    const int K = std::max(1, std::min(k, NS_MAX_K));
    std::vector<std::vector<double>> lat(nthreads);
    std::atomic<int> failed{NS_OK};
    std::vector<std::string> errs(nthreads);
    std::vector<std::thread> th;
    const auto t0 = std::chrono::steady_clock::now();
    for (uint32_t t = 0; t < nthreads; t++) {
        th.emplace_back([&, t] {
            std::vector<ns_hit> hits((size_t)K);
            uint32_t nh;
            uint64_t fo;
            uint8_t has;
            lat[t].reserve(per_thread);
            for (uint32_t i = 0; i < per_thread; i++) {
                const char* q = starts[((uint64_t)t * per_thread + i) % Q];
                const auto a = std::chrono::steady_clock::now();
                const int rc = ns_engine_search_one(e, q, K, hits.data(), &nh, &fo, &has);
                if (rc != NS_OK) {
                    failed = rc;
                    errs[t] = ns_last_error();
                    return;
                }
                lat[t].push_back(std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count());
            }
        });
    }
    for (auto& x : th) x.join();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (failed != NS_OK) {
        for (auto& m : errs)
            if (!m.empty()) { set_error(m); break; }
        return failed;
    }
    std::vector<double> all;
    for (auto& v : lat) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end());
    if (qps) *qps = secs > 0 ? (double)all.size() / secs : 0.0;
    if (p50_us) *p50_us = all.empty() ? 0.0 : all[all.size() / 2];
    if (p99_us) *p99_us = all.empty() ? 0.0 : all[(size_t)((double)(all.size() - 1) * 0.99)];
    return NS_OK;
}

extern "C" int ns_engine_load_test(ns_engine* e, uint32_t nthreads, uint32_t per_thread, uint32_t Q, const char* zqueries,
                                   size_t nbytes, int k, double* qps, double* p50_us, double* p99_us) {
    return abi_guard("ns_engine_load_test", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_engine_load_test_impl(e, nthreads, per_thread, Q, zqueries, nbytes, k, qps, p50_us, p99_us); });
}

// Engine::search (src/api_engine.cpp:369-542) as JSON text, byte-compatible with nlohmann's dump() of the
// reference's object: keys in lexicographic order, floats via Grisu2 (json_text.hpp).
static int ns_engine_search_json_impl(ns_engine* e, const char* query, int k, char* buf, size_t cap, size_t* needed) {
    if (!e || !query) { set_error("ns_engine_search_json: null argument"); return NS_ERR_INVALID; }
    if (e->idx.empty()) { set_error("engine was created without a CUDA device; there is no CPU search path"); return NS_ERR_STATE; }
    const int K = std::max(1, std::min(k, NS_MAX_K));  // src/api_engine.cpp:377
    std::vector<ns_hit> hits((size_t)K);
    uint32_t nh = 0;
    uint64_t found = 0;
    uint8_t has = 0;
    std::shared_ptr<const Generation> gen;
    int rc = search_one(e, query, K, hits.data(), &nh, &found, &has, &gen);
    if (rc != NS_OK) return rc;
    const Generation& g = *gen;  // names, uids and metadata of the generation that produced the hits
    std::string out = "{";
    bool ok = true;
    if (has) out += "\"found\":" + std::to_string(found) + ",";
    out += "\"k\":" + std::to_string(K) + ",\"query\":";
    ok = ok && jsontext::append_string(out, query, std::strlen(query));
    out += ",\"results\":[";
    MetaFields mf;
    for (uint32_t i = 0; i < nh && ok; i++) {
        const ns_hit& h = hits[i];
        if (i) out += ",";
        const std::string uid = (h.seg < g.segs.size() && g.segs[h.seg]) ? g.segs[h.seg]->cord_uid(h.doc) : std::string();
        const bool decorated = g.meta.enabled() && g.meta.lookup(uid, mf);  // :516-532
        out += "{";
        if (decorated && !mf.author.empty()) {
            out += "\"author\":";
            ok = ok && jsontext::append_string(out, mf.author);
            out += ",";
        }
        out += "\"cord_uid\":";
        ok = ok && jsontext::append_string(out, uid);
        out += ",\"docId\":" + std::to_string(h.doc);
        if (decorated && !mf.publish_time.empty()) {
            out += ",\"publish_time\":";
            ok = ok && jsontext::append_string(out, mf.publish_time);
        }
        out += ",\"score\":";
        jsontext::append_double(out, (double)h.score);  // r["score"] = h.s widens the f32 (:511)
        out += ",\"segment\":";
        ok = ok && jsontext::append_string(out, h.seg < g.seg_names.size() ? g.seg_names[h.seg] : std::string());
        if (decorated) {
            if (!mf.title.empty()) {
                out += ",\"title\":";
                ok = ok && jsontext::append_string(out, mf.title);
            }
            std::string url = mf.url;
            const size_t semi = url.find(';');
            if (semi != std::string::npos) url = url.substr(0, semi);  // :524-526
            if (!url.empty()) {
                out += ",\"url\":";
                ok = ok && jsontext::append_string(out, url);
            }
        }
        out += "}";
    }
    out += "],\"segments\":" + std::to_string(g.seg_names.size()) + "}";
    if (!ok) {
        // nlohmann::json::dump throws type_error.316 on invalid UTF-8 and the reference's HTTP handler answers 500
        set_error("query or result field is not valid UTF-8 (the reference's json::dump throws type_error.316)");
        return NS_ERR_INVALID;
    }
    if (needed) *needed = out.size();
    if (buf && cap) {
        size_t n = std::min(cap - 1, out.size());
        std::memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return NS_OK;
}

extern "C" int ns_engine_search_json(ns_engine* e, const char* query, int k, char* buf, size_t cap, size_t* needed) {
    return abi_guard("ns_engine_search_json", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_engine_search_json_impl(e, query, k, buf, cap, needed); });
}

// ------------------------------------------------------------------------------------------

static int ns_text_query_terms_impl(const char* query, char* buf, size_t cap) {
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

extern "C" int ns_text_query_terms(const char* query, char* buf, size_t cap) {
    return abi_guard("ns_text_query_terms", -1, -1, [&]() -> int { return ns_text_query_terms_impl(query, buf, cap); });
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

static int ns_corpus_write_segment_impl(const ns_corpus_spec* spec, uint64_t doc_base, uint32_t ndocs, const char* segdir,
                                       int write_forward, const char* dump_path, int nthreads) {
    if (!spec || !segdir) { set_error("ns_corpus_write_segment: null argument"); return NS_ERR_INVALID; }
    GenSegment g;
    generate_segment(to_spec(spec), doc_base, ndocs, nthreads > 0 ? nthreads : hw_threads(), g);
    if (!write_segment_files(g, segdir, write_forward != 0)) return NS_ERR_IO;
    if (dump_path && !write_corpus_dump(g, dump_path)) { set_error("cannot write corpus dump"); return NS_ERR_IO; }
    return NS_OK;
}

extern "C" int ns_corpus_write_segment(const ns_corpus_spec* spec, uint64_t doc_base, uint32_t ndocs, const char* segdir,
                                       int write_forward, const char* dump_path, int nthreads) {
    return abi_guard("ns_corpus_write_segment", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_corpus_write_segment_impl(spec, doc_base, ndocs, segdir, write_forward, dump_path, nthreads); });
}

static int ns_corpus_write_manifest_impl(const char* index_dir, uint32_t nseg, const char* const* names) {
    if (!index_dir || (nseg && !names)) return NS_ERR_INVALID;
    if (!make_dirs(index_dir)) { set_error("cannot create index dir"); return NS_ERR_IO; }
    std::vector<std::string> v;
    for (uint32_t i = 0; i < nseg; i++) v.push_back(names[i]);
    return save_manifest(std::string(index_dir) + "/manifest.bin", v) ? NS_OK : NS_ERR_IO;
}

extern "C" int ns_corpus_write_manifest(const char* index_dir, uint32_t nseg, const char* const* names) {
    return abi_guard("ns_corpus_write_manifest", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_corpus_write_manifest_impl(index_dir, nseg, names); });
}

static int ns_corpus_make_queries_impl(const ns_corpus_spec* spec, uint64_t query_seed, uint32_t nq, uint32_t min_terms,
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

extern "C" int ns_corpus_make_queries(const ns_corpus_spec* spec, uint64_t query_seed, uint32_t nq, uint32_t min_terms,
                                      uint32_t max_terms, uint32_t head_ranks, char* buf, size_t cap, size_t* needed) {
    return abi_guard("ns_corpus_make_queries", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_corpus_make_queries_impl(spec, query_seed, nq, min_terms, max_terms, head_ranks, buf, cap, needed); });
}
