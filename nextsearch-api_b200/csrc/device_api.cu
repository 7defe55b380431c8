// C-ABI implementation of the device path: ns_index_*, ns_batch_*, ns_search_batch, ns_merge_device.
// Host-side glue only; all arithmetic on postings happens in bm25_kernels.cuh.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/nextsearch_b200.h"
#include "bm25_kernels.cuh"
#include "semantic_kernels.cuh"
#include "device_internal.hpp"
#include "host/common.hpp"
#include "host/segment_io.hpp"

using namespace nsb;

namespace {

#define NS_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            set_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr);    \
            return NS_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

constexpr size_t kAlign = 256;
inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

// BM25 constants of src/api_engine.cpp:375-376, evaluated in f32 like the reference does
const float kK1 = 1.2f;
const float kB = 0.75f;

// Environment overrides, read ONCE when the index handle is created (never on the query path).
struct Tunables {
    bool no_pack = false, no_resident = false, no_fast = false, explicit_items = false, trace = false;
    long item_postings = 0;   // NSB200_ITEM_POSTINGS (0 = default)
    int window_tiles = -1;    // NSB200_WINDOW_TILES (-1 = from the batch shape)
    int l2_prefetch = -1;     // NSB200_L2_PREFETCH (-1 = auto)
    int impact = -1;          // NSB200_IMPACT (-1 = auto)
    static Tunables from_env() {
        Tunables t;
        t.no_pack = std::getenv("NSB200_NO_PACK") != nullptr;
        t.no_resident = std::getenv("NSB200_NO_RESIDENT") != nullptr;
        t.no_fast = std::getenv("NSB200_NO_FAST") != nullptr;
        t.explicit_items = std::getenv("NSB200_EXPLICIT_ITEMS") != nullptr;
        t.trace = std::getenv("NSB200_TRACE") != nullptr;
        if (const char* s = std::getenv("NSB200_ITEM_POSTINGS")) t.item_postings = std::max(0L, std::atol(s));
        if (const char* s = std::getenv("NSB200_WINDOW_TILES")) t.window_tiles = std::max(0, std::atoi(s));
        if (const char* s = std::getenv("NSB200_L2_PREFETCH")) t.l2_prefetch = std::atoi(s) != 0 ? 1 : 0;
        if (const char* s = std::getenv("NSB200_IMPACT")) t.impact = std::atoi(s) != 0 ? 1 : 0;
        return t;
    }
};

struct SegState {
    uint32_t gseg = 0, ndocs = 0, T = 0, ntiles = 0;
    uint64_t P = 0;
    float avgdl = 0.f;
    uint2* d_post = nullptr;   // raw postings; nullptr when the segment keeps only its resident impacts (NS_SEG_DROP_RAW)
    float* d_norm = nullptr;   // per doc (unpacked segments)
    float* d_lut = nullptr;    // per distinct doc length (packed segments)
    bool packed = false;
    uint32_t* d_tileoff = nullptr;
    uint2* d_imp = nullptr;    // resident impacts, same index space as d_post (nullptr: not built)
    std::vector<uint32_t> h_idf_bits;  // per row: bits of bm25_idf(N, count) the resident impacts were built with
    std::vector<uint32_t> h_count;  // LexEntry.count per row (query weights, row validation)
    std::vector<uint32_t> h_begin;  // first posting per row (impact pre-pass source ranges)
    uint64_t bytes = 0;
    bool norm_in_range = true;      // every norm[] value validated for div_rn_inrange
    void release() {
        if (d_post) cudaFree(d_post);
        if (d_norm) cudaFree(d_norm);
        if (d_lut) cudaFree(d_lut);
        if (d_tileoff) cudaFree(d_tileoff);
        if (d_imp) cudaFree(d_imp);
        d_imp = nullptr;
        d_post = nullptr;
        d_norm = nullptr;
        d_lut = nullptr;
        d_tileoff = nullptr;
    }
};

struct IndexState {
    int device = 0;
    uint32_t tile_docs = kTileDocs;
    std::vector<SegState> segs;  // ascending gseg
    std::unordered_map<uint32_t, uint32_t> slot_of;
    DevSeg* d_segs = nullptr;
    uint32_t* d_tile_base = nullptr;
    uint32_t total_tiles = 0;
    uint64_t bytes = 0;
    ~IndexState() {
        cudaSetDevice(device);
        for (auto& s : segs) s.release();
        if (d_segs) cudaFree(d_segs);
        if (d_tile_base) cudaFree(d_tile_base);
    }
};

// Reusable per-call resources: one device blob, two pinned blobs, a stream and events.
struct BatchRes {
    int device = 0;
    uint8_t* d_blob = nullptr;
    size_t d_cap = 0;
    uint8_t* h_in = nullptr;
    size_t h_in_cap = 0;
    uint8_t* h_out = nullptr;
    size_t h_out_cap = 0;
    uint8_t* d_scratch = nullptr;  // per-batch impact array, grown on demand
    size_t scratch_cap = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_h2d = nullptr;  // the descriptor upload of the batch using this resource set has finished
    cudaEvent_t ev_copy = nullptr; // result copy finished (host waits on it, sleeping)
    ~BatchRes() {
        cudaSetDevice(device);
        if (d_blob) cudaFree(d_blob);
        if (d_scratch) cudaFree(d_scratch);
        if (h_in) cudaFreeHost(h_in);
        if (h_out) cudaFreeHost(h_out);
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        if (ev_h2d) cudaEventDestroy(ev_h2d);
        if (ev_copy) cudaEventDestroy(ev_copy);
        if (stream) cudaStreamDestroy(stream);
    }
};

}  // namespace

// What a batch needs from the index handle that created it.  Shared (shared_ptr) so that destroying a batch
// AFTER its index handle — e.g. garbage-collected wrappers going away in arbitrary order — is harmless.
struct IndexShared {
    int device = 0;
    int sm_count = 148;
    Tunables tun;
    std::unordered_map<const void*, int> occupancy;  // kernel variant -> resident CTAs per SM (filled at create)
    std::mutex mu;                                    // guards pool / closed
    std::vector<std::unique_ptr<BatchRes>> pool;
    bool closed = false;
};

struct ns_index {
    int device = 0;
    uint32_t tile_docs = kTileDocs;
    std::mutex mu;
    std::shared_ptr<IndexState> live;
    std::vector<SegState> staged;
    std::shared_ptr<IndexShared> sh;
};

struct ns_batch {
    std::shared_ptr<IndexShared> owner;
    std::shared_ptr<IndexState> st;
    std::unique_ptr<BatchRes> res;
    uint32_t Q = 0, k = 0;
    uint32_t nitems = 0, max_split = 1;
    bool window_major = false;  // items ordered by doc window first (build_items)
    bool implicit_items = false;  // window-major with one nsplit for all queries: the item list is order[] + arithmetic
    uint32_t nsplit_all = 1;
    bool scan_always = false;
    bool fast = false;  // operand ranges validated + unit weights: FAST kernel variant
    bool impact = false;           // per-batch shared term scores (impact pre-pass)
    uint32_t max_in_seg = 0;       // most terms any (query, segment) has
    uint32_t ndist = 0;            // distinct (segment, row, idf) terms of the batch
    uint64_t dist_postings = 0;    // postings the pre-pass evaluates
    DevDistinct* d_dist = nullptr;
    uint32_t* d_dstart = nullptr;
    uint64_t nterms = 0, postings = 0;
    std::vector<uint64_t> weight;  // postings per query (host copy, for re-splitting)
    bool needs_raw = false;        // some term is scored from raw postings (per-batch pre-pass or non-impact kernel)
    // Device blob of one batch:
    //   [0, up_bytes)                      uploaded per batch: qoff | terms | order (or the explicit item list) | dist | dstart
    //   [off_zero, off_zero + zero_bytes)  zeroed per launch: queue head | per-query locks | per-query done counts |
    //                                      published count | result blob (hits | nhits | found)
    uint32_t* d_qoff = nullptr;
    DevTerm* d_terms = nullptr;
    DevItem* d_items = nullptr;   // explicit items, or order[] when implicit_items
    uint32_t* d_counter = nullptr;
    uint32_t* d_qlock = nullptr;
    unsigned long long* d_qdone = nullptr;
    uint32_t* d_npub = nullptr;
    size_t off_items = 0, up_bytes = 0, items_cap = 0, off_zero = 0, zero_bytes = 0;
    uint8_t* d_out = nullptr;  // hits | nhits | found, contiguous == h_out layout
    ns_hit* d_out_hits = nullptr;
    uint32_t* d_out_n = nullptr;
    unsigned long long* d_out_found = nullptr;
    size_t out_bytes = 0, off_n = 0, off_found = 0;
    bool launched = false;
};

// Pinned staging + device posting array of one segment being uploaded (ns_upload_*).
struct ns_upload {
    ns_index* owner = nullptr;
    uint64_t P = 0;
    uint2* d_post = nullptr;
    uint8_t* h_pinned = nullptr;
    cudaStream_t stream = nullptr;
    std::mutex mu;  // ns_upload_push may be called from the threads that read the barrel files
};

// Peer exchange of result blobs (see PublishDest in bm25_kernels.cuh).  One per (rank, device).
struct ns_exchange {
    int device = 0;
    uint32_t world = 1, rank = 0, slots = 1, max_q = 0;
    size_t stride = 0;        // bytes reserved per rank inside a gather region (blob of max_q queries at NS_MAX_K)
    size_t slot_bytes = 0;    // world * stride
    size_t off_flags = 0, off_status = 0, off_merged = 0, off_pub = 0, total = 0;
    uint8_t* d_mem = nullptr; // gather[slots][world][stride] | flags[slots][64] | status[slots] | merged[slots][stride] | PublishDest[slots]
    struct Dest {
        uint32_t rank;
        uint8_t* base;        // the destination's d_mem as mapped into this process / device
        bool ipc;
    };
    std::vector<Dest> dests;
    bool receiver = false;    // this rank is one of its own destinations: it waits for all ranks and merges
    bool publisher_only = false;  // created without gather / merged / pinned buffers (cannot receive, cannot be a destination)
    uint64_t timeout_ns = 10000000000ull;  // NSB200_EXCHANGE_TIMEOUT_MS, default 10 s
    uint8_t* h_out = nullptr; // pinned, stride + 256 bytes
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_done[8] = {};  // per slot: merge of the last step using the slot has been enqueued
    cudaEvent_t ev_copy = nullptr;  // result copy finished (host waits on it, sleeping)
    std::mutex mu;
    uint8_t* gather(uint8_t* base, uint32_t slot) const { return base + (size_t)slot * slot_bytes; }
    uint32_t* flags(uint8_t* base, uint32_t slot) const { return reinterpret_cast<uint32_t*>(base + off_flags + (size_t)slot * 256); }
    uint32_t* status(uint32_t slot) const { return reinterpret_cast<uint32_t*>(d_mem + off_status + (size_t)slot * 256); }
    uint8_t* merged(uint32_t slot) const { return d_mem + off_merged + (size_t)slot * stride; }
    PublishDest* pub(uint32_t slot) const { return reinterpret_cast<PublishDest*>(d_mem + off_pub) + slot; }
};

extern "C" const char* ns_last_error(void) { return last_error(); }

extern "C" int ns_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

namespace {

struct KernelCfg {
    const void* fn;
    size_t smem;
    int threads = kThreads;
};

template <int TDW, int KCAP, bool FAST, bool IMPACT, int NG, bool PUB>
KernelCfg cfg_of() {
    return KernelCfg{(const void*)bm25_score_topk_kernel<TDW, KCAP, FAST, IMPACT, NG, PUB>,
                     sizeof(WarpSmem<TDW, KCAP>) * kWarpsPerBlock};
}

template <int TDW, int KCAP, int NG, bool PUB>
KernelCfg pick_kernel_tk(bool fast, bool impact) {
    if (fast) return impact ? cfg_of<TDW, KCAP, true, true, NG, PUB>() : cfg_of<TDW, KCAP, true, false, NG, PUB>();
    return impact ? cfg_of<TDW, KCAP, false, true, NG, PUB>() : cfg_of<TDW, KCAP, false, false, NG, PUB>();
}

template <int NG, bool PUB>
KernelCfg pick_kernel_ng(uint32_t k, bool fast, bool impact) {
    return k <= 16 ? pick_kernel_tk<kTileDocs, 16, NG, PUB>(fast, impact) : pick_kernel_tk<kTileDocs, 100, NG, PUB>(fast, impact);
}

// The long-query variants (more than 64 terms in one (query, segment)) exist in the generic-arithmetic form only:
// FAST differs from it by a cheaper, equally rounded division and a skipped multiplication by 1.0f, never in results.
template <int TDW, int KCAP, int NG, bool PUB>
KernelCfg pick_kernel_long(bool impact) {
    return impact ? cfg_of<TDW, KCAP, false, true, NG, PUB>() : cfg_of<TDW, KCAP, false, false, NG, PUB>();
}

template <bool PUB>
KernelCfg pick_kernel_p(uint32_t k, bool fast, bool impact, uint32_t ng) {
    switch (ng) {
        case 1: return pick_kernel_ng<1, PUB>(k, fast, impact);
        case 2: return pick_kernel_ng<2, PUB>(k, fast, impact);
        case 4: return k <= 16 ? pick_kernel_long<kTileDocs, 16, 4, PUB>(impact) : pick_kernel_long<kTileDocs, 100, 4, PUB>(impact);
        default: return k <= 16 ? pick_kernel_long<kTileDocs, 16, 8, PUB>(impact) : pick_kernel_long<kTileDocs, 100, 8, PUB>(impact);
    }
}

// 32-term register groups per lane for a batch whose largest (query, segment) term list has `max_in_seg` entries
uint32_t term_groups(uint32_t max_in_seg) { return max_in_seg <= 32u ? 1u : max_in_seg <= 64u ? 2u : max_in_seg <= 128u ? 4u : 8u; }
static_assert(NS_MAX_TERMS == 8 * 32, "the widest kernel variant holds 8 groups of 32 terms");

// ng:  32-term register groups per lane (term_groups)
// pub: the score kernel publishes finished queries to peer GPUs (multi-GPU exchange)
KernelCfg pick_kernel(uint32_t k, bool fast, bool impact, uint32_t ng, bool pub) {
    return pub ? pick_kernel_p<true>(k, fast, impact, ng) : pick_kernel_p<false>(k, fast, impact, ng);
}

// The dynamic shared-memory opt-in and the occupancy query of every kernel variant, done once per index
// handle: neither belongs on the launch path (cudaFuncSetAttribute may wait for the device, which must not
// happen while a peer-exchange wait kernel is spinning).
int init_kernel_table(IndexShared* idx) {
    for (int kk = 0; kk < 2; kk++)
        for (int fast = 0; fast < 2; fast++)
            for (int impact = 0; impact < 2; impact++)
                for (int wp = 0; wp < 8; wp++) {
                    const KernelCfg cfg = pick_kernel(kk ? 100u : 10u, fast != 0, impact != 0, 1u << (wp & 3), (wp & 4) != 0);
                    int per_sm = 0;
                    NS_CUDA(cudaFuncSetAttribute(cfg.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
                    NS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cfg.fn, cfg.threads, cfg.smem));
                    idx->occupancy[cfg.fn] = per_sm;
                }
    return NS_OK;
}

}  // namespace

static int ns_index_create_impl(int device, ns_index** out) {
    if (!out) { set_error("ns_index_create: out is null"); return NS_ERR_INVALID; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                  "); this library has no CPU fallback");
        return NS_ERR_CUDA;
    }
    if (device < 0 || device >= n) { set_error("ns_index_create: device out of range"); return NS_ERR_INVALID; }
    NS_CUDA(cudaSetDevice(device));
    auto* idx = new ns_index();
    idx->device = device;
    idx->sh = std::make_shared<IndexShared>();
    idx->sh->device = device;
    idx->sh->tun = Tunables::from_env();
    cudaDeviceProp prop;
    NS_CUDA(cudaGetDeviceProperties(&prop, device));
    idx->sh->sm_count = prop.multiProcessorCount;
    int rc = init_kernel_table(idx->sh.get());
    if (rc != NS_OK) {
        delete idx;
        return rc;
    }
    *out = idx;
    return NS_OK;
}

extern "C" int ns_index_create(int device, ns_index** out) {
    return abi_guard("ns_index_create", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_index_create_impl(device, out); });
}

extern "C" void ns_index_destroy(ns_index* idx) {
    if (!idx) return;
    cudaSetDevice(idx->device);
    for (auto& s : idx->staged) s.release();
    {
        std::lock_guard<std::mutex> lk(idx->sh->mu);
        idx->sh->closed = true;  // batches still alive free their resources themselves
        idx->sh->pool.clear();
    }
    idx->live.reset();
    delete idx;
}

namespace {

// Everything after the raw postings are on the device: validate, tile table, doc-length factors, packing,
// resident impacts; stages the segment on success.  Takes ownership of d_post (freed on failure).
int add_segment_core(ns_index* idx, uint32_t global_seg, uint32_t N, float avgdl, const uint32_t* doc_len, uint32_t T,
                     const uint64_t* term_begin, const uint32_t* term_count, const float* row_idf, uint2* d_post,
                     uint64_t P, uint32_t flags) {
    SegState s;
    s.d_post = d_post;
    std::vector<uint32_t> begin32(T);
    for (uint32_t t = 0; t < T; t++) {
        // overflow-safe form of begin + count <= P
        if ((uint64_t)term_count[t] > P || term_begin[t] > P - (uint64_t)term_count[t]) {
            s.release();
            set_error("ns_index_add_segment: row " + std::to_string(t) + " exceeds the posting array");
            return NS_ERR_FORMAT;
        }
        begin32[t] = (uint32_t)term_begin[t];
    }
    // One resident score per posting (and its in-place build) needs every posting to belong to at most one row.
    // Rows that share postings are legal for the reference — its loop only follows (offset, count),
    // src/api_engine.cpp:469-476 — and would need one score per (row, posting): such a segment stays on the
    // raw-posting path (no resident scores, {docId, tf} kept whatever the flags say; the per-batch pre-pass
    // evaluates each named row under its own idf).
    bool rows_overlap = false;
    {
        std::vector<uint32_t> by_begin;
        by_begin.reserve(T);
        for (uint32_t t = 0; t < T; t++)
            if (term_count[t]) by_begin.push_back(t);
        std::sort(by_begin.begin(), by_begin.end(), [&](uint32_t a, uint32_t b) { return begin32[a] < begin32[b]; });
        for (size_t i = 1; i < by_begin.size() && !rows_overlap; i++)
            rows_overlap = (uint64_t)begin32[by_begin[i - 1]] + term_count[by_begin[i - 1]] > begin32[by_begin[i]];
    }
    s.gseg = global_seg;
    s.ndocs = N;
    s.T = T;
    s.P = P;
    s.avgdl = avgdl;
    s.ntiles = (N + idx->tile_docs - 1) / idx->tile_docs;
    if (s.ntiles == 0) s.ntiles = 1;
    s.h_count.assign(term_count, term_count + T);
    s.h_begin = begin32;

    // Distinct doc lengths -> 16-bit codes (packed payload).  Falls back to the per-doc norm array
    // when the segment has more than 65536 distinct lengths or (checked on the device) a tf >= 65536.
    std::vector<uint32_t> uniq(doc_len, doc_len + N);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    bool can_pack = uniq.size() <= 65536 && !idx->sh->tun.no_pack;
    std::vector<unsigned short> code;
    if (can_pack) {
        code.resize(N);
        for (uint32_t d = 0; d < N; d++)
            code[d] = (unsigned short)(std::lower_bound(uniq.begin(), uniq.end(), doc_len[d]) - uniq.begin());
    }

    uint32_t *d_len = nullptr, *d_begin = nullptr, *d_count = nullptr;
    unsigned short* d_code = nullptr;
    float* d_idf = nullptr;
    unsigned int* d_err = nullptr;  // [0] posting-order violations, [1] norms outside the fast-division range, [2] tf > 0xFFFF
    auto cleanup = [&]() {
        if (d_len) cudaFree(d_len);
        if (d_begin) cudaFree(d_begin);
        if (d_count) cudaFree(d_count);
        if (d_code) cudaFree(d_code);
        if (d_idf) cudaFree(d_idf);
        if (d_err) cudaFree(d_err);
    };
#define NS_CUDA_SEG(expr)                                                                       \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            set_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr);     \
            cleanup();                                                                          \
            s.release();                                                                        \
            return NS_ERR_CUDA;                                                                 \
        }                                                                                       \
    } while (0)

    const size_t tile_entries = (size_t)T * (s.ntiles + 1);
    const size_t nlen = can_pack ? uniq.size() : (size_t)N;  // lengths the norm kernel evaluates
    NS_CUDA_SEG(cudaMalloc(&s.d_tileoff, std::max<size_t>(16, tile_entries * sizeof(uint32_t))));
    NS_CUDA_SEG(cudaMalloc(&d_len, std::max<size_t>(16, nlen * 4)));
    NS_CUDA_SEG(cudaMalloc(&d_begin, std::max<size_t>(16, (size_t)T * 4)));
    NS_CUDA_SEG(cudaMalloc(&d_count, std::max<size_t>(16, (size_t)T * 4)));
    NS_CUDA_SEG(cudaMalloc(&d_err, 3 * sizeof(unsigned int)));
    NS_CUDA_SEG(cudaMemset(d_err, 0, 3 * sizeof(unsigned int)));
    unsigned int h_errs[3] = {0, 0, 0};
    if (T) {
        NS_CUDA_SEG(cudaMemcpy(d_begin, begin32.data(), (size_t)T * 4, cudaMemcpyHostToDevice));
        NS_CUDA_SEG(cudaMemcpy(d_count, term_count, (size_t)T * 4, cudaMemcpyHostToDevice));
        const int blocks = (int)std::min<uint64_t>(((uint64_t)T + 7) / 8, (uint64_t)idx->sh->sm_count * 32);
        validate_rows_kernel<<<blocks, 256>>>(s.d_post, d_begin, d_count, T, N, d_err);
        NS_CUDA_SEG(cudaGetLastError());
        NS_CUDA_SEG(cudaMemcpy(h_errs, d_err, sizeof(h_errs), cudaMemcpyDeviceToHost));
        // Reject BEFORE any kernel indexes a per-doc array with a posting's docId: every later kernel
        // (tile table, packing, impacts) relies on docId < N and sorted rows.
        if (h_errs[0]) {
            cleanup();
            s.release();
            set_error("segment " + std::to_string(global_seg) + ": " + std::to_string(h_errs[0]) +
                      " postings are out of order, duplicated or have docId >= N");
            return NS_ERR_FORMAT;
        }
        const uint64_t tb = (tile_entries + 255) / 256;
        tile_table_kernel<<<(int)std::min<uint64_t>(tb, (uint64_t)idx->sh->sm_count * 64), 256>>>(
            s.d_post, d_begin, d_count, T, s.ntiles, idx->tile_docs, s.d_tileoff);
        NS_CUDA_SEG(cudaGetLastError());
    }
    if (h_errs[2]) can_pack = false;
    float* d_normdst = nullptr;
    if (can_pack) {
        NS_CUDA_SEG(cudaMalloc(&s.d_lut, std::max<size_t>(16, uniq.size() * sizeof(float))));
        if (!uniq.empty()) NS_CUDA_SEG(cudaMemcpy(d_len, uniq.data(), uniq.size() * 4, cudaMemcpyHostToDevice));
        d_normdst = s.d_lut;
    } else {
        if (nlen < (size_t)N) {  // d_len was sized for the distinct lengths: regrow
            cudaFree(d_len);
            d_len = nullptr;
            NS_CUDA_SEG(cudaMalloc(&d_len, std::max<size_t>(16, (size_t)N * 4)));
        }
        NS_CUDA_SEG(cudaMalloc(&s.d_norm, std::max<size_t>(16, (size_t)N * sizeof(float))));
        if (N) NS_CUDA_SEG(cudaMemcpy(d_len, doc_len, (size_t)N * 4, cudaMemcpyHostToDevice));
        d_normdst = s.d_norm;
    }
    const uint32_t nnorm = can_pack ? (uint32_t)uniq.size() : N;
    if (nnorm) {
        doc_norm_kernel<<<(nnorm + 255) / 256, 256>>>(d_len, d_normdst, nnorm, avgdl, kK1, kB, d_err + 1);
        NS_CUDA_SEG(cudaGetLastError());
    }
    if (can_pack && P) {
        NS_CUDA_SEG(cudaMalloc(&d_code, std::max<size_t>(16, (size_t)N * 2)));
        NS_CUDA_SEG(cudaMemcpy(d_code, code.data(), (size_t)N * 2, cudaMemcpyHostToDevice));
        pack_postings_kernel<<<idx->sh->sm_count * 16, 256>>>(s.d_post, P, d_code, N);
        NS_CUDA_SEG(cudaGetLastError());
    }
    s.packed = can_pack;
    // Resident impacts: the term score of every posting under the row's idf — the caller's row_idf[]
    // (the engine passes bm25_idf(N, df) of its lexicon) or bm25_idf(N, count) (src/api_engine.cpp:45-47 with
    // df = LexEntry.count, what every writer stores: include/segment_writer.hpp:147-149).  A query term whose
    // idf differs bit-wise from the row's goes through the per-batch pre-pass, which needs the raw postings.
    const bool drop_raw = (flags & NS_SEG_DROP_RAW) != 0u;
    if (P && T && !idx->sh->tun.no_resident && !rows_overlap) {
        std::vector<float> h_idf(T);
        s.h_idf_bits.resize(T);
        for (uint32_t t = 0; t < T; t++) {
            h_idf[t] = row_idf ? row_idf[t] : bm25_idf(N, term_count[t]);
            std::memcpy(&s.h_idf_bits[t], &h_idf[t], 4);
        }
        NS_CUDA_SEG(cudaMalloc(&d_idf, (size_t)T * 4));
        NS_CUDA_SEG(cudaMemcpy(d_idf, h_idf.data(), (size_t)T * 4, cudaMemcpyHostToDevice));
        if (drop_raw) {
            s.d_imp = s.d_post;  // in place: the raw payload of every row-covered posting is replaced by its score
            s.d_post = nullptr;
        } else {
            NS_CUDA_SEG(cudaMalloc(&s.d_imp, (P + 2) * sizeof(uint2)));
        }
        const int blocks = (int)std::min<uint64_t>(((uint64_t)T + 7) / 8, (uint64_t)idx->sh->sm_count * 32);
        build_impacts_kernel<<<blocks, 256>>>(drop_raw ? s.d_imp : s.d_post, d_begin, d_count, d_idf, T, d_normdst,
                                              can_pack ? 1u : 0u, kK1 + 1.0f, s.d_imp);
        NS_CUDA_SEG(cudaGetLastError());
    }
    NS_CUDA_SEG(cudaDeviceSynchronize());
    NS_CUDA_SEG(cudaMemcpy(h_errs, d_err, sizeof(h_errs), cudaMemcpyDeviceToHost));
    s.norm_in_range = (h_errs[1] == 0);
    cleanup();
#undef NS_CUDA_SEG
    s.bytes = ((s.d_imp ? 1 : 0) + (s.d_post ? 1 : 0)) * P * sizeof(uint2) +
              (can_pack ? (uint64_t)uniq.size() * 4 : (uint64_t)N * 4) + tile_entries * 4;
    std::lock_guard<std::mutex> lk(idx->mu);
    for (auto& o : idx->staged) {
        if (o.gseg == global_seg) {
            s.release();
            set_error("segment " + std::to_string(global_seg) + " staged twice");
            return NS_ERR_INVALID;
        }
    }
    idx->staged.push_back(std::move(s));
    return NS_OK;
}

int check_segment_args(ns_index* idx, uint32_t global_seg, uint32_t N, const uint32_t* doc_len, uint32_t T,
                       const uint64_t* term_begin, const uint32_t* term_count, uint64_t P) {
    if (!idx || (N && !doc_len) || (T && (!term_begin || !term_count))) {
        set_error("ns_index_add_segment: null argument");
        return NS_ERR_INVALID;
    }
    if (P >= 0xFFFFFFFFull) { set_error("segment has >= 2^32 postings; split it"); return NS_ERR_INVALID; }
    if (global_seg >= 0x80000000u) { set_error("ns_index_add_segment: global_seg must be < 2^31"); return NS_ERR_INVALID; }
    return NS_OK;
}

}  // namespace

static int ns_index_add_segment_ex_impl(ns_index* idx, uint32_t global_seg, uint32_t N, float avgdl,
                                       const uint32_t* doc_len, uint32_t T, const uint64_t* term_begin,
                                       const uint32_t* term_count, const float* row_idf, const void* postings,
                                       uint64_t P, uint32_t flags) {
    int rc = check_segment_args(idx, global_seg, N, doc_len, T, term_begin, term_count, P);
    if (rc != NS_OK) return rc;
    if (P && !postings) { set_error("ns_index_add_segment: null argument"); return NS_ERR_INVALID; }
    NS_CUDA(cudaSetDevice(idx->device));
    uint2* d_post = nullptr;
    NS_CUDA(cudaMalloc(&d_post, (P + 2) * sizeof(uint2)));  // +2: bulk prefetches round slices up to 16 B
    if (P) {
        cudaError_t e = cudaMemcpy(d_post, postings, P * sizeof(uint2), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(d_post);
            set_error(std::string("CUDA error uploading postings: ") + cudaGetErrorString(e));
            return NS_ERR_CUDA;
        }
    }
    return add_segment_core(idx, global_seg, N, avgdl, doc_len, T, term_begin, term_count, row_idf, d_post, P, flags);
}

extern "C" int ns_index_add_segment_ex(ns_index* idx, uint32_t global_seg, uint32_t N, float avgdl,
                                       const uint32_t* doc_len, uint32_t T, const uint64_t* term_begin,
                                       const uint32_t* term_count, const float* row_idf, const void* postings,
                                       uint64_t P, uint32_t flags) {
    return abi_guard("ns_index_add_segment_ex", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_index_add_segment_ex_impl(idx, global_seg, N, avgdl, doc_len, T, term_begin, term_count, row_idf, postings, P, flags); });
}

extern "C" int ns_index_add_segment(ns_index* idx, uint32_t global_seg, uint32_t N, float avgdl,
                                    const uint32_t* doc_len, uint32_t T, const uint64_t* term_begin,
                                    const uint32_t* term_count, const void* postings, uint64_t P) {
    return ns_index_add_segment_ex(idx, global_seg, N, avgdl, doc_len, T, term_begin, term_count, nullptr, postings, P, 0u);
}

// ---- streamed upload: the loader reads barrel files straight into pinned memory and pushes every
// ---- finished range; the copies overlap the remaining file reads ----

static int ns_upload_begin_impl(ns_index* idx, uint64_t P, ns_upload** out) {
    if (!idx || !out) { set_error("ns_upload_begin: null argument"); return NS_ERR_INVALID; }
    *out = nullptr;
    if (P >= 0xFFFFFFFFull) { set_error("segment has >= 2^32 postings; split it"); return NS_ERR_INVALID; }
    NS_CUDA(cudaSetDevice(idx->device));
    auto u = std::make_unique<ns_upload>();
    u->owner = idx;
    u->P = P;
    cudaError_t e = cudaMalloc(&u->d_post, (P + 2) * sizeof(uint2));
    if (e == cudaSuccess) e = cudaHostAlloc(&u->h_pinned, std::max<size_t>(16, P * sizeof(uint2)), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&u->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        if (u->d_post) cudaFree(u->d_post);
        if (u->h_pinned) cudaFreeHost(u->h_pinned);
        set_error(std::string("ns_upload_begin: ") + cudaGetErrorString(e));
        return NS_ERR_CUDA;
    }
    *out = u.release();
    return NS_OK;
}

extern "C" int ns_upload_begin(ns_index* idx, uint64_t P, ns_upload** out) {
    return abi_guard("ns_upload_begin", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_upload_begin_impl(idx, P, out); });
}

extern "C" void* ns_upload_buffer(ns_upload* u) { return u ? u->h_pinned : nullptr; }

extern "C" int ns_upload_push(ns_upload* u, uint64_t first, uint64_t count) {
    if (!u || first > u->P || count > u->P - first) { set_error("ns_upload_push: range outside the segment"); return NS_ERR_INVALID; }
    if (count == 0) return NS_OK;
    std::lock_guard<std::mutex> lk(u->mu);
    NS_CUDA(cudaSetDevice(u->owner->device));
    NS_CUDA(cudaMemcpyAsync(u->d_post + first, u->h_pinned + first * sizeof(uint2), count * sizeof(uint2),
                            cudaMemcpyHostToDevice, u->stream));
    return NS_OK;
}

extern "C" void ns_upload_abort(ns_upload* u) {
    if (!u) return;
    cudaSetDevice(u->owner->device);
    if (u->stream) {
        cudaStreamSynchronize(u->stream);
        cudaStreamDestroy(u->stream);
    }
    if (u->d_post) cudaFree(u->d_post);
    if (u->h_pinned) cudaFreeHost(u->h_pinned);
    delete u;
}

static int ns_upload_finish_impl(ns_upload* u, uint32_t global_seg, uint32_t N, float avgdl, const uint32_t* doc_len,
                                uint32_t T, const uint64_t* term_begin, const uint32_t* term_count,
                                const float* row_idf, uint32_t flags) {
    if (!u) { set_error("ns_upload_finish: null"); return NS_ERR_INVALID; }
    ns_index* idx = u->owner;
    int rc = check_segment_args(idx, global_seg, N, doc_len, T, term_begin, term_count, u->P);
    if (rc != NS_OK) { ns_upload_abort(u); return rc; }
    cudaSetDevice(idx->device);
    cudaError_t e = cudaStreamSynchronize(u->stream);
    uint2* d_post = u->d_post;
    const uint64_t P = u->P;
    u->d_post = nullptr;
    ns_upload_abort(u);  // frees the staging, not the device array
    if (e != cudaSuccess) {
        cudaFree(d_post);
        set_error(std::string("ns_upload_finish: ") + cudaGetErrorString(e));
        return NS_ERR_CUDA;
    }
    return add_segment_core(idx, global_seg, N, avgdl, doc_len, T, term_begin, term_count, row_idf, d_post, P, flags);
}

extern "C" int ns_upload_finish(ns_upload* u, uint32_t global_seg, uint32_t N, float avgdl, const uint32_t* doc_len,
                                uint32_t T, const uint64_t* term_begin, const uint32_t* term_count,
                                const float* row_idf, uint32_t flags) {
    return abi_guard("ns_upload_finish", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_upload_finish_impl(u, global_seg, N, avgdl, doc_len, T, term_begin, term_count, row_idf, flags); });
}

extern "C" int ns_index_abort(ns_index* idx) {
    if (!idx) return NS_ERR_INVALID;
    cudaSetDevice(idx->device);
    std::lock_guard<std::mutex> lk(idx->mu);
    for (auto& s : idx->staged) s.release();
    idx->staged.clear();
    return NS_OK;
}

static int ns_index_commit_impl(ns_index* idx) {
    if (!idx) { set_error("ns_index_commit: null"); return NS_ERR_INVALID; }
    NS_CUDA(cudaSetDevice(idx->device));
    std::lock_guard<std::mutex> lk(idx->mu);
    auto st = std::make_shared<IndexState>();
    st->device = idx->device;
    st->tile_docs = idx->tile_docs;
    std::sort(idx->staged.begin(), idx->staged.end(), [](const SegState& a, const SegState& b) { return a.gseg < b.gseg; });
    std::vector<DevSeg> h_segs;
    std::vector<uint32_t> h_base(1, 0);
    for (auto& s : idx->staged) {
        DevSeg d;
        d.post = s.d_post;
        d.norm = s.d_norm;
        d.lut = s.d_lut;
        d.packed = s.packed ? 1u : 0u;
        d.pad_ = 0;
        d.tileoff = s.d_tileoff;
        d.imp = s.d_imp;
        d.ndocs = s.ndocs;
        d.T = s.T;
        d.ntiles = s.ntiles;
        d.gseg = s.gseg;
        h_segs.push_back(d);
        h_base.push_back(h_base.back() + s.ntiles);
        st->bytes += s.bytes;
    }
    const size_t nseg = h_segs.size();
    cudaError_t e = cudaMalloc(&st->d_segs, std::max<size_t>(16, nseg * sizeof(DevSeg)));
    if (e == cudaSuccess) e = cudaMalloc(&st->d_tile_base, (nseg + 1) * sizeof(uint32_t));
    if (e == cudaSuccess && nseg) e = cudaMemcpy(st->d_segs, h_segs.data(), nseg * sizeof(DevSeg), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(st->d_tile_base, h_base.data(), (nseg + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        // st's destructor frees what was allocated; staged segments stay staged, live index untouched
        set_error(std::string("CUDA error in ns_index_commit: ") + cudaGetErrorString(e));
        return NS_ERR_CUDA;
    }
    st->total_tiles = h_base.back();
    st->segs = std::move(idx->staged);
    idx->staged.clear();
    for (uint32_t i = 0; i < st->segs.size(); i++) st->slot_of[st->segs[i].gseg] = i;
    idx->live = st;  // in-flight batches keep their own shared_ptr to the old state
    return NS_OK;
}

extern "C" int ns_index_commit(ns_index* idx) {
    return abi_guard("ns_index_commit", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_index_commit_impl(idx); });
}

extern "C" int ns_index_num_segments(const ns_index* idx) {
    if (!idx) return 0;
    auto* m = const_cast<ns_index*>(idx);
    std::lock_guard<std::mutex> lk(m->mu);
    return m->live ? (int)m->live->segs.size() : 0;
}

extern "C" uint64_t ns_index_device_bytes(const ns_index* idx) {
    if (!idx) return 0;
    auto* m = const_cast<ns_index*>(idx);
    std::lock_guard<std::mutex> lk(m->mu);
    return m->live ? m->live->bytes : 0;
}

std::shared_ptr<const void> nsb::index_live_state(ns_index* idx) {
    if (!idx) return nullptr;
    std::lock_guard<std::mutex> lk(idx->mu);
    return std::static_pointer_cast<const void>(idx->live);
}

// ---------------------------------------------------------------------------------------------

namespace {

int acquire_res(IndexShared* idx, size_t d_need, size_t in_need, size_t out_need, std::unique_ptr<BatchRes>& out) {
    {
        std::lock_guard<std::mutex> lk(idx->mu);
        for (size_t i = 0; i < idx->pool.size(); i++) {
            auto& r = idx->pool[i];
            if (r->d_cap >= d_need && r->h_in_cap >= in_need && r->h_out_cap >= out_need) {
                out = std::move(r);
                idx->pool.erase(idx->pool.begin() + i);
                return NS_OK;
            }
        }
    }
    auto r = std::make_unique<BatchRes>();
    r->device = idx->device;
    // round capacities up so that a stream of similar batches reuses one resource set
    auto grow = [](size_t n) { size_t c = 1 << 16; while (c < n) c <<= 1; return c; };
    r->d_cap = grow(d_need);
    r->h_in_cap = grow(in_need);
    r->h_out_cap = grow(out_need);
    NS_CUDA(cudaMalloc(&r->d_blob, r->d_cap));
    NS_CUDA(cudaMallocHost(&r->h_in, r->h_in_cap));
    NS_CUDA(cudaMallocHost(&r->h_out, r->h_out_cap));
    NS_CUDA(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    // Blocking-sync events: a host thread that waits for the GPU sleeps instead of spinning, so that many
    // concurrent callers (and the front-end worker threads) do not starve each other of cores.
    for (auto& e : r->ev) NS_CUDA(cudaEventCreateWithFlags(&e, cudaEventBlockingSync));
    NS_CUDA(cudaEventCreateWithFlags(&r->ev_h2d, cudaEventDisableTiming | cudaEventBlockingSync));
    NS_CUDA(cudaEventCreateWithFlags(&r->ev_copy, cudaEventDisableTiming | cudaEventBlockingSync));
    out = std::move(r);
    return NS_OK;
}

constexpr uint32_t kMaxSplit = 64;

// Cut queries into items — (query, window of consecutive tiles) — and write what the kernel needs into the
// pinned input blob at off_items.  Normal case: window-major order with the SAME number of windows for
// every query, so only the query order[] is written and the kernel derives item i = (order[i % Q], i / Q).
// forced > 0 (tests) gives every query exactly min(forced, tiles) items, written as an explicit list.
// Returns the number of bytes written at off_items.
size_t build_items(ns_batch* b, uint32_t forced) {
    const IndexState& st = *b->st;
    const Tunables& tun = b->owner->tun;
    const uint32_t Q = b->Q;
    const uint32_t tiles = std::max<uint32_t>(1, st.total_tiles);
    uint64_t target = tun.item_postings > 0 ? (uint64_t)tun.item_postings : 32768;
    // small batches: cut finer so that every resident warp has work
    const uint64_t want_items = (uint64_t)b->owner->sm_count * 24 * 2;
    if (Q > 0 && b->postings / target + Q < want_items) target = std::max<uint64_t>(1024, b->postings / want_items);
    // Items per query: ~10 items per resident warp over the whole batch, at least tiles/24 (windows of
    // <= 24 tiles keep the batch's hot slices in L2), at most tiles/4 — and at most 16 for small batches,
    // where all items of a query run at the same time and serialise on the query's result-list lock.
    uint32_t window;
    {
        const uint64_t warps = (uint64_t)b->owner->sm_count * 24;
        uint64_t ns = warps * 10 / std::max<uint32_t>(1, Q);
        ns = std::max<uint64_t>(ns, (tiles + 23) / 24);
        ns = std::min<uint64_t>(ns, std::max<uint32_t>(1, tiles / 4));
        if (Q < 256) ns = std::min<uint64_t>(ns, 16);
        ns = std::max<uint64_t>(1, ns);
        window = (uint32_t)((tiles + ns - 1) / ns);
    }
    if (tun.window_tiles >= 0) window = (uint32_t)tun.window_tiles;  // 0 = query-major
    std::vector<uint32_t> nsplit(Q);
    uint32_t maxs = 1;
    uint64_t nitems = 0;
    for (uint32_t q = 0; q < Q; q++) {
        uint64_t ns = forced ? forced : (b->weight[q] + target - 1) / target;
        if (!forced && window) ns = (tiles + window - 1) / window;
        ns = std::max<uint64_t>(1, std::min<uint64_t>(ns, std::min<uint64_t>(tiles, kMaxSplit)));
        nsplit[q] = (uint32_t)ns;
        maxs = std::max(maxs, nsplit[q]);
        nitems += ns;
    }
    // Heaviest queries first inside a window (query-major mode: heaviest items first).
    std::vector<uint32_t> order(Q);
    std::iota(order.begin(), order.end(), 0u);
    std::vector<uint64_t> iw(Q);
    for (uint32_t q = 0; q < Q; q++) iw[q] = b->weight[q] / nsplit[q];
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return iw[x] > iw[y]; });
    uint8_t* dst = b->res->h_in + b->off_items;
    b->window_major = window && !forced;
    b->implicit_items = b->window_major && !tun.explicit_items;
    b->nsplit_all = maxs;
    b->nitems = (uint32_t)nitems;
    b->max_split = maxs;
    if (b->implicit_items) {
        if (Q) std::memcpy(dst, order.data(), (size_t)Q * 4);
        return (size_t)Q * 4;
    }
    DevItem* h_items = reinterpret_cast<DevItem*>(dst);
    size_t at = 0;
    if (b->window_major) {
        for (uint32_t sp = 0; sp < maxs; sp++)
            for (uint32_t q : order)
                if (sp < nsplit[q]) h_items[at++] = DevItem{q, (sp << 16) | nsplit[q]};
    } else {
        for (uint32_t q : order)
            for (uint32_t sp = 0; sp < nsplit[q]; sp++) h_items[at++] = DevItem{q, (sp << 16) | nsplit[q]};
    }
    return at * sizeof(DevItem);
}

}  // namespace

namespace {

// Second half of every prepare: sizes the batch's device blob, fills the pinned input blob, builds the items and
// starts the upload.  `kept` are the batch's terms in kernel form; dist/dstart describe the terms whose scores are
// not resident (empty on the engine's trusted path).
int finish_prepare(ns_index* idx, const std::shared_ptr<IndexState>& st, std::unique_ptr<ns_batch> b, uint32_t Q, uint32_t k,
                   const DevTerm* kept, size_t nkept, const uint32_t* qoff32, const std::vector<DevDistinct>& dist,
                   const std::vector<uint32_t>& dstart, uint64_t dist_post, uint64_t total_post, uint64_t nonres_post,
                   uint64_t n_resident, uint32_t max_in_seg, bool scan_always, bool fast, bool all_raw, ns_batch** out) {
    b->owner = idx->sh;
    b->st = st;
    b->Q = Q;
    b->k = k;
    b->nterms = nkept;
    b->postings = total_post;
    b->scan_always = scan_always;
    b->max_in_seg = max_in_seg;
    b->fast = fast && !scan_always;
    b->ndist = (uint32_t)dist.size();
    b->dist_postings = dist_post;
    // Share term scores across the batch when that removes enough evaluations: the pre-pass reads
    // and writes every distinct posting once, the scoring kernel then skips ~2/3 of its arithmetic.
    // Segments without raw postings can only be scored from impacts.
    {
        const double share = dist_post ? (double)nonres_post / (double)dist_post : 0.0;
        const int env = idx->sh->tun.impact;
        b->impact = !all_raw || n_resident > 0 || (env >= 0 ? (env != 0 && dist_post > 0) : (share >= 1.5));
    }

    const size_t tiles = std::max<uint32_t>(1, st->total_tiles);
    b->items_cap = (size_t)Q * std::min<size_t>(tiles, kMaxSplit);
    const size_t sz_qoff = align_up(((size_t)Q + 1) * 4);
    const size_t sz_terms = align_up(std::max<size_t>(1, nkept) * sizeof(DevTerm));
    const size_t sz_dist = align_up(std::max<size_t>(1, dist.size()) * sizeof(DevDistinct));
    const size_t sz_dstart = align_up(dstart.size() * 4);
    // Room for an explicit item list (ns_batch_set_splits, query-major mode) is only reserved while it is small;
    // large batches run with implicit items (order[] only) and keep their pinned / device blobs small.
    const bool explicit_mode = idx->sh->tun.explicit_items || idx->sh->tun.window_tiles == 0;
    if (!explicit_mode && b->items_cap * sizeof(DevItem) > (1u << 20)) b->items_cap = ((size_t)Q + 1) / 2;  // = Q * 4 bytes of order[]
    const size_t sz_items = align_up(std::max<size_t>(1, b->items_cap) * sizeof(DevItem));
    const size_t off_dist = sz_qoff + sz_terms;
    const size_t off_dstart = off_dist + sz_dist;
    b->off_items = off_dstart + sz_dstart;  // last uploaded region: only its used prefix is copied
    const size_t in_cap = b->off_items + sz_items;
    // zeroed per launch: queue head | locks[Q] | done[Q] | published | result blob
    const size_t off_qdone = ((2 + (size_t)Q) * 4 + 7) / 8 * 8;  // u64 [Q] after queue head, published count and locks
    const size_t sz_ctrl = align_up(off_qdone + (size_t)Q * 8);
    const size_t sz_hits = align_up(std::max<size_t>(1, (size_t)Q * k) * sizeof(ns_hit));
    const size_t sz_n = align_up(std::max<size_t>(1, Q) * 4);
    const size_t sz_found = align_up(std::max<size_t>(1, Q) * 8);
    b->out_bytes = sz_hits + sz_n + sz_found;
    b->off_n = sz_hits;
    b->off_found = sz_hits + sz_n;
    b->off_zero = in_cap;
    b->zero_bytes = sz_ctrl + b->out_bytes;

    int rc = acquire_res(idx->sh.get(), in_cap + b->zero_bytes, in_cap, b->out_bytes, b->res);
    if (rc != NS_OK) return rc;
    BatchRes& r = *b->res;
    std::memcpy(r.h_in, qoff32, ((size_t)Q + 1) * 4);
    if (nkept != 0) std::memcpy(r.h_in + sz_qoff, kept, nkept * sizeof(DevTerm));
    if (!dist.empty()) std::memcpy(r.h_in + off_dist, dist.data(), dist.size() * sizeof(DevDistinct));
    std::memcpy(r.h_in + off_dstart, dstart.data(), dstart.size() * 4);
    b->d_dist = reinterpret_cast<DevDistinct*>(r.d_blob + off_dist);
    b->d_dstart = reinterpret_cast<uint32_t*>(r.d_blob + off_dstart);
    if (b->impact && dist_post) {
        const size_t need = (dist_post + 4) * sizeof(uint2);
        if (r.scratch_cap < need) {
            if (r.d_scratch) cudaFree(r.d_scratch);
            r.d_scratch = nullptr;
            r.scratch_cap = 0;
            size_t cap = 1 << 20;
            while (cap < need) cap <<= 1;
            NS_CUDA(cudaMalloc(&r.d_scratch, cap));
            r.scratch_cap = cap;
        }
    }
    const size_t items_bytes = build_items(b.get(), 0);
    b->up_bytes = b->off_items + items_bytes;
    b->d_qoff = reinterpret_cast<uint32_t*>(r.d_blob);
    b->d_terms = reinterpret_cast<DevTerm*>(r.d_blob + sz_qoff);
    b->d_items = reinterpret_cast<DevItem*>(r.d_blob + b->off_items);
    uint32_t* ctrl = reinterpret_cast<uint32_t*>(r.d_blob + b->off_zero);
    b->d_counter = ctrl;
    b->d_npub = ctrl + 1;
    b->d_qlock = ctrl + 2;
    b->d_qdone = reinterpret_cast<unsigned long long*>(r.d_blob + b->off_zero + off_qdone);
    b->d_out = r.d_blob + b->off_zero + sz_ctrl;
    b->d_out_hits = reinterpret_cast<ns_hit*>(b->d_out);
    b->d_out_n = reinterpret_cast<uint32_t*>(b->d_out + b->off_n);
    b->d_out_found = reinterpret_cast<unsigned long long*>(b->d_out + b->off_found);
    // The upload is NOT waited for here: launches on any stream order themselves after ev_h2d, so the
    // caller's host thread goes on (to the launch, or to preparing the next batch) while the copy runs.
    NS_CUDA(cudaMemcpyAsync(r.d_blob, r.h_in, b->up_bytes, cudaMemcpyHostToDevice, r.stream));
    NS_CUDA(cudaEventRecord(r.ev_h2d, r.stream));
    *out = b.release();
    return NS_OK;
}

}  // namespace

int nsb::batch_prepare_on(ns_index* idx, const std::shared_ptr<const void>& state, uint32_t Q, int k_in,
                          const uint64_t* q_off, const ns_qterm* terms, ns_batch** out) {
    if (!idx || !out || !q_off || (Q && q_off[Q] && !terms)) { set_error("ns_batch_prepare: null argument"); return NS_ERR_INVALID; }
    *out = nullptr;
    std::shared_ptr<IndexState> st =
        std::const_pointer_cast<IndexState>(std::static_pointer_cast<const IndexState>(state));
    if (!st) { set_error("ns_batch_prepare: index has no committed segments"); return NS_ERR_STATE; }
    NS_CUDA(cudaSetDevice(idx->device));
    const uint32_t k = (uint32_t)std::max(1, std::min(k_in, NS_MAX_K));  // src/api_engine.cpp:377

    // ---- host pass: keep the terms of segments this index holds, validate, weigh queries ----
    const uint64_t nin = q_off[Q];
    std::vector<DevTerm> kept;
    kept.reserve(nin);
    struct DistKey {
        uint64_t slot_row;
        uint32_t idf_bits;
        bool operator==(const DistKey& o) const { return slot_row == o.slot_row && idf_bits == o.idf_bits; }
    };
    struct DistHash {
        size_t operator()(const DistKey& k) const { return (size_t)(k.slot_row * 0x9E3779B97F4A7C15ull ^ k.idf_bits); }
    };
    std::unordered_map<DistKey, uint32_t, DistHash> dist_of;  // only touched by terms whose scores are not resident
    std::vector<DevDistinct> dist;
    std::vector<uint32_t> dstart;
    uint64_t dist_post = 0;
    std::vector<uint32_t> qoff32((size_t)Q + 1, 0);
    auto b = std::make_unique<ns_batch>();
    b->weight.assign(Q, 0);
    uint64_t total_post = 0, nonres_post = 0, n_resident = 0;
    uint32_t max_in_seg = 0;
    bool scan_always = false;
    bool fast = true;
    bool all_raw = true;
    for (auto& sg : st->segs) {
        fast = fast && sg.norm_in_range;
        all_raw = all_raw && (sg.d_post != nullptr || sg.P == 0);
        // a doc-length factor outside the validated range may be negative or NaN (e.g. avgdl <= 0 in stats.bin): term
        // scores can then be negative and partial sums are not monotone -> dense selection per tile
        if (!sg.norm_in_range) scan_always = true;
    }
    uint32_t memo_seg = 0xFFFFFFFFu;
    int64_t memo_slot = -1;
    {
        auto it = st->slot_of.find(memo_seg);
        if (it != st->slot_of.end()) memo_slot = (int64_t)it->second;
    }
    for (uint32_t q = 0; q < Q; q++) {
        if (q_off[q + 1] < q_off[q]) { set_error("ns_batch_prepare: q_off not monotone"); return NS_ERR_INVALID; }
        uint32_t prev_slot = 0, in_seg = 0;
        bool have_prev = false;
        for (uint64_t e = q_off[q]; e < q_off[q + 1]; e++) {
            const ns_qterm& t = terms[e];
            if (t.seg != memo_seg) {  // terms come grouped by segment: one map lookup per run
                auto it = st->slot_of.find(t.seg);
                memo_seg = t.seg;
                memo_slot = it == st->slot_of.end() ? -1 : (int64_t)it->second;
            }
            if (memo_slot < 0) continue;  // another device's / rank's segment
            const uint32_t slot = (uint32_t)memo_slot;
            const SegState& sg = st->segs[slot];
            if (t.row >= sg.T) { set_error("ns_batch_prepare: row out of range"); return NS_ERR_INVALID; }
            if (have_prev && slot < prev_slot) { set_error("ns_batch_prepare: terms of a query must be ordered by segment"); return NS_ERR_INVALID; }
            if (!have_prev || slot != prev_slot) in_seg = 0;
            have_prev = true;
            prev_slot = slot;
            const uint32_t cnt = sg.h_count[t.row];
            if (cnt == 0) continue;
            max_in_seg = std::max(max_in_seg, in_seg + 1);
            if (++in_seg > NS_MAX_TERMS) { set_error("ns_batch_prepare: more than NS_MAX_TERMS terms for one (query, segment)"); return NS_ERR_INVALID; }
            // negative or NaN contribution: partial sums are not monotone -> dense scan per tile
            if (!(t.weight >= 0.0f) || !(t.idf >= 0.0f)) scan_always = true;
            // FAST kernel: qweight == 1.0f and idf in [2^-40, 2^6] (see div_rn_inrange)
            if (t.weight != 1.0f || !(t.idf >= 9.094947017729282e-13f && t.idf <= 64.0f)) fast = false;
            uint32_t idf_bits;
            std::memcpy(&idf_bits, &t.idf, 4);
            if (sg.d_imp && sg.h_idf_bits[t.row] == idf_bits) {  // scores are resident: nothing to evaluate
                kept.push_back(DevTerm{slot, t.row, t.idf, t.weight, 0u, 0u});
                b->weight[q] += cnt;
                n_resident++;
                continue;
            }
            if (!sg.d_post) {
                set_error("ns_batch_prepare: a term's idf differs from the one its segment's resident scores were built with, "
                          "and the segment was uploaded without raw postings (NS_SEG_DROP_RAW)");
                return NS_ERR_STATE;
            }
            // distinct (segment, row, idf): one evaluation of the term's scores per batch
            nonres_post += cnt;
            const DistKey key{((uint64_t)slot << 32) | t.row, idf_bits};
            auto ins = dist_of.emplace(key, (uint32_t)dist.size());
            if (ins.second) {
                if (dist_post + cnt >= 0xFFFFFF00ull) { set_error("ns_batch_prepare: batch touches >= 2^32 distinct postings"); return NS_ERR_INVALID; }
                dist.push_back(DevDistinct{slot, sg.h_begin[t.row], (uint32_t)dist_post, t.idf});
                dstart.push_back((uint32_t)dist_post);
                dist_post += cnt;
            }
            const DevDistinct& dd = dist[ins.first->second];
            kept.push_back(DevTerm{slot, t.row, t.idf, t.weight, dd.dst_begin - dd.src_begin, 1u});
            b->weight[q] += cnt;
        }
        if (kept.size() > 0xFFFFFFF0ull) { set_error("ns_batch_prepare: too many terms"); return NS_ERR_INVALID; }
        qoff32[q + 1] = (uint32_t)kept.size();
        total_post += b->weight[q];
    }

    dstart.push_back((uint32_t)dist_post);
    return finish_prepare(idx, st, std::move(b), Q, k, kept.data(), kept.size(), qoff32.data(), dist, dstart, dist_post, total_post,
                          nonres_post, n_resident, max_in_seg, scan_always, fast, all_raw, out);
}

int nsb::batch_prepare_trusted(ns_index* idx, const std::shared_ptr<const void>& state, uint32_t Q, int k_in,
                               const PreparedBatch& pb, ns_batch** out) {
    static_assert(sizeof(PreparedTerm) == sizeof(DevTerm), "PreparedTerm mirrors DevTerm");
    if (!idx || !out || !pb.qoff || (Q && !pb.weight) || (Q && pb.qoff[Q] && !pb.terms)) { set_error("batch_prepare_trusted: null argument"); return NS_ERR_INVALID; }
    *out = nullptr;
    std::shared_ptr<IndexState> st = std::const_pointer_cast<IndexState>(std::static_pointer_cast<const IndexState>(state));
    if (!st) { set_error("ns_batch_prepare: index has no committed segments"); return NS_ERR_STATE; }
    bool fast = pb.unit_weights;
    bool scan_always = pb.scan_always;
    for (auto& sg : st->segs) {
        if (!sg.d_imp && sg.P) { set_error("batch_prepare_trusted: a segment has no resident scores"); return NS_ERR_STATE; }
        fast = fast && sg.norm_in_range;
        if (!sg.norm_in_range) scan_always = true;  // see batch_prepare_on
    }
    if (pb.max_in_seg > NS_MAX_TERMS) { set_error("ns_batch_prepare: more than NS_MAX_TERMS terms for one (query, segment)"); return NS_ERR_INVALID; }
    NS_CUDA(cudaSetDevice(idx->device));
    const uint32_t k = (uint32_t)std::max(1, std::min(k_in, NS_MAX_K));
    auto b = std::make_unique<ns_batch>();
    if (Q) b->weight.assign(pb.weight, pb.weight + Q);
    uint64_t total_post = 0;
    for (uint32_t q = 0; q < Q; q++) total_post += pb.weight[q];
    const std::vector<DevDistinct> dist;
    const std::vector<uint32_t> dstart(1, 0u);
    return finish_prepare(idx, st, std::move(b), Q, k, reinterpret_cast<const DevTerm*>(pb.terms), pb.qoff[Q], pb.qoff, dist, dstart,
                          0, total_post, 0, pb.qoff[Q], pb.max_in_seg, scan_always, fast, false, out);
}

static int ns_batch_prepare_impl(ns_index* idx, uint32_t Q, int k_in, const uint64_t* q_off, const ns_qterm* terms,
                                ns_batch** out) {
    if (!idx) { set_error("ns_batch_prepare: null argument"); return NS_ERR_INVALID; }
    return nsb::batch_prepare_on(idx, nsb::index_live_state(idx), Q, k_in, q_off, terms, out);
}

extern "C" int ns_batch_prepare(ns_index* idx, uint32_t Q, int k_in, const uint64_t* q_off, const ns_qterm* terms,
                                ns_batch** out) {
    return abi_guard("ns_batch_prepare", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_batch_prepare_impl(idx, Q, k_in, q_off, terms, out); });
}

extern "C" int ns_batch_set_splits(ns_batch* b, uint32_t splits) {
    if (!b) return NS_ERR_INVALID;
    NS_CUDA(cudaSetDevice(b->st->device));
    if (b->launched) NS_CUDA(cudaEventSynchronize(b->res->ev[2]));
    {
        const size_t tiles = std::max<uint32_t>(1, b->st->total_tiles);
        const size_t need = (size_t)b->Q * std::min<size_t>(std::min<size_t>(splits ? splits : kMaxSplit, tiles), kMaxSplit);
        if (need > b->items_cap) {
            set_error("ns_batch_set_splits: this batch was prepared without room for an explicit item list (large batch); "
                      "set NSB200_EXPLICIT_ITEMS=1 before creating the index");
            return NS_ERR_STATE;
        }
    }
    const size_t items_bytes = build_items(b, std::min<uint32_t>(splits, kMaxSplit));
    BatchRes& r = *b->res;
    NS_CUDA(cudaMemcpyAsync(r.d_blob + b->off_items, r.h_in + b->off_items, items_bytes, cudaMemcpyHostToDevice, r.stream));
    NS_CUDA(cudaEventRecord(r.ev_h2d, r.stream));
    return NS_OK;
}

namespace {

// Enqueues [zero control + results] [impact pre-pass] [score + top-k (+ publish to peers)] on stream s.
int launch_score(ns_batch* b, cudaStream_t s, const PublishDest* d_pub, uint32_t epoch) {
    NS_CUDA(cudaStreamWaitEvent(s, b->res->ev_h2d, 0));
    NS_CUDA(cudaEventRecord(b->res->ev[0], s));
    if (b->Q > 0) {
        // One memset: queue head, locks, completion counters AND the shared result lists.  The kernel relies on the lists
        // starting as all-zero: items read hits[q][k-1] / nhits[q] WITHOUT the lock as a lower bound of the final k-th
        // score, and a reader may see nhits == k before the entries another warp is writing are visible — what it then
        // reads is either a former k-th score or this initial 0, both valid lower bounds while all scores are >= 0
        // (batches with a negative weight or idf set scan_always and never take that shortcut).
        NS_CUDA(cudaMemsetAsync(b->res->d_blob + b->off_zero, 0, b->zero_bytes, s));
        ScoreArgs a;
        a.segs = b->st->d_segs;
        a.tile_base = b->st->d_tile_base;
        a.nseg = (uint32_t)b->st->segs.size();
        a.total_tiles = b->st->total_tiles;
        a.qoff = b->d_qoff;
        a.terms = b->d_terms;
        a.items = b->implicit_items ? nullptr : b->d_items;
        a.order = reinterpret_cast<const uint32_t*>(b->d_items);
        a.nq = b->Q;
        a.nsplit = b->nsplit_all;
        a.counter = b->d_counter;
        a.qlock = b->d_qlock;
        a.nitems = b->nitems;
        a.k = b->k;
        a.scan_always = b->scan_always ? 1u : 0u;
        a.k1p1 = kK1 + 1.0f;
        a.zero = 0u;
        a.impacts = reinterpret_cast<const uint2*>(b->res->d_scratch);
        a.any_scratch = (b->impact && b->ndist > 0) ? 1u : 0u;
        // next-window L2 prefetch pays when the items of a window really run together: window-major
        // order and enough queries to fill the machine (NSB200_L2_PREFETCH=0/1 overrides)
        const int pf = b->owner->tun.l2_prefetch;
        a.l2_prefetch = pf >= 0 ? (uint32_t)pf : (b->window_major && b->Q >= 256 ? 1u : 0u);
        a.hits = b->d_out_hits;
        a.nhits = b->d_out_n;
        a.found = b->d_out_found;
        a.pub = d_pub;
        a.pub_epoch = epoch;
        a.pub_pad = 0;
        a.pub_off_n = b->off_n;
        a.pub_off_found = b->off_found;
        a.q_done = b->d_qdone;
        a.n_published = b->d_npub;
        const bool fast = b->fast && !b->owner->tun.no_fast;
        const KernelCfg cfg = pick_kernel(b->k, fast, b->impact, term_groups(b->max_in_seg), d_pub != nullptr);
        const uint32_t warps_per_block = (uint32_t)cfg.threads / 32u;
        int per_sm = 0;
        {
            auto it = b->owner->occupancy.find(cfg.fn);  // filled at ns_index_create, read-only afterwards
            if (it != b->owner->occupancy.end()) per_sm = it->second;
        }
        if (per_sm < 1) { set_error("score kernel does not fit on an SM"); return NS_ERR_CUDA; }
        uint32_t grid = (uint32_t)b->owner->sm_count * (uint32_t)per_sm;
        grid = std::min<uint32_t>(grid, (b->nitems + warps_per_block - 1) / warps_per_block);
        grid = std::max<uint32_t>(grid, 1);
        if (b->impact && b->ndist > 0) {
            ImpactArgs ia;
            ia.segs = b->st->d_segs;
            ia.dist = b->d_dist;
            ia.dstart = b->d_dstart;
            ia.ndist = b->ndist;
            ia.total = (uint32_t)b->dist_postings;
            ia.impacts = reinterpret_cast<uint2*>(b->res->d_scratch);
            ia.k1p1 = kK1 + 1.0f;
            const uint64_t warps = (b->dist_postings + 127) / 128;
            const uint32_t ig = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((warps + 7) / 8, (uint64_t)b->owner->sm_count * 8));
            if (fast) impact_kernel<true><<<ig, 256, 0, s>>>(ia);
            else impact_kernel<false><<<ig, 256, 0, s>>>(ia);
            NS_CUDA(cudaGetLastError());
        }
        void* kargs[] = {(void*)&a};
        NS_CUDA(cudaLaunchKernel(cfg.fn, dim3(grid), dim3(cfg.threads), kargs, cfg.smem, s));
    }
    NS_CUDA(cudaEventRecord(b->res->ev[1], s));
    return NS_OK;
}

}  // namespace

extern "C" int ns_batch_launch(ns_batch* b, void* stream) {
    if (!b) { set_error("ns_batch_launch: null"); return NS_ERR_INVALID; }
    NS_CUDA(cudaSetDevice(b->st->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : b->res->stream;
    int rc = launch_score(b, s, nullptr, 0u);
    if (rc != NS_OK) return rc;
    NS_CUDA(cudaEventRecord(b->res->ev[2], s));
    b->launched = true;
    return NS_OK;
}

extern "C" int ns_batch_sync(ns_batch* b) {
    if (!b) return NS_ERR_INVALID;
    NS_CUDA(cudaSetDevice(b->st->device));
    if (b->launched) NS_CUDA(cudaEventSynchronize(b->res->ev[2]));
    return NS_OK;
}

extern "C" int ns_batch_fetch(ns_batch* b, ns_hit* out_hits, uint32_t* out_nhits, uint64_t* out_found) {
    if (!b || !b->launched) { set_error("ns_batch_fetch: batch was not launched"); return NS_ERR_STATE; }
    NS_CUDA(cudaSetDevice(b->st->device));
    BatchRes& r = *b->res;
    // order the copy after the kernels even when they ran on a caller-supplied stream
    NS_CUDA(cudaStreamWaitEvent(r.stream, r.ev[2], 0));
    NS_CUDA(cudaMemcpyAsync(r.h_out, b->d_out, b->out_bytes, cudaMemcpyDeviceToHost, r.stream));
    NS_CUDA(cudaEventRecord(r.ev_copy, r.stream));
    NS_CUDA(cudaEventSynchronize(r.ev_copy));
    if (out_hits) std::memcpy(out_hits, r.h_out, (size_t)b->Q * b->k * sizeof(ns_hit));
    if (out_nhits) std::memcpy(out_nhits, r.h_out + b->off_n, (size_t)b->Q * 4);
    if (out_found) std::memcpy(out_found, r.h_out + b->off_found, (size_t)b->Q * 8);
    return NS_OK;
}

extern "C" void ns_batch_destroy(ns_batch* b) {
    if (!b) return;
    cudaSetDevice(b->st->device);
    if (b->launched) cudaEventSynchronize(b->res->ev[2]);
    else if (b->res) cudaEventSynchronize(b->res->ev_h2d);  // the pinned blob goes back to the pool: the copy must be done
    if (b->owner && b->res) {
        std::lock_guard<std::mutex> lk(b->owner->mu);
        if (!b->owner->closed && b->owner->pool.size() < 64) b->owner->pool.push_back(std::move(b->res));
    }
    delete b;
}

extern "C" int ns_batch_device_results(ns_batch* b, void** d_hits, void** d_nhits, void** d_found) {
    if (!b) return NS_ERR_INVALID;
    if (d_hits) *d_hits = b->d_out_hits;
    if (d_nhits) *d_nhits = b->d_out_n;
    if (d_found) *d_found = b->d_out_found;
    return NS_OK;
}

extern "C" uint64_t ns_batch_posting_count(const ns_batch* b) { return b ? b->postings : 0; }
extern "C" void* ns_batch_stream(ns_batch* b) { return b && b->res ? (void*)b->res->stream : nullptr; }
extern "C" uint32_t ns_batch_num_launches(const ns_batch* b) {
    if (!b || b->Q == 0) return 0u;
    return 1u + (b->impact && b->ndist > 0 ? 1u : 0u);
}
extern "C" uint64_t ns_batch_upload_bytes(const ns_batch* b) { return b ? b->up_bytes : 0; }
extern "C" uint64_t ns_batch_result_bytes(const ns_batch* b) { return b ? b->out_bytes : 0; }

extern "C" float ns_batch_last_kernel_ms(ns_batch* b, int which) {
    if (!b || !b->launched || which < 0 || which > 1) return -1.0f;
    cudaSetDevice(b->st->device);
    if (cudaEventSynchronize(b->res->ev[2]) != cudaSuccess) return -1.0f;
    float ms = -1.0f;
    if (cudaEventElapsedTime(&ms, b->res->ev[which], b->res->ev[which + 1]) != cudaSuccess) return -1.0f;
    return ms;
}

static int ns_search_batch_impl(ns_index* idx, uint32_t Q, int k, const uint64_t* q_off, const ns_qterm* terms,
                               ns_hit* out_hits, uint32_t* out_nhits, uint64_t* out_found) {
    if (!idx) { set_error("ns_search_batch: null index"); return NS_ERR_INVALID; }
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto t0 = now();
    ns_batch* b = nullptr;
    int rc = ns_batch_prepare(idx, Q, k, q_off, terms, &b);
    if (rc != NS_OK) return rc;
    auto t1 = now();
    rc = ns_batch_launch(b, nullptr);
    auto t2 = now();
    if (rc == NS_OK) rc = ns_batch_fetch(b, out_hits, out_nhits, out_found);
    auto t3 = now();
    const uint32_t nitems = b->nitems;
    ns_batch_destroy(b);
    auto t4 = now();
    if (idx->sh->tun.trace) {
        auto ms = [](auto a, auto c) { return std::chrono::duration<double, std::milli>(c - a).count(); };
        std::fprintf(stderr, "[nsb200] Q=%u items=%u prepare %.3f ms, launch %.3f, fetch(+kernels) %.3f, destroy %.3f\n", Q,
                     nitems, ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4));
    }
    return rc;
}

extern "C" int ns_search_batch(ns_index* idx, uint32_t Q, int k, const uint64_t* q_off, const ns_qterm* terms,
                               ns_hit* out_hits, uint32_t* out_nhits, uint64_t* out_found) {
    return abi_guard("ns_search_batch", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_search_batch_impl(idx, Q, k, q_off, terms, out_hits, out_nhits, out_found); });
}

static int merge_launch(int device, uint32_t Q, int k_in, uint32_t nlists, const void* d_hits, const void* d_nhits,
                        const void* d_found, uint64_t hits_lsb, uint64_t n_lsb, uint64_t f_lsb, void* d_out_hits,
                        void* d_out_nhits, void* d_out_found, void* stream) {
    if (!d_hits || !d_nhits || !d_found || !d_out_hits || !d_out_nhits || !d_out_found) {
        set_error("ns_merge: null argument");
        return NS_ERR_INVALID;
    }
    if (nlists == 0 || nlists > (uint32_t)kMergeMaxLists) { set_error("ns_merge: nlists out of range"); return NS_ERR_INVALID; }
    NS_CUDA(cudaSetDevice(device));
    const uint32_t k = (uint32_t)std::max(1, std::min(k_in, NS_MAX_K));
    if (Q == 0) return NS_OK;
    MergeArgs m;
    m.hits = static_cast<const unsigned char*>(d_hits);
    m.nhits = static_cast<const unsigned char*>(d_nhits);
    m.found = static_cast<const unsigned char*>(d_found);
    m.hits_lsb = hits_lsb;
    m.n_lsb = n_lsb;
    m.f_lsb = f_lsb;
    m.qs = k;
    m.qs2 = 1;
    m.list_off = nullptr;
    m.Q = Q;
    m.k = k;
    m.nlists = nlists;
    m.out_hits = static_cast<ns_hit*>(d_out_hits);
    m.out_nhits = static_cast<uint32_t*>(d_out_nhits);
    m.out_found = static_cast<unsigned long long*>(d_out_found);
    const size_t smem = (size_t)kMergeWarps * nlists * sizeof(unsigned short);
    topk_merge_kernel<<<(Q + kMergeWarps - 1) / kMergeWarps, kMergeWarps * 32, smem, (cudaStream_t)stream>>>(m);
    NS_CUDA(cudaGetLastError());
    return NS_OK;
}

extern "C" int ns_merge_device(int device, uint32_t Q, int k_in, uint32_t nlists, const void* d_hits,
                               const void* d_nhits, const void* d_found, void* d_out_hits, void* d_out_nhits,
                               void* d_out_found, void* stream) {
    const uint32_t k = (uint32_t)std::max(1, std::min(k_in, NS_MAX_K));
    return merge_launch(device, Q, k_in, nlists, d_hits, d_nhits, d_found, (uint64_t)Q * k * sizeof(ns_hit),
                        (uint64_t)Q * 4, (uint64_t)Q * 8, d_out_hits, d_out_nhits, d_out_found, stream);
}

extern "C" int ns_batch_result_blob(ns_batch* b, void** d_blob, uint64_t* bytes, uint64_t* off_nhits, uint64_t* off_found) {
    if (!b) return NS_ERR_INVALID;
    if (d_blob) *d_blob = b->d_out;
    if (bytes) *bytes = b->out_bytes;
    if (off_nhits) *off_nhits = b->off_n;
    if (off_found) *off_found = b->off_found;
    return NS_OK;
}

extern "C" int ns_merge_blobs_device(int device, uint32_t Q, int k, uint32_t nlists, const void* d_blobs,
                                     uint64_t blob_stride, uint64_t off_nhits, uint64_t off_found, void* d_out_hits,
                                     void* d_out_nhits, void* d_out_found, void* stream) {
    if (!d_blobs) { set_error("ns_merge_blobs_device: null argument"); return NS_ERR_INVALID; }
    const unsigned char* base = static_cast<const unsigned char*>(d_blobs);
    return merge_launch(device, Q, k, nlists, base, base + off_nhits, base + off_found, blob_stride, blob_stride,
                        blob_stride, d_out_hits, d_out_nhits, d_out_found, stream);
}

// ---------------------------------------------------------------------------------------------
// Peer exchange (multi-GPU): result blobs travel as P2P stores issued by the score kernel itself.
// ---------------------------------------------------------------------------------------------

namespace {

size_t blob_capacity(uint32_t max_q) {
    return align_up(std::max<size_t>(1, (size_t)max_q * NS_MAX_K) * sizeof(ns_hit)) + align_up(std::max<size_t>(1, max_q) * 4) +
           align_up(std::max<size_t>(1, max_q) * 8);
}

int upload_pub(ns_exchange* x) {
    std::vector<PublishDest> h(x->slots);
    for (uint32_t s = 0; s < x->slots; s++) {
        PublishDest& p = h[s];
        std::memset(&p, 0, sizeof(p));
        p.ndest = (uint32_t)x->dests.size();
        p.src = x->rank;
        p.stride = x->stride;
        for (size_t d = 0; d < x->dests.size(); d++) {
            // addresses inside the DESTINATION's memory: always the full (receiver) layout, whatever this rank's own is
            p.blob[d] = x->dests[d].base + (size_t)s * x->slot_bytes;
            p.flag[d] = reinterpret_cast<uint32_t*>(x->dests[d].base + (size_t)x->slots * x->slot_bytes + (size_t)s * 256);
        }
    }
    NS_CUDA(cudaSetDevice(x->device));
    NS_CUDA(cudaMemcpy(x->pub(0), h.data(), h.size() * sizeof(PublishDest), cudaMemcpyHostToDevice));
    return NS_OK;
}

}  // namespace

static int exchange_create_impl(int device, uint32_t world, uint32_t rank, uint32_t max_queries, uint32_t slots,
                                bool publisher_only, ns_exchange** out);

static int ns_exchange_create_impl(int device, uint32_t world, uint32_t rank, uint32_t max_queries, uint32_t slots,
                                  ns_exchange** out) {
    return exchange_create_impl(device, world, rank, max_queries, slots, false, out);
}

extern "C" int ns_exchange_create(int device, uint32_t world, uint32_t rank, uint32_t max_queries, uint32_t slots,
                                  ns_exchange** out) {
    return abi_guard("ns_exchange_create", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_exchange_create_impl(device, world, rank, max_queries, slots, out); });
}

int nsb::exchange_create_publisher(int device, uint32_t world, uint32_t rank, uint32_t max_queries, uint32_t slots,
                                   ns_exchange** out) {
    return exchange_create_impl(device, world, rank, max_queries, slots, true, out);
}

static int exchange_create_impl(int device, uint32_t world, uint32_t rank, uint32_t max_queries, uint32_t slots,
                                bool publisher_only, ns_exchange** out) {
    if (!out || world == 0 || world > NS_MAX_PEERS || rank >= world || slots == 0 || slots > 8 || max_queries == 0) {
        set_error("ns_exchange_create: bad argument (world <= NS_MAX_PEERS, rank < world, 1 <= slots <= 8)");
        return NS_ERR_INVALID;
    }
    *out = nullptr;
    NS_CUDA(cudaSetDevice(device));
    auto x = std::make_unique<ns_exchange>();
    x->device = device;
    x->world = world;
    x->rank = rank;
    x->slots = slots;
    x->max_q = max_queries;
    x->stride = blob_capacity(max_queries);
    x->slot_bytes = (size_t)world * x->stride;
    x->publisher_only = publisher_only;
    if (publisher_only) {
        // a rank that only publishes needs neither gather regions nor a merged blob nor a pinned result buffer:
        // its device memory holds the flags / status words (unused) and the PublishDest table
        x->off_flags = 0;
        x->off_status = (size_t)slots * 256;
        x->off_merged = x->off_status + (size_t)slots * 256;
        x->off_pub = x->off_merged;
    } else {
        x->off_flags = (size_t)slots * x->slot_bytes;
        x->off_status = x->off_flags + (size_t)slots * 256;
        x->off_merged = x->off_status + (size_t)slots * 256;
        x->off_pub = x->off_merged + (size_t)slots * x->stride;
    }
    x->total = x->off_pub + align_up((size_t)slots * sizeof(PublishDest));
    if (const char* s = std::getenv("NSB200_EXCHANGE_TIMEOUT_MS")) x->timeout_ns = (uint64_t)std::max(1L, std::atol(s)) * 1000000ull;
    cudaError_t e = cudaMalloc(&x->d_mem, x->total);
    // flags and status start at 0; epochs start at 1
    if (e == cudaSuccess) e = cudaMemset(x->d_mem + x->off_flags, 0, x->off_merged - x->off_flags);
    if (e == cudaSuccess && !publisher_only) e = cudaMallocHost(&x->h_out, x->stride + 256);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&x->stream, cudaStreamNonBlocking);
    for (uint32_t s = 0; e == cudaSuccess && s < slots; s++) e = cudaEventCreateWithFlags(&x->ev_done[s], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&x->ev_copy, cudaEventDisableTiming | cudaEventBlockingSync);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        set_error(std::string("ns_exchange_create: ") + cudaGetErrorString(e));
        ns_exchange_destroy(x.release());
        return NS_ERR_CUDA;
    }
    *out = x.release();
    return NS_OK;
}

extern "C" void ns_exchange_destroy(ns_exchange* x) {
    if (!x) return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (auto& d : x->dests)
        if (d.ipc && d.base) cudaIpcCloseMemHandle(d.base);
    if (x->d_mem) cudaFree(x->d_mem);
    if (x->h_out) cudaFreeHost(x->h_out);
    for (auto& ev : x->ev_done)
        if (ev) cudaEventDestroy(ev);
    if (x->ev_copy) cudaEventDestroy(x->ev_copy);
    if (x->stream) cudaStreamDestroy(x->stream);
    delete x;
}

extern "C" int ns_exchange_ipc_handle(ns_exchange* x, void* handle) {
    if (!x || !handle || x->publisher_only) { set_error("ns_exchange_ipc_handle: null or publisher-only exchange"); return NS_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == NS_IPC_HANDLE_BYTES, "handle size");
    NS_CUDA(cudaSetDevice(x->device));
    cudaIpcMemHandle_t h;
    NS_CUDA(cudaIpcGetMemHandle(&h, x->d_mem));
    std::memcpy(handle, &h, sizeof(h));
    return NS_OK;
}

static int add_dest(ns_exchange* x, uint32_t peer_rank, uint8_t* base, bool ipc) {
    std::lock_guard<std::mutex> lk(x->mu);
    if (x->dests.size() >= (size_t)kMaxPeers) { set_error("ns_exchange: too many destinations"); return NS_ERR_INVALID; }
    for (auto& d : x->dests)
        if (d.rank == peer_rank) { set_error("ns_exchange: destination attached twice"); return NS_ERR_INVALID; }
    x->dests.push_back(ns_exchange::Dest{peer_rank, base, ipc});
    if (peer_rank == x->rank) x->receiver = true;
    return upload_pub(x);
}

extern "C" int ns_exchange_attach_ipc(ns_exchange* x, uint32_t peer_rank, const void* handle) {
    if (!x || !handle || peer_rank >= x->world || peer_rank == x->rank) { set_error("ns_exchange_attach_ipc: bad argument"); return NS_ERR_INVALID; }
    NS_CUDA(cudaSetDevice(x->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    NS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    return add_dest(x, peer_rank, static_cast<uint8_t*>(p), true);
}

extern "C" int ns_exchange_attach_local(ns_exchange* x, ns_exchange* peer) {
    if (!x || !peer) { set_error("ns_exchange_attach_local: null"); return NS_ERR_INVALID; }
    if (peer->publisher_only) { set_error("ns_exchange_attach_local: the destination was created as a publisher only"); return NS_ERR_INVALID; }
    if (peer->world != x->world || peer->stride != x->stride || peer->slots != x->slots) {
        set_error("ns_exchange_attach_local: the two exchanges were created with different shapes");
        return NS_ERR_INVALID;
    }
    NS_CUDA(cudaSetDevice(x->device));
    if (peer->device != x->device) {
        int can = 0;
        NS_CUDA(cudaDeviceCanAccessPeer(&can, x->device, peer->device));
        if (!can) { set_error("ns_exchange_attach_local: no peer access between the two devices"); return NS_ERR_CUDA; }
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) { set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); return NS_ERR_CUDA; }
    }
    return add_dest(x, peer->rank, peer->d_mem, false);
}

extern "C" int ns_batch_launch_exchange(ns_batch* b, ns_exchange* x, uint64_t step, void* stream) {
    if (!b || !x) { set_error("ns_batch_launch_exchange: null"); return NS_ERR_INVALID; }
    if (b->st->device != x->device) { set_error("ns_batch_launch_exchange: batch and exchange live on different devices"); return NS_ERR_INVALID; }
    if (b->out_bytes > x->stride) { set_error("ns_batch_launch_exchange: batch larger than the exchange was created for"); return NS_ERR_INVALID; }
    if (x->dests.empty()) { set_error("ns_batch_launch_exchange: no destination attached"); return NS_ERR_STATE; }
    NS_CUDA(cudaSetDevice(x->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : b->res->stream;
    const uint32_t slot = (uint32_t)(step % x->slots);
    const uint32_t epoch = (uint32_t)(step / x->slots) + 1u;
    // Q == 0: nothing to score, nothing to publish (the merge of an empty batch is a no-op as well)
    int rc = launch_score(b, s, b->Q ? x->pub(slot) : nullptr, epoch);
    if (rc != NS_OK) return rc;
    NS_CUDA(cudaEventRecord(b->res->ev[2], s));
    b->launched = true;
    return NS_OK;
}

extern "C" int ns_exchange_merge(ns_exchange* x, uint64_t step, uint32_t Q, int k_in, int spin, void* stream) {
    if (!x) { set_error("ns_exchange_merge: null"); return NS_ERR_INVALID; }
    if (!x->receiver) { set_error("ns_exchange_merge: this rank does not receive (it is not among its own destinations)"); return NS_ERR_STATE; }
    const uint32_t k = (uint32_t)std::max(1, std::min(k_in, NS_MAX_K));
    const size_t off_n = align_up(std::max<size_t>(1, (size_t)Q * k) * sizeof(ns_hit));
    const size_t off_found = off_n + align_up(std::max<size_t>(1, Q) * 4);
    if (off_found + align_up(std::max<size_t>(1, Q) * 8) > x->stride) { set_error("ns_exchange_merge: Q, k larger than the exchange was created for"); return NS_ERR_INVALID; }
    NS_CUDA(cudaSetDevice(x->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : x->stream;
    const uint32_t slot = (uint32_t)(step % x->slots);
    const uint32_t epoch = (uint32_t)(step / x->slots) + 1u;
    if (Q) {
        if (spin) {
            exchange_wait_kernel<<<1, 32, 0, s>>>(x->flags(x->d_mem, slot), x->world, epoch, x->timeout_ns, x->status(slot));
            NS_CUDA(cudaGetLastError());
        }
        uint8_t* g = x->gather(x->d_mem, slot);
        uint8_t* m = x->merged(slot);
        int rc = merge_launch(x->device, Q, (int)k, x->world, g, g + off_n, g + off_found, x->stride, x->stride, x->stride, m,
                              m + off_n, m + off_found, s);
        if (rc != NS_OK) return rc;
    }
    NS_CUDA(cudaEventRecord(x->ev_done[slot], s));
    return NS_OK;
}

int nsb::exchange_root_merge(ns_exchange* root, ns_batch* const* batches, int ndev, uint64_t step, uint32_t Q, int k) {
    if (!root || !batches || ndev < 1 || !batches[0]) { set_error("exchange_root_merge: null"); return NS_ERR_INVALID; }
    NS_CUDA(cudaSetDevice(root->device));
    cudaStream_t s = batches[0]->res->stream;
    for (int d = 1; d < ndev; d++) {
        if (!batches[d] || !batches[d]->launched) { set_error("exchange_root_merge: a device's batch was not launched"); return NS_ERR_STATE; }
        NS_CUDA(cudaStreamWaitEvent(s, batches[d]->res->ev[2], 0));  // recorded right after that device's score kernel
    }
    return ns_exchange_merge(root, step, Q, k, 0, s);
}

extern "C" int ns_exchange_result_device(ns_exchange* x, uint64_t step, void** d_blob) {
    if (!x || !d_blob) return NS_ERR_INVALID;
    *d_blob = x->merged((uint32_t)(step % x->slots));
    return NS_OK;
}

extern "C" int ns_exchange_fetch(ns_exchange* x, uint64_t step, uint32_t Q, int k_in, ns_hit* out_hits,
                                 uint32_t* out_nhits, uint64_t* out_found) {
    if (!x) { set_error("ns_exchange_fetch: null"); return NS_ERR_INVALID; }
    if (!x->receiver) { set_error("ns_exchange_fetch: this rank does not receive (it is not among its own destinations)"); return NS_ERR_STATE; }
    const uint32_t k = (uint32_t)std::max(1, std::min(k_in, NS_MAX_K));
    const size_t sz_hits = align_up(std::max<size_t>(1, (size_t)Q * k) * sizeof(ns_hit));
    const size_t sz_n = align_up(std::max<size_t>(1, Q) * 4);
    const size_t sz_found = align_up(std::max<size_t>(1, Q) * 8);
    const size_t bytes = sz_hits + sz_n + sz_found;
    if (bytes > x->stride) { set_error("ns_exchange_fetch: Q, k larger than the exchange was created for"); return NS_ERR_INVALID; }
    NS_CUDA(cudaSetDevice(x->device));
    const uint32_t slot = (uint32_t)(step % x->slots);
    std::lock_guard<std::mutex> lk(x->mu);
    NS_CUDA(cudaStreamWaitEvent(x->stream, x->ev_done[slot], 0));
    if (Q) NS_CUDA(cudaMemcpyAsync(x->h_out, x->merged(slot), bytes, cudaMemcpyDeviceToHost, x->stream));
    NS_CUDA(cudaMemcpyAsync(x->h_out + x->stride, x->status(slot), 4, cudaMemcpyDeviceToHost, x->stream));
    NS_CUDA(cudaEventRecord(x->ev_copy, x->stream));
    NS_CUDA(cudaEventSynchronize(x->ev_copy));
    uint32_t st = 0;
    std::memcpy(&st, x->h_out + x->stride, 4);
    if (st != 0) {
        cudaMemsetAsync(x->status(slot), 0, 4, x->stream);
        cudaStreamSynchronize(x->stream);
        char buf[160];
        std::snprintf(buf, sizeof(buf), "ns_exchange_fetch: step %llu timed out waiting for the result blobs of ranks (bit mask) 0x%x",
                      (unsigned long long)step, st);
        set_error(buf);
        return NS_ERR_STATE;
    }
    if (Q) {
        if (out_hits) std::memcpy(out_hits, x->h_out, (size_t)Q * k * sizeof(ns_hit));
        if (out_nhits) std::memcpy(out_nhits, x->h_out + sz_hits, (size_t)Q * 4);
        if (out_found) std::memcpy(out_found, x->h_out + sz_hits + sz_n, (size_t)Q * 8);
    }
    return NS_OK;
}

// ---------------------------------------------------------------------------------------------
// Semantic expansion: the similarity scan on the device (semantic_kernels.cuh)
// ---------------------------------------------------------------------------------------------

struct ns_semantic {
    int device = 0;
    uint32_t rows = 0, dim = 0;
    float* d_vT = nullptr;  // [dim][rows]
    std::mutex mu;          // one scan at a time per handle (scratch below)
    float* d_q = nullptr;
    uint32_t* d_rows = nullptr;
    float* d_sims = nullptr;
    uint32_t* d_count = nullptr;
    uint32_t cap_m = 0, cap_per = 0;
    cudaStream_t stream = nullptr;
};

static int ns_semantic_upload_impl(int device, uint32_t rows, uint32_t dim, const float* vecs, ns_semantic** out) {
    if (!out || !vecs || rows == 0 || dim == 0) { set_error("ns_semantic_upload: bad argument"); return NS_ERR_INVALID; }
    *out = nullptr;
    NS_CUDA(cudaSetDevice(device));
    auto s = std::make_unique<ns_semantic>();
    s->device = device;
    s->rows = rows;
    s->dim = dim;
    std::vector<float> t((size_t)rows * dim);  // transpose on the host, once
    for (uint32_t r = 0; r < rows; r++)
        for (uint32_t i = 0; i < dim; i++) t[(size_t)i * rows + r] = vecs[(size_t)r * dim + i];
    cudaError_t e = cudaMalloc(&s->d_vT, t.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(s->d_vT, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        if (s->d_vT) cudaFree(s->d_vT);
        set_error(std::string("ns_semantic_upload: ") + cudaGetErrorString(e));
        return NS_ERR_CUDA;
    }
    *out = s.release();
    return NS_OK;
}

extern "C" int ns_semantic_upload(int device, uint32_t rows, uint32_t dim, const float* vecs, ns_semantic** out) {
    return abi_guard("ns_semantic_upload", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_semantic_upload_impl(device, rows, dim, vecs, out); });
}

extern "C" void ns_semantic_destroy(ns_semantic* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) {
        cudaStreamSynchronize(s->stream);
        cudaStreamDestroy(s->stream);
    }
    for (void* p : {(void*)s->d_vT, (void*)s->d_q, (void*)s->d_rows, (void*)s->d_sims, (void*)s->d_count})
        if (p) cudaFree(p);
    delete s;
}

static int ns_semantic_scan_impl(ns_semantic* s, uint32_t M, const float* qvecs, float min_sim, uint32_t cap,
                                uint32_t* out_rows, float* out_sims, uint32_t* out_count) {
    if (!s || !qvecs || !out_rows || !out_sims || !out_count || cap == 0) { set_error("ns_semantic_scan: bad argument"); return NS_ERR_INVALID; }
    if (M == 0) return NS_OK;
    std::lock_guard<std::mutex> lk(s->mu);
    NS_CUDA(cudaSetDevice(s->device));
    if (s->cap_m < M || s->cap_per < cap) {
        for (void* p : {(void*)s->d_q, (void*)s->d_rows, (void*)s->d_sims, (void*)s->d_count})
            if (p) cudaFree(p);
        s->d_q = nullptr; s->d_rows = nullptr; s->d_sims = nullptr; s->d_count = nullptr;
        const uint32_t cm = std::max(M, s->cap_m), cp = std::max(cap, s->cap_per);
        NS_CUDA(cudaMalloc(&s->d_q, (size_t)cm * s->dim * sizeof(float)));
        NS_CUDA(cudaMalloc(&s->d_rows, (size_t)cm * cp * 4));
        NS_CUDA(cudaMalloc(&s->d_sims, (size_t)cm * cp * 4));
        NS_CUDA(cudaMalloc(&s->d_count, (size_t)cm * 4));
        s->cap_m = cm;
        s->cap_per = cp;
    }
    const uint32_t cp = s->cap_per;
    NS_CUDA(cudaMemcpyAsync(s->d_q, qvecs, (size_t)M * s->dim * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    NS_CUDA(cudaMemsetAsync(s->d_count, 0, (size_t)M * 4, s->stream));
    const size_t smem = (size_t)kSemChunk * s->dim * sizeof(float);
    if (smem > 48 * 1024) { set_error("ns_semantic_scan: embedding dimension too large for the query tile"); return NS_ERR_INVALID; }
    cosine_scan_kernel<<<(s->rows + 255) / 256, 256, smem, s->stream>>>(s->d_vT, s->rows, s->dim, s->d_q, M, min_sim, cp, s->d_rows,
                                                                        s->d_sims, s->d_count);
    NS_CUDA(cudaGetLastError());
    NS_CUDA(cudaMemcpyAsync(out_count, s->d_count, (size_t)M * 4, cudaMemcpyDeviceToHost, s->stream));
    NS_CUDA(cudaStreamSynchronize(s->stream));
    for (uint32_t m = 0; m < M; m++) {
        const uint32_t n = std::min(out_count[m], cap);
        if (n == 0) continue;
        NS_CUDA(cudaMemcpyAsync(out_rows + (size_t)m * cap, s->d_rows + (size_t)m * cp, (size_t)n * 4, cudaMemcpyDeviceToHost, s->stream));
        NS_CUDA(cudaMemcpyAsync(out_sims + (size_t)m * cap, s->d_sims + (size_t)m * cp, (size_t)n * 4, cudaMemcpyDeviceToHost, s->stream));
    }
    NS_CUDA(cudaStreamSynchronize(s->stream));
    return NS_OK;
}

extern "C" int ns_semantic_scan(ns_semantic* s, uint32_t M, const float* qvecs, float min_sim, uint32_t cap,
                                uint32_t* out_rows, float* out_sims, uint32_t* out_count) {
    return abi_guard("ns_semantic_scan", NS_ERR_NOMEM, NS_ERR_STATE, [&]() -> int { return ns_semantic_scan_impl(s, M, qvecs, min_sim, cap, out_rows, out_sims, out_count); });
}

// Debug build only (make debug -> libnsb200_dbg.so): the per-class counts of index checks that failed inside the
// score kernel on `device` since the library was loaded (DbgClass in bm25_kernels.cuh).
extern "C" int ns_debug_violations(int device, uint64_t* counts, int n) {
    if (!counts || n < 0) { set_error("ns_debug_violations: bad argument"); return NS_ERR_INVALID; }
#ifdef NSB_DEBUG_CHECKS
    NS_CUDA(cudaSetDevice(device));
    NS_CUDA(cudaDeviceSynchronize());
    unsigned long long h[kDbgClasses];
    NS_CUDA(cudaMemcpyFromSymbol(h, g_dbg_violations, sizeof(h)));
    for (int i = 0; i < n; i++) counts[i] = i < (int)kDbgClasses ? h[i] : 0;
    return NS_OK;
#else
    (void)device;
    set_error("ns_debug_violations: not a debug build (make -C nextsearch-api_b200/csrc debug builds libnsb200_dbg.so)");
    return NS_ERR_STATE;
#endif
}

// Debug build only: one thread fails one check of the last class on purpose, so that a reader of
// ns_debug_violations knows the counters are live (they must then show exactly one more "list" violation).
extern "C" int ns_debug_selftest(int device) {
#ifdef NSB_DEBUG_CHECKS
    NS_CUDA(cudaSetDevice(device));
    dbg_selftest_kernel<<<1, 1>>>(0);
    NS_CUDA(cudaGetLastError());
    NS_CUDA(cudaDeviceSynchronize());
    return NS_OK;
#else
    (void)device;
    set_error("ns_debug_selftest: not a debug build");
    return NS_ERR_STATE;
#endif
}

extern "C" int ns_selftest_fastdiv(int device, uint64_t n, uint64_t seed, uint64_t* mismatches) {
    if (!mismatches) return NS_ERR_INVALID;
    NS_CUDA(cudaSetDevice(device));
    unsigned long long* d = nullptr;
    NS_CUDA(cudaMalloc(&d, 8));
    NS_CUDA(cudaMemset(d, 0, 8));
    selftest_fastdiv_kernel<<<148 * 8, 256>>>(n, seed, d);
    cudaError_t e = cudaGetLastError();
    unsigned long long h = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) { set_error(std::string("ns_selftest_fastdiv: ") + cudaGetErrorString(e)); return NS_ERR_CUDA; }
    *mismatches = h;
    return NS_OK;
}
