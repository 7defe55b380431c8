/*
 * nextsearch_b200.h — C ABI of the B200-native BM25 scoring + top-k path.
 *
 * This is the drop-in boundary for NextSearch's query-time hot path
 * (reference: src/api_engine.cpp:426-505, "for segId … score[docId] += … / top-K heap").
 * The reference has no FFI; the seam it offers is the pair of Engine members
 *   bool  Engine::reload()                      (include/api_engine.hpp:65, src/api_engine.cpp:50-162)
 *   json  Engine::search(const string&, int k)  (include/api_engine.hpp:66, src/api_engine.cpp:369-542)
 * and, one level down, the inner loop over (segment, query term, posting).
 * The entry points below are what a cgo/JNI/ctypes/C++ binding for that seam
 * would bind.  Plain pointers and sizes only; no C++ / torch types; no
 * exceptions cross this boundary; every call returns an ns_status.
 *
 * Two layers:
 *   ns_index_* / ns_batch_* / ns_search_batch / ns_merge_*   — the device path
 *       proper: segments become device-resident CSR arrays, a batch of
 *       lexicon-resolved queries is scored and top-k selected on the GPU.
 *   ns_engine_*  — host mirror of cord19::Engine (reads the real on-disk
 *       format, tokenises, resolves the lexicon, computes IDF with the host's
 *       logf, calls the device path, packages the reference's JSON fields).
 *   ns_corpus_* / ns_text_*  — the synthetic-corpus generator / segment writer
 *       (replaces include/segment_writer.hpp:23-169 for large corpora) and the
 *       tokenizer (include/textutil.hpp:13-37), exposed for tests and bench.
 *
 * Tie-break (total order, stated in DESIGN.md): score desc, then global
 * segment index asc, then docId asc.  Scores are the exact f32 values the
 * reference computes (same operation tree, round-to-nearest, no FMA).
 */
#ifndef NEXTSEARCH_B200_H
#define NEXTSEARCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ns_status {
    NS_OK = 0,
    NS_ERR_INVALID = 1,      /* bad argument (null pointer, k out of range after clamp, too many terms …) */
    NS_ERR_CUDA = 2,         /* CUDA runtime error or no usable device; ns_last_error() has the text */
    NS_ERR_IO = 3,           /* missing / truncated index file (reference: reload() returns false) */
    NS_ERR_FORMAT = 4,       /* posting list not strictly increasing in docId, docId >= N, offset misaligned */
    NS_ERR_NOMEM = 5,
    NS_ERR_STATE = 6         /* e.g. search before commit */
} ns_status;

#define NS_MAX_K 100          /* reference clamps k to 1..100: src/api_engine.cpp:377 */
#define NS_MAX_TERMS 256      /* per (query, segment); the reference has no cap (its expansion stops at 40: src/api_engine.cpp:417) */
#define NS_BARREL_COUNT 64    /* include/barrels.hpp:12 */

/* thread-local text of the last error raised on this thread */
const char* ns_last_error(void);

/* ------------------------------------------------------------------ */
/* Device index (one GPU).  Replaces Segment{docs, lex, inv_barrels}   */
/* (include/api_types.hpp:46-60) for the scoring loop.                 */
/* ------------------------------------------------------------------ */
typedef struct ns_index ns_index;

/* One weighted query term resolved against one segment's lexicon
 * (reference: the (term, qweight) pair of src/api_engine.cpp:449-461 after
 * seg.lex.find(term) and bm25_idf(seg.N, e.df)). */
typedef struct ns_qterm {
    uint32_t seg;      /* global segment index (position in seg_names) */
    uint32_t row;      /* CSR row of the term inside that segment, as uploaded */
    float    idf;      /* bm25_idf(N, df), computed on the host (src/api_engine.cpp:45-47) */
    float    weight;   /* qweight; 1.0f when semantic expansion is off (src/api_engine.cpp:420) */
} ns_qterm;

/* Hit{s, segId, docId} of src/api_engine.cpp:427-431 */
typedef struct ns_hit {
    float    score;
    uint32_t seg;
    uint32_t doc;
} ns_hit;

int ns_device_count(void);

int ns_index_create(int device, ns_index** out);
void ns_index_destroy(ns_index* idx);

/* Stage one segment for the next commit.
 *   doc_len[N]          DocInfo.doc_len (include/api_types.hpp:18-21)
 *   term_begin[T]       first posting of row t, in postings (not bytes)
 *   term_count[T]       LexEntry.count (include/api_types.hpp:23-29)
 *   postings[P]         interleaved little-endian {u32 docId, u32 tf} exactly as in
 *                       inverted_bNNN.bin (include/segment_writer.hpp:151-155)
 * Validates: docId < N and strictly increasing inside every row. */
int ns_index_add_segment(ns_index* idx, uint32_t global_seg, uint32_t N, float avgdl,
                         const uint32_t* doc_len, uint32_t T, const uint64_t* term_begin,
                         const uint32_t* term_count, const void* postings, uint64_t P);

/* Same with two options.
 *   row_idf[T]   (may be NULL) the idf every row's resident term scores are built with; NULL means
 *                bm25_idf(N, term_count[t]).  The host engine passes bm25_idf(N, LexEntry.df) of its lexicon
 *                (src/api_engine.cpp:461), so its queries never need the per-batch pre-pass even if a writer
 *                ever stored df != count.
 *   flags        NS_SEG_DROP_RAW: keep only {docId, resident term score} per posting (built in place); halves
 *                the segment's footprint.  A batch that names a row with an idf different from the resident
 *                one is then refused with NS_ERR_STATE.
 * Rows that SHARE postings (two lexicon entries pointing at overlapping ranges — legal for the reference, whose
 * loop only follows (offset, count), src/api_engine.cpp:469-476) cannot have one resident score per posting:
 * such a segment gets no resident scores and keeps its raw postings whatever the flags say; its terms are scored
 * through the per-batch pre-pass, each row under its own idf. */
#define NS_SEG_DROP_RAW 1u
int ns_index_add_segment_ex(ns_index* idx, uint32_t global_seg, uint32_t N, float avgdl,
                            const uint32_t* doc_len, uint32_t T, const uint64_t* term_begin,
                            const uint32_t* term_count, const float* row_idf, const void* postings, uint64_t P,
                            uint32_t flags);

/* Streamed form of the same upload (reference: load_segment_barrels, src/api_segment.cpp:69-103, reads 64
 * inverted_bNNN.bin files): begin allocates the device posting array and a PINNED host buffer of P postings;
 * the loader reads each barrel file straight into its place in ns_upload_buffer() and calls ns_upload_push for
 * the finished range (thread-safe, asynchronous: the copies overlap the remaining file reads); finish waits
 * for the copies and does what ns_index_add_segment_ex does.  finish and abort both free the ticket. */
typedef struct ns_upload ns_upload;
int ns_upload_begin(ns_index* idx, uint64_t P, ns_upload** out);
void* ns_upload_buffer(ns_upload* u);
int ns_upload_push(ns_upload* u, uint64_t first_posting, uint64_t count);
int ns_upload_finish(ns_upload* u, uint32_t global_seg, uint32_t N, float avgdl, const uint32_t* doc_len,
                     uint32_t T, const uint64_t* term_begin, const uint32_t* term_count, const float* row_idf,
                     uint32_t flags);
void ns_upload_abort(ns_upload* u);

/* Atomically replace the searchable index with the staged segments
 * (reference: `segments = std::move(loaded)` src/api_engine.cpp:90).  In-flight
 * batches keep the index they started with.  On failure the previous index stays. */
int ns_index_commit(ns_index* idx);
/* Drop staged, uncommitted segments. */
int ns_index_abort(ns_index* idx);

int ns_index_num_segments(const ns_index* idx);
uint64_t ns_index_device_bytes(const ns_index* idx);

/* ------------------------------------------------------------------ */
/* Batched search.                                                     */
/* ------------------------------------------------------------------ */

/* One-shot, HOST buffers in and out (the e2e path; H2D + kernels + D2H inside).
 *   q_off[Q+1]   prefix offsets into terms[]; terms of one query are ordered by
 *                (seg asc, then query-term order); duplicates are kept
 *                (src/api_engine.cpp:391-397).
 *   k            clamped to 1..100 like src/api_engine.cpp:377
 *   out_hits     [Q][k_clamped] best first;  out_nhits[Q];  out_found[Q]
 *                (found = Σ_segments |{docs with ≥1 matching posting}|, src/api_engine.cpp:495)
 * Thread-safe: may be called concurrently from many host threads. */
int ns_search_batch(ns_index* idx, uint32_t Q, int k, const uint64_t* q_off, const ns_qterm* terms,
                    ns_hit* out_hits, uint32_t* out_nhits, uint64_t* out_found);

/* Split form of the same call: prepare uploads the descriptors (H2D) and
 * allocates device outputs; launch enqueues the kernels on `stream`
 * (a cudaStream_t passed as void*, NULL = the batch's own stream); fetch copies
 * results back and synchronises.  bench.py's device-resident `value` times
 * ns_batch_launch alone. */
typedef struct ns_batch ns_batch;
int ns_batch_prepare(ns_index* idx, uint32_t Q, int k, const uint64_t* q_off, const ns_qterm* terms,
                     ns_batch** out);
int ns_batch_launch(ns_batch* b, void* stream);
int ns_batch_sync(ns_batch* b);
int ns_batch_fetch(ns_batch* b, ns_hit* out_hits, uint32_t* out_nhits, uint64_t* out_found);
void ns_batch_destroy(ns_batch* b);
/* device pointers of a batch's results ([Q][k] ns_hit, [Q] u32, [Q] u64) for
 * torch.distributed all_gather without a host round trip */
int ns_batch_device_results(ns_batch* b, void** d_hits, void** d_nhits, void** d_found);
/* introspection for bench.py */
uint64_t ns_batch_posting_count(const ns_batch* b);   /* Σ LexEntry.count over all (query, term, segment) */
uint32_t ns_batch_num_launches(const ns_batch* b);    /* kernels one ns_batch_launch enqueues */
float ns_batch_last_kernel_ms(ns_batch* b, int which); /* CUDA-event time of the last launch: 0 = zeroing + score+topk, 1 = what follows (merge) */
void* ns_batch_stream(ns_batch* b);                   /* the batch's own cudaStream_t (what a NULL stream argument means) */
uint64_t ns_batch_upload_bytes(const ns_batch* b);    /* host->device bytes prepare copied for this batch */
uint64_t ns_batch_result_bytes(const ns_batch* b);    /* device->host bytes one fetch copies */
int ns_batch_set_splits(ns_batch* b, uint32_t splits); /* 0 = auto */

/* Merge `nlists` per-rank (or per-split) result sets on the device.
 *   d_hits   [nlists][Q][k] ns_hit, each list best-first; d_nhits [nlists][Q]; d_found [nlists][Q]
 *   outputs  [Q][k], [Q], [Q]  (device pointers)
 * Order: score desc, seg asc, doc asc. */
int ns_merge_device(int device, uint32_t Q, int k, uint32_t nlists, const void* d_hits,
                    const void* d_nhits, const void* d_found, void* d_out_hits, void* d_out_nhits,
                    void* d_out_found, void* stream);

/* The batch's whole result as one contiguous device blob (hits | nhits | found at the returned
 * byte offsets): the unit ranks exchange with ONE all-gather per batch. */
int ns_batch_result_blob(ns_batch* b, void** d_blob, uint64_t* bytes, uint64_t* off_nhits, uint64_t* off_found);
/* Merge `nlists` such blobs laid out blob_stride bytes apart (the all-gather output). */
int ns_merge_blobs_device(int device, uint32_t Q, int k, uint32_t nlists, const void* d_blobs,
                          uint64_t blob_stride, uint64_t off_nhits, uint64_t off_found, void* d_out_hits,
                          void* d_out_nhits, void* d_out_found, void* stream);

/* ------------------------------------------------------------------ */
/* Multi-GPU: peer exchange of the per-GPU result blobs.               */
/* (reference: the cross-segment state of src/api_engine.cpp:435,495 — */
/*  the global top-K heap and the found sum — is all that is exchanged)*/
/* ------------------------------------------------------------------ */
/* Segments shard across GPUs; every GPU scores the same query batch against its own segments.  The score
 * kernel itself stores each finished query's list into the gather buffer of every destination GPU (P2P stores
 * over NVLink, no separate collective), raises a flag when its blob is complete, and a receiving GPU merges
 * the `world` blobs under the total order once all flags of the step are up.
 *   same process  : ns_exchange_attach_local (cudaDeviceEnablePeerAccess)
 *   one process per GPU (torchrun): exchange the 64-byte handle of ns_exchange_ipc_handle out of band
 *                   (e.g. torch.distributed.all_gather_object) and ns_exchange_attach_ipc it.
 * A rank that attaches ITSELF (attach_local(x, x)) is a receiver: it waits and merges.  Every rank must use the
 * same (world, max_queries, slots).  Steps are numbered by the caller, consecutive steps use consecutive
 * slots; with launches and merges of one rank on ONE stream two slots suffice (a rank can be at most one
 * step ahead of the slowest one, because its own merge of step i waits for everybody's step i). */
typedef struct ns_exchange ns_exchange;
#define NS_IPC_HANDLE_BYTES 64
#define NS_MAX_PEERS 16
int ns_exchange_create(int device, uint32_t world, uint32_t rank, uint32_t max_queries, uint32_t slots,
                       ns_exchange** out);
void ns_exchange_destroy(ns_exchange* x);
int ns_exchange_ipc_handle(ns_exchange* x, void* handle /* NS_IPC_HANDLE_BYTES */);
int ns_exchange_attach_ipc(ns_exchange* x, uint32_t peer_rank, const void* handle);
int ns_exchange_attach_local(ns_exchange* x, ns_exchange* peer);
/* ns_batch_launch whose score kernel also publishes to every attached destination. */
int ns_batch_launch_exchange(ns_batch* b, ns_exchange* x, uint64_t step, void* stream);
/* Receiver side of `step`, enqueued on `stream` (NULL = the exchange's own): wait until every rank's blob of
 * the step has arrived, then merge the `world` blobs into the exchange's merged blob of slot step % slots
 * (same layout as ns_batch_result_blob).
 *   spin = 1: a one-warp kernel polls the arrival flags (bounded by the timeout) — the cross-process case,
 *             where nothing else can order this stream after another process's kernel.  Enqueue it AFTER this
 *             rank's own ns_batch_launch_exchange of the step.
 *   spin = 0: the caller has ordered `stream` after every publisher's score kernel itself
 *             (cudaStreamWaitEvent on events of the other devices, same process): no polling. */
int ns_exchange_merge(ns_exchange* x, uint64_t step, uint32_t Q, int k, int spin, void* stream);
int ns_exchange_result_device(ns_exchange* x, uint64_t step, void** d_blob);
/* D2H of the merged result of `step` (waits for it).  NS_ERR_STATE if a rank's blob did not arrive within
 * NSB200_EXCHANGE_TIMEOUT_MS (default 10000): the wait is bounded, a missing peer cannot hang the GPU. */
int ns_exchange_fetch(ns_exchange* x, uint64_t step, uint32_t Q, int k, ns_hit* out_hits, uint32_t* out_nhits,
                      uint64_t* out_found);

/* ------------------------------------------------------------------ */
/* Engine mirror (host).  Same surface as cord19::Engine for this path.*/
/* ------------------------------------------------------------------ */
typedef struct ns_engine ns_engine;

/* device < 0: host-only engine (lexicon / tokeniser / resolve available, search fails loudly) */
int ns_engine_create(const char* index_dir, int device, ns_engine** out);
/* One engine spanning `ndev` GPUs of the box (ns_index_create(ndev, dev_ids)): segment j of the
 * engine's share lives on devices[j % ndev]; a batch is tokenised and resolved once, scored on every device,
 * the per-device results are stored into devices[0]'s gather buffer by the score kernels (peer memory) and
 * merged there.  This is what a C++ api_server links: one process, one handle, all GPUs.  ndev = 0: host-only. */
int ns_engine_create_multi(const char* index_dir, int ndev, const int* devices, ns_engine** out);
int ns_engine_num_devices(const ns_engine* e);
void ns_engine_destroy(ns_engine* e);

/* Shard selection for multi-GPU: this engine uploads only segments with
 * (global_seg % world) == rank; set before reload.  Default world=1. */
int ns_engine_set_shard(ns_engine* e, int rank, int world);

/* Engine::reload (src/api_engine.cpp:50-90): manifest.bin or sorted scan of
 * segments/seg_*, load every segment, upload, commit.  NS_ERR_IO if any file is missing. */
int ns_engine_reload(ns_engine* e);

int ns_engine_num_segments(const ns_engine* e);
/* copies seg_names[i] into buf; returns length or -1 */
int ns_engine_segment_name(const ns_engine* e, int i, char* buf, size_t cap);
int ns_engine_segment_stats(const ns_engine* e, int i, uint32_t* N, float* avgdl, uint32_t* T, uint64_t* P);
/* df / count of `term` in segment i; returns 0 and sets *df=*count=0 if absent */
int ns_engine_term_stats(const ns_engine* e, int i, const char* term, uint32_t* df, uint32_t* count);

/* Engine::search (src/api_engine.cpp:369-542) minus the LRU cache: returns the reference's JSON object as
 * text, byte-identical to nlohmann's dump() of it (keys: query, k, segments, found,
 * results[{score, segment, docId, cord_uid} + title, url, publish_time, author when INDEX_DIR/metadata.csv
 * has a row for the cord_uid, src/api_engine.cpp:516-532]).  NS_ERR_INVALID when the query or a result field
 * is not valid UTF-8 (dump() throws there).
 * Writes at most cap-1 bytes + NUL; returns the full length needed (like snprintf)
 * in *needed. */
int ns_engine_search_json(ns_engine* e, const char* query, int k, char* buf, size_t cap, size_t* needed);

/* Batched form: `queries` is Q NUL-terminated strings.  has_found[q] = 0 when the
 * reference would omit "found" (no usable terms: src/api_engine.cpp:407). */
int ns_engine_search_batch(ns_engine* e, uint32_t Q, const char* const* queries, int k,
                           ns_hit* out_hits, uint32_t* out_nhits, uint64_t* out_found,
                           uint8_t* has_found);

/* Same, with the Q query strings packed back to back, each NUL-terminated, in one buffer of
 * nbytes bytes (what a request-coalescing front end accumulates; avoids a pointer array). */
int ns_engine_search_batch_packed(ns_engine* e, uint32_t Q, const char* zqueries, size_t nbytes, int k,
                                  ns_hit* out_hits, uint32_t* out_nhits, uint64_t* out_found,
                                  uint8_t* has_found);

/* Front end only: tokenise + filter + lexicon lookup + IDF for a batch, producing
 * the arrays ns_search_batch takes.  Two-call protocol: pass terms=NULL to get
 * the count in *n_terms.  Only segments owned by this engine's shard are emitted. */
int ns_engine_resolve_batch(ns_engine* e, uint32_t Q, const char* const* queries,
                            uint64_t* q_off, ns_qterm* terms, uint64_t terms_cap, uint64_t* n_terms,
                            uint8_t* has_terms);
/* Same with the query strings packed back to back (each NUL-terminated) in one buffer. */
int ns_engine_resolve_batch_packed(ns_engine* e, uint32_t Q, const char* zqueries, size_t nbytes,
                                   uint64_t* q_off, ns_qterm* terms, uint64_t terms_cap, uint64_t* n_terms,
                                   uint8_t* has_terms);
/* Front end + prepare without the launch, on a one-device engine: tokenise, dictionary, descriptors, H2D.  The caller
 * launches (ns_batch_launch / ns_batch_launch_exchange), fetches and destroys the batch.  has_found as above. */
int ns_engine_prepare_batch_packed(ns_engine* e, uint32_t Q, const char* zqueries, size_t nbytes, int k, ns_batch** out,
                                   uint8_t* has_found);
ns_index* ns_engine_index(ns_engine* e);                    /* device slot 0 */
ns_index* ns_engine_device_index(ns_engine* e, int slot);
/* timing and size of the last successful reload: total seconds, seconds in barrel reads + upload (overlapped),
 * seconds building the term dictionary, posting bytes on disk (8 B x sum of LexEntry.count), device bytes */
int ns_engine_reload_stats(const ns_engine* e, double* total_s, double* read_upload_s, double* dict_s,
                           uint64_t* posting_bytes, uint64_t* device_bytes);

/* Explicit (term, qweight) lists — the reference's qterms_w (src/api_engine.cpp:410-421) — one per query:
 * t_off[Q+1] indexes terms[] / weights[].  No tokenisation, filter or expansion: what the scoring loop
 * (:426-505) receives.  A query with an empty list has no "found" (:424). */
int ns_engine_search_terms_batch(ns_engine* e, uint32_t Q, const uint64_t* t_off, const char* const* terms,
                                 const float* weights, int k, ns_hit* out_hits, uint32_t* out_nhits,
                                 uint64_t* out_found, uint8_t* has_found);
/* SemanticIndex::expand for one query (src/semantic_embedding.cpp:148-229 with the constants of
 * src/api_engine.cpp:412-417): terms NUL-separated into buf, weights into weights[wcap]; returns the count
 * (0 with *enabled = 0 when no embeddings file was found at reload), -1 if a buffer is too small. */
int ns_engine_expand(ns_engine* e, const char* query, char* buf, size_t cap, float* weights, int wcap, int* enabled);

/* One query, blocking — the shape of the reference's engine.search(q, k) call per HTTP request
 * (src/api_server.cpp:117-178).  out_hits has room for clamp(k) entries.  With the coalescer running,
 * concurrent callers are gathered into one GPU batch: a dispatcher takes up to max_batch queued requests,
 * or whatever has arrived max_wait_us after the oldest one, and scores them together (the reference
 * serialises them on Engine::mtx, src/api_engine.cpp:372).  ns_engine_search_json uses the same path. */
int ns_engine_search_one(ns_engine* e, const char* query, int k, ns_hit* out_hits, uint32_t* out_nhits,
                         uint64_t* out_found, uint8_t* has_found);
int ns_engine_coalescer_start(ns_engine* e, uint32_t max_batch, uint32_t max_wait_us, int dispatchers);
int ns_engine_coalescer_stop(ns_engine* e);
int ns_engine_coalescer_stats(ns_engine* e, uint64_t* batches, uint64_t* queries, uint64_t* max_batch_seen);
/* Load generator (bench tooling): nthreads host threads x per_thread blocking ns_engine_search_one calls over the Q
 * NUL-separated queries in zqueries; reports queries/s and the p50 / p99 latency of one call in microseconds. */
int ns_engine_load_test(ns_engine* e, uint32_t nthreads, uint32_t per_thread, uint32_t Q, const char* zqueries,
                        size_t nbytes, int k, double* qps, double* p50_us, double* p99_us);
/* cord_uid of (segment, doc); returns length or -1 */
int ns_engine_cord_uid(const ns_engine* e, uint32_t seg, uint32_t doc, char* buf, size_t cap);

/* ------------------------------------------------------------------ */
/* Semantic expansion: similarity scan on the device.                  */
/* (reference: SemanticIndex::most_similar_to_vec, the scan loop of    */
/*  src/semantic_embedding.cpp:119-127)                                */
/* ------------------------------------------------------------------ */
typedef struct ns_semantic ns_semantic;
/* vecs[rows][dim]: the L2-normalised vectors in row order (SemanticIndex::vecs, include/semantic_embedding.hpp:24). */
int ns_semantic_upload(int device, uint32_t rows, uint32_t dim, const float* vecs, ns_semantic** out);
void ns_semantic_destroy(ns_semantic* s);
/* For each of the M query vectors qvecs[M][dim]: every row whose similarity is not below min_sim, with that
 * similarity computed exactly as the reference's dot() does (sequential f32 multiply-then-add).  out_count[m] is the
 * number of such rows; the first min(out_count[m], cap) are written to out_rows / out_sims [m][cap] in NO particular
 * order (sort by row to replay the reference's scan order); out_count[m] > cap means the caller must rescan that vector
 * with a larger cap or on the host.  Banned rows are the caller's to drop. */
int ns_semantic_scan(ns_semantic* s, uint32_t M, const float* qvecs, float min_sim, uint32_t cap, uint32_t* out_rows,
                     float* out_sims, uint32_t* out_count);

/* ------------------------------------------------------------------ */
/* Text + synthetic corpus tooling.                                    */
/* ------------------------------------------------------------------ */

/* tokenize + (len<2 | stopword) filter (include/textutil.hpp:13-37,
 * src/api_engine.cpp:391-397).  Tokens are written NUL-separated into buf;
 * returns the number of kept tokens, or -1 if buf is too small. */
int ns_text_query_terms(const char* query, char* buf, size_t cap);

typedef struct ns_corpus_spec {
    uint64_t seed;        /* corpus seed (20260101) */
    uint32_t vocab;       /* V */
    double   zipf_s;      /* 1.0 */
    double   zipf_q;      /* 25.0 : p(r) ∝ 1/(r+q)^s, r = 1..V */
    uint32_t len_lo;      /* doc length L ~ U[len_lo, len_hi) */
    uint32_t len_hi;
} ns_corpus_spec;

/* Generate docs [doc_base, doc_base+ndocs) of the synthetic corpus and write them as
 * one segment directory in the reference's barrelized on-disk format
 * (include/segment_writer.hpp:65-168).  write_forward!=0 also emits forward.bin and
 * terms.bin (not read at query time).  dump_path!=NULL additionally writes the
 * documents as a flat binary dump that oracle/ref_driver feeds to the reference's own
 * SegmentWriter for byte-for-byte comparison. */
int ns_corpus_write_segment(const ns_corpus_spec* spec, uint64_t doc_base, uint32_t ndocs,
                            const char* segdir, int write_forward, const char* dump_path, int nthreads);
/* manifest.bin (src/api_segment.cpp:29-35) */
int ns_corpus_write_manifest(const char* index_dir, uint32_t nseg, const char* const* names);
/* Deterministic query strings: query i has 1..max_terms terms "t<rank>" drawn from the
 * corpus distribution.  head_ranks>0 forces the first term's rank into [1, head_ranks]
 * (the high-df stress of BASELINE configs[3]).  Output: NUL-separated strings. */
int ns_corpus_make_queries(const ns_corpus_spec* spec, uint64_t query_seed, uint32_t nq,
                           uint32_t min_terms, uint32_t max_terms, uint32_t head_ranks,
                           char* buf, size_t cap, size_t* needed);

/* Device self-test: compares the kernels' inline correctly-rounded division with div.rn.f32 on n
 * pseudo-random operand pairs drawn from the validated range; *mismatches must come back 0. */
int ns_selftest_fastdiv(int device, uint64_t n, uint64_t seed, uint64_t* mismatches);

/* Debug library only (libnsb200_dbg.so, built with -DNSB_DEBUG_CHECKS): the score kernel checks every index it
 * derives from a posting, a tile table or a descriptor before using it — what the reference leaves unchecked
 * (seg.docs[docId], src/api_engine.cpp:477) — and counts failures per class; counts[i] receives class i
 * (0 item, 1 term descriptor, 2 tile window, 3 posting slice, 4 accumulator slot, 5 docId, 6 candidate, 7 result
 * list).  The product library returns NS_ERR_STATE. */
int ns_debug_violations(int device, uint64_t* counts, int n);
/* Debug library only: fails ONE check of class 7 on purpose (the counters are live). */
int ns_debug_selftest(int device);

#ifdef __cplusplus
}
#endif
#endif /* NEXTSEARCH_B200_H */
