// Internal C++ seam between the host engine (host/engine.cpp) and the device layer (device_api.cu);
// not part of the C ABI.  The engine binds every call to ONE committed index generation: it snapshots
// the device state together with its host lexicons and prepares batches against exactly that state,
// so a concurrent ns_engine_reload can never pair old lexicon rows with a new device index.
#pragma once
#include <cstdint>
#include <memory>

#include "../../include/nextsearch_b200.h"

namespace nsb {

// The committed device index of `idx` right now (nullptr before the first commit).  Holding the pointer
// keeps that generation's device memory alive.
std::shared_ptr<const void> index_live_state(ns_index* idx);

// ns_batch_prepare against an explicit generation.
int batch_prepare_on(ns_index* idx, const std::shared_ptr<const void>& state, uint32_t Q, int k, const uint64_t* q_off,
                     const ns_qterm* terms, ns_batch** out);

// The engine's fast path.  Its front end already knows, from its own generation, everything ns_batch_prepare has
// to look up and check per term (which slot a segment occupies on the device, that the row exists, its posting
// count, that the idf is the one the resident scores were built with), so it hands over the terms in the
// kernel's own form together with the per-query posting totals.  Layout of PreparedTerm == DevTerm.
struct PreparedTerm {
    uint32_t slot;   // position of the term's segment among the device's segments (ascending global index)
    uint32_t row;
    float idf;
    float w;
    uint32_t delta, scratch;  // 0, 0: scores are resident
};
struct PreparedBatch {
    const uint32_t* qoff;        // [Q+1] into terms
    const PreparedTerm* terms;   // per query: by slot, then query-term order
    const uint64_t* weight;      // [Q] postings each query touches on this device
    uint32_t max_in_seg;         // most terms any (query, segment) has
    bool unit_weights;           // every w == 1.0f and every idf inside the fast-division range
    bool scan_always;            // some w or idf is negative / NaN
};
// NS_ERR_STATE when a segment of `state` has no resident scores (the caller then uses batch_prepare_on).
int batch_prepare_trusted(ns_index* idx, const std::shared_ptr<const void>& state, uint32_t Q, int k, const PreparedBatch& pb,
                          ns_batch** out);

// An exchange that only publishes (the non-root device slots of a multi-device engine): no gather regions, no
// merged blob, no pinned result buffer — creating it costs one small device allocation.
int exchange_create_publisher(int device, uint32_t world, uint32_t rank, uint32_t max_queries, uint32_t slots,
                              ns_exchange** out);

// Multi-device engine, root side of one step: order the root batch's stream after the score kernels of every
// other device's batch (events — same process, so nothing has to poll) and enqueue the merge of the `ndev`
// blobs the score kernels stored into the root's gather buffer.  batches[0] is the root's.
int exchange_root_merge(ns_exchange* root, ns_batch* const* batches, int ndev, uint64_t step, uint32_t Q, int k);

}  // namespace nsb
