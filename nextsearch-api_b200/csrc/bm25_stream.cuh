// Streamed variant of the BM25 score + top-k kernel (reference loop: src/api_engine.cpp:441-505).
//
// Same data layout, work items, accumulator tiles, candidate scheme, shared result lists and float
// operation order as bm25_score_topk_kernel (bm25_kernels.cuh) — it reuses that file's helpers — but
// the postings of an item are consumed as ONE STREAM OF CHUNKS (<= 128 postings of one (term, tile)
// slice) that runs across term and tile boundaries, and the loads of chunk i+1 are issued before
// chunk i is accumulated.  The nest kernel exposed one L2/HBM round trip per slice (plus one per
// tail): ~3.5 per (query, tile), 30 % of all stall samples on the first use of a loaded posting
// (profiles/r1_v7_final_ncu.txt).  Here only the first chunk of an (item, segment) is exposed.
//
// Loader side:  (b0, b1) per lane = bounds of term `lane` in the loader's tile jL; nx = tile-table entry
//               of the tile after, in flight.  next_chunk() pops terms in query order, advances over
//               empty tiles, and tags each chunk with its tile, its length and "first term of the tile".
// Consumer side: when the tile tag changes, the finished tile is folded into the top-k list and reset.
#pragma once
#include "bm25_kernels.cuh"

namespace nsb {

// Fold the finished tile's candidates into the warp's sorted list, then reset the tile.
// scan_mode: the candidate buffer was not maintained for this tile (list not full / negative weights).
template <int TDW, int KCAP>
__device__ __forceinline__ void fold_and_reset(WarpSmem<TDW, KCAP>& ws, uint32_t k, bool scan_mode, float thr_c,
                                               uint32_t base, uint32_t gseg, uint32_t& ntop, float& thr, uint32_t lane) {
    float* acc = ws.acc;
    float4* acc4 = reinterpret_cast<float4*>(ws.acc);
    const float4 sent4 = make_float4(__uint_as_float(kSentinel), __uint_as_float(kSentinel), __uint_as_float(kSentinel),
                                     __uint_as_float(kSentinel));
    uint32_t cnt = ws.cnt;
    if (scan_mode || cnt != 0u) {
        bool slow = false;
        if (scan_mode || cnt > kCandCap) {
            slow = true;
            if (k <= 32u) {
                // Dense tile, short list: threshold T0 = k-th largest of the per-lane maxima (>= k
                // accumulators are >= T0, so nothing below T0 can be in the tile's top k); collect
                // everything >= T0 and above thr as candidates.
                float lm = -INFINITY;
                for (uint32_t i = lane; i < TDW / 4; i += 32) {
                    const float4 v = acc4[i];
                    lm = fmaxf(lm, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));  // fmaxf drops the NaN sentinel
                }
                if (!(lm > thr_c)) lm = -INFINITY;
                const float t0 = warp_kth_largest(lm, k, lane);
                if (lane == 0) ws.cnt = 0;
                __syncwarp();
                for (uint32_t i = lane; i < TDW / 4; i += 32) {
                    const float4 v = acc4[i];
                    const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
                    if (mx >= t0 && mx > thr_c) {
                        const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            if (x[c] >= t0 && x[c] > thr_c) {
                                const uint32_t at = atomicAdd(&ws.cnt, 1u);
                                if (at < kCandCap) ws.cand[at] = base + 4u * i + (uint32_t)c;
                            }
                        }
                    }
                }
                __syncwarp();
                cnt = ws.cnt;
                slow = cnt > kCandCap;
            }
        }
        if (slow) {
            // general path: extract the tile's hits above thr in order until one fails to enter
            float prev_s = INFINITY;
            uint32_t prev_d = 0;
            for (;;) {
                float bs = -INFINITY;
                uint32_t bd = kNone;
                for (uint32_t i = lane; i < TDW / 4; i += 32) {
                    const float4 v = acc4[i];
                    const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const uint32_t d = base + 4u * i + (uint32_t)c;
                        const bool after_prev = (x[c] < prev_s) || (x[c] == prev_s && d > prev_d);
                        if (x[c] > thr_c && after_prev) {  // false for the NaN sentinel
                            if (bd == kNone || x[c] > bs || (x[c] == bs && d < bd)) {
                                bs = x[c];
                                bd = d;
                            }
                        }
                    }
                }
                warp_best(bs, bd);
                if (bd == kNone) break;
                if (!list_insert<KCAP>(ws.top_s, ws.top_d, ws.top_g, ntop, thr, k, bs, gseg, bd, lane)) break;
                prev_s = bs;
                prev_d = bd;
            }
        } else if (cnt > 0) {
            float cs = -INFINITY;
            uint32_t cd = kNone;
            if (lane < cnt) {
                cd = ws.cand[lane];
                cs = acc[cd - base];  // final value: all terms of this tile are done
            }
            for (uint32_t r = 0; r < cnt; r++) {
                float bs = cs;
                uint32_t bd = cd;
                warp_best(bs, bd);
                if (bd == kNone) break;
                if (!list_insert<KCAP>(ws.top_s, ws.top_d, ws.top_g, ntop, thr, k, bs, gseg, bd, lane)) break;
                if (cd == bd) cd = kNone;  // a doc may have been recorded more than once
            }
        }
        if (lane == 0) ws.cnt = 0;
    }
#pragma unroll
    for (uint32_t i = 0; i < TDW / 128; i++) acc4[32u * i + lane] = sent4;
    __syncwarp();
}

// One chunk in registers: <= 128 consecutive postings of a tile's FLAT posting sequence — the tile's
// slices concatenated in query-term order.  A chunk that lies inside one slice is "single" (docIds
// unique: the four steps are batched); one that spans a slice boundary is "mixed" (two lanes of a
// step may name the same doc through different terms: steps are sequential and ordered by lane).
// tag = rem (bits 0..7, 1..128) | kFirst | kMixed | tile index within the segment << 10
struct Chunk {
    uint2 e[4];
    uint32_t tag;
    uint32_t tls;  // term (lane index) of this lane's posting in step u, 8 bits per step; only read by
                   // the variants that evaluate idf/weight in the kernel (dead in the FAST impact variant)
};
constexpr uint32_t kFirst = 0x100u;  // single chunk of the tile's first non-empty term: accumulators still unset
constexpr uint32_t kMixed = 0x200u;

// e = *p when ok; otherwise e keeps its (stale) value — no branch, no clamp.
__device__ __forceinline__ void ld_stream_u2_if(uint2& e, const uint2* p, bool ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t"
        "@p ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];\n\t}"
        : "+r"(e.x), "+r"(e.y)
        : "l"(p), "r"((uint32_t)ok));
}

// Per-lane BM25 contribution of step u (src/api_engine.cpp:477-480).  FAST impact: the stored score.
template <bool FAST, int PAY, bool PER_LANE_TERM>
__device__ __forceinline__ float step_contrib(const PassCtx& c, uint2 e, bool valid, uint32_t tl, float t_idf, float t_w) {
    float idf = 0.f, w = 0.f;
    if (PER_LANE_TERM) {
        idf = __shfl_sync(0xffffffffu, t_idf, tl);
        w = __shfl_sync(0xffffffffu, t_w, tl);
    }
    if (PAY == kPayImpact) {
        const float sc = __uint_as_float(e.y);
        return FAST ? sc : __fmul_rn(w, sc);
    }
    uint32_t gi = PAY == kPayPacked ? (e.y >> 16) : e.x;
    if (!valid) gi = 0u;  // stale registers: stay inside the table
    const float nr = ld_norm(c.norm + gi);
    return bm25_contrib<FAST>(PAY == kPayPacked ? (e.y & 0xFFFFu) : e.y, nr, idf, w, c.k1p1);
}

// rare: after this chunk some accumulator is above the threshold — remember the docs
template <int NS>
__device__ __forceinline__ void chunk_crossers(const PassCtx& c, const uint2 (&e)[4], uint32_t rem) {
#pragma unroll
    for (int u = 0; u < NS; u++) {
        if (32u * u + c.lane < rem) {
            const float v = lds_f32(c.sacc + 4u * e[u].x);
            if (v > c.thr_eff) {
                const uint32_t at = atomicAdd(c.cnt, 1u);
                if (at < kCandCap) asm volatile("st.shared.u32 [%0], %1;" ::"r"(c.scand + 4u * at), "r"(e[u].x) : "memory");
            }
        }
    }
}

// Single chunk, NS steps (1: rem <= 32; 4: 32 < rem <= 128).  The registers of lanes beyond `rem` are
// STALE: every use is predicated.  All reads, then the arithmetic, then all writes (docIds are
// unique inside a slice).
template <int NS, bool FIRST, bool FAST, int PAY>
__device__ __forceinline__ void single_steps(const PassCtx& c, const Chunk& s, uint32_t rem, float t_idf, float t_w,
                                             uint32_t& my_found) {
    constexpr bool PLT = !(FAST && PAY == kPayImpact);
    const uint32_t lane = c.lane;
    bool valid[4];
#pragma unroll
    for (int u = 0; u < NS; u++) valid[u] = (NS > 1 && u == 0) ? true : (32u * u + lane < rem);
    float x[4];
#pragma unroll
    for (int u = 0; u < NS; u++) x[u] = step_contrib<FAST, PAY, PLT>(c, s.e[u], valid[u], (s.tls >> (8 * u)) & 0xFFu, t_idf, t_w);
    float old[4];
    if (!FIRST) {
#pragma unroll
        for (int u = 0; u < NS; u++) {
            old[u] = 0.0f;
            if (valid[u]) old[u] = lds_f32(c.sacc + 4u * s.e[u].x);
        }
    }
    bool cross = false;
#pragma unroll
    for (int u = 0; u < NS; u++) {
        const uint32_t addr = c.sacc + 4u * s.e[u].x;
        float nv;
        if (FIRST) {
            nv = FAST ? x[u] : __fadd_rn(0.0f, x[u]);
        } else {
            const bool fresh = __float_as_uint(old[u]) == kSentinel;
            nv = __fadd_rn(fresh ? 0.0f : old[u], x[u]);
            my_found += (fresh && valid[u]) ? 1u : 0u;
        }
        if (valid[u]) sts_f32(addr, nv);
        cross |= valid[u] && (nv > c.thr_eff);
    }
    if (__any_sync(0xffffffffu, cross)) chunk_crossers<NS>(c, s.e, rem);
}

// Mixed chunk: step by step; inside a step, lanes that name the same doc (through different terms —
// lower lane = earlier term) take turns in lane order, so every doc still sees its additions in
// query-term order.
template <int TDW, bool FAST, int PAY>
__device__ __forceinline__ void mixed_steps(const PassCtx& c, const Chunk& s, uint32_t rem, float t_idf, float t_w,
                                            uint32_t& my_found) {
    constexpr bool PLT = !(FAST && PAY == kPayImpact);
    const uint32_t lane = c.lane;
    bool cross = false;
#pragma unroll
    for (int u = 0; u < 4; u++) {
        if (u > 0 && rem <= 32u * u) break;  // warp-uniform
        const bool valid = 32u * u + lane < rem;
        const float x = step_contrib<FAST, PAY, PLT>(c, s.e[u], valid, (s.tls >> (8 * u)) & 0xFFu, t_idf, t_w);
        const uint32_t addr = c.sacc + 4u * s.e[u].x;
        // key: the accumulator address for valid lanes (inside the tile), a unique value outside it otherwise
        const uint32_t key = valid ? addr : (0xFFFFFF00u + lane);
        const uint32_t grp = __match_any_sync(0xffffffffu, key);
        const uint32_t rank = __popc(grp & ((1u << lane) - 1u));
        uint32_t r = 0;
        for (;;) {
            if (valid && rank == r) {
                const float old = lds_f32(addr);
                const bool fresh = __float_as_uint(old) == kSentinel;
                const float nv = __fadd_rn(fresh ? 0.0f : old, x);
                my_found += fresh ? 1u : 0u;
                sts_f32(addr, nv);
                cross |= nv > c.thr_eff;
            }
            __syncwarp();
            r++;
            if (!__any_sync(0xffffffffu, valid && rank >= r)) break;
        }
    }
    if (__any_sync(0xffffffffu, cross)) chunk_crossers<4>(c, s.e, rem);
}

// All tiles [j0, j1) of one segment for one item (at most 32 terms per (query, segment): lane t
// holds term t).  PAY is uniform per segment.
template <int TDW, int KCAP, bool FAST, bool IMPACT, int PAY>
__device__ __forceinline__ void stream_slot(const ScoreArgs& a, WarpSmem<TDW, KCAP>& ws, PassCtx& ctx, const DevSeg& seg,
                                            const uint32_t* to, uint32_t t_delta, uint32_t t_scr, float t_idf, float t_w,
                                            uint32_t nt, uint32_t j0, uint32_t j1, uint32_t k, uint32_t& ntop, float& thr,
                                            uint32_t& my_found, uint32_t acc_saddr) {
    constexpr bool PLT = !(FAST && PAY == kPayImpact);
    const uint32_t lane = ctx.lane;
    const uint2* seg_post = IMPACT ? seg.imp : seg.post;
    const bool scratch = IMPACT && a.any_scratch != 0u;

    // ---- loader state ----
    uint32_t jL = j0;
    uint32_t b0 = 0u, b1 = 0u, nx = 0u;  // per lane: slice of term `lane` in tile jL; tile-table entry of the tile after
    if (lane < nt) {
        b0 = __ldg(to + j0) + t_delta;
        b1 = __ldg(to + j0 + 1) + t_delta;
        if (j0 + 1 < j1) nx = __ldg(to + j0 + 2);
    }
    uint32_t sincl = 0u, dlt = 0u;  // per lane: inclusive prefix of slice lengths; posting index minus flat index
    uint32_t f = 0u, M = 0u;        // flat cursor / flat length of tile jL (warp-uniform from here on)
    uint32_t mm = 0u;               // non-empty terms after tcur
    uint32_t tcur = 0u, tfirst = 0u, tE = 0u, dcur = 0u;
    const uint2* pcur = seg_post;

    auto next_term = [&]() {
        tcur = (uint32_t)__ffs((int)mm) - 1u;
        mm &= mm - 1u;
        tE = __shfl_sync(0xffffffffu, sincl, tcur);
        dcur = __shfl_sync(0xffffffffu, dlt, tcur);
        if (scratch) pcur = __shfl_sync(0xffffffffu, t_scr, tcur) != 0u ? a.impacts : seg_post;
    };
    auto setup_tile = [&](uint32_t nonempty) {
        const uint32_t len = b1 - b0;  // 0 on lanes without a term
        uint32_t sc = len;
        for (uint32_t d = 1; d < nt; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, sc, d);
            if (lane >= d) sc += o;
        }
        sincl = sc;
        dlt = b0 - (sc - len);
        M = __shfl_sync(0xffffffffu, sc, nt - 1u);
        f = 0u;
        mm = nonempty;
        next_term();
        tfirst = tcur;
    };
    // move the loader to the next tile that has postings; false when the segment window is exhausted
    auto advance_tile = [&]() -> bool {
        for (;;) {
            if (++jL >= j1) return false;
            b0 = b1;
            b1 = nx + t_delta;  // nx was requested one tile ago; 0 + 0 on lanes without a term
            if (lane < nt && jL + 1 < j1) nx = __ldg(to + jL + 2);
            const uint32_t ne = __ballot_sync(0xffffffffu, b1 != b0);
            if (ne != 0u) {
                setup_tile(ne);
                return true;
            }
        }
    };
    {
        const uint32_t ne = __ballot_sync(0xffffffffu, b1 != b0);
        if (ne != 0u) setup_tile(ne);
    }

    auto load_chunk = [&](Chunk& s) -> bool {
        if (f == M) {
            if (!advance_tile()) return false;
        }
        const uint32_t rem = min(M - f, 128u);
        uint32_t tag = rem | (jL << 10);
        if (f + rem <= tE) {
            // inside the slice of term tcur
            if (tcur == tfirst) tag |= kFirst;
            const uint2* p = pcur + (f + dcur + lane);
            ld_stream_u2_if(s.e[0], p, lane < rem);
            if (rem > 32u) {
                ld_stream_u2_if(s.e[1], p + 32, 32u + lane < rem);
                ld_stream_u2_if(s.e[2], p + 64, 64u + lane < rem);
                ld_stream_u2_if(s.e[3], p + 96, 96u + lane < rem);
            }
            if (PLT) s.tls = tcur * 0x01010101u;
            f += rem;
            if (f == tE && f < M) next_term();
        } else {
            tag |= kMixed;
            uint32_t tl = tcur, el = tE, tls = 0u;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (u > 0 && rem <= 32u * u) break;  // warp-uniform
                const uint32_t fl = f + 32u * u + lane;
                const bool valid = 32u * u + lane < rem;
                for (;;) {  // advance this lane's term until its slice contains fl
                    const bool adv = valid && fl >= el;
                    if (!__any_sync(0xffffffffu, adv)) break;
                    tl += adv ? 1u : 0u;
                    el = __shfl_sync(0xffffffffu, sincl, tl);
                }
                const uint32_t dl = __shfl_sync(0xffffffffu, dlt, tl);
                const uint2* pb = seg_post;
                if (scratch) pb = __shfl_sync(0xffffffffu, t_scr, tl) != 0u ? a.impacts : seg_post;
                ld_stream_u2_if(s.e[u], pb + (fl + dl), valid);
                if (PLT) tls |= tl << (8 * u);
            }
            if (PLT) s.tls = tls;
            f += rem;
            if (f < M) {
                while (f >= tE) next_term();
            }
        }
        s.tag = tag;
        return true;
    };

    // ---- consumer state ----
    uint32_t jA = kNone;  // tile being accumulated
    bool scan_mode = false;
    float thr_c = thr;

    Chunk cur, nxt;
    cur.e[0] = cur.e[1] = cur.e[2] = cur.e[3] = make_uint2(0u, 0u);
    nxt.e[0] = nxt.e[1] = nxt.e[2] = nxt.e[3] = make_uint2(0u, 0u);
    cur.tag = nxt.tag = 0u;
    cur.tls = nxt.tls = 0u;
    bool have = load_chunk(cur);
#pragma unroll 1
    for (;;) {
        bool have_n = false;
        if (have) have_n = load_chunk(nxt);  // in flight while `cur` is accumulated
        const uint32_t tile = have ? (cur.tag >> 10) : kNone;
        if (tile != jA) {
            if (jA != kNone) fold_and_reset<TDW, KCAP>(ws, k, scan_mode, thr_c, jA * (uint32_t)TDW, seg.gseg, ntop, thr, lane);
            jA = tile;
            const uint32_t base = tile * (uint32_t)TDW;
            // dense selection is only needed while the list is not full
            scan_mode = (ntop < k) || (a.scan_always != 0u);
            // What a doc must exceed to be a candidate: the k-th score — or its predecessor when a
            // doc of this tile could still win a tie against the k-th entry on (segment, docId).
            thr_c = thr;
            if (ntop == k) {
                const uint32_t kg = ws.top_g[k - 1] & ~kForeign, kd = ws.top_d[k - 1];
                if (seg.gseg < kg || (seg.gseg == kg && base < kd)) thr_c = float_pred(thr);
            }
            ctx.thr_eff = scan_mode ? INFINITY : thr_c;
            ctx.sacc = acc_saddr - 4u * base;
        }
        if (!have) break;  // (also when the window held no posting at all: jA == tile == kNone)
        const uint32_t rem = cur.tag & 0xFFu;
        if (cur.tag & kMixed) {
            mixed_steps<TDW, FAST, PAY>(ctx, cur, rem, t_idf, t_w, my_found);
        } else if (cur.tag & kFirst) {
            my_found += (rem > lane) ? ((rem - lane + 31u) >> 5) : 0u;  // each posting is a new doc
            if (rem <= 32u) single_steps<1, true, FAST, PAY>(ctx, cur, rem, t_idf, t_w, my_found);
            else single_steps<4, true, FAST, PAY>(ctx, cur, rem, t_idf, t_w, my_found);
        } else {
            if (rem <= 32u) single_steps<1, false, FAST, PAY>(ctx, cur, rem, t_idf, t_w, my_found);
            else single_steps<4, false, FAST, PAY>(ctx, cur, rem, t_idf, t_w, my_found);
        }
        __syncwarp();
        cur = nxt;
        have = have_n;
    }
}

// At most 32 terms per (query, segment); wider batches run bm25_score_topk_kernel<..., NG = 2>.
template <int TDW, int KCAP, bool FAST, bool IMPACT>
__global__ void __launch_bounds__(kThreads, 3) bm25_stream_kernel(const ScoreArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using WS = WarpSmem<TDW, KCAP>;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    WS& ws = reinterpret_cast<WS*>(smem_raw)[warp];
    float4* acc4 = reinterpret_cast<float4*>(ws.acc);
    const float4 sent4 = make_float4(__uint_as_float(kSentinel), __uint_as_float(kSentinel),
                                     __uint_as_float(kSentinel), __uint_as_float(kSentinel));
    const uint32_t k = a.k;
    const uint32_t acc_saddr = (uint32_t)__cvta_generic_to_shared(ws.acc);

    PassCtx ctx;
    ctx.k1p1 = a.k1p1;
    ctx.lane = lane;
    ctx.zero = a.zero;
    ctx.scand = (uint32_t)__cvta_generic_to_shared(ws.cand);
    ctx.cnt = &ws.cnt;

    for (uint32_t i = lane; i < TDW / 4; i += 32) acc4[i] = sent4;
    if (lane == 0) ws.cnt = 0;
    __syncwarp();

    for (;;) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(a.counter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= a.nitems) break;
        const DevItem it = a.items[item];
        const uint32_t q = it.q, split = it.split_ns >> 16, nsplit = it.split_ns & 0xFFFFu;
        const uint32_t e0 = a.qoff[q], e1 = a.qoff[q + 1];
        const uint32_t g0 = (uint32_t)(((uint64_t)a.total_tiles * split) / nsplit);
        const uint32_t g1 = (uint32_t)(((uint64_t)a.total_tiles * (split + 1)) / nsplit);

        uint32_t ntop = 0;
        float thr = -INFINITY;   // k-th best of THIS item's list (-inf until it is full)
        uint32_t my_found = 0;
        uint32_t ecur = e0;      // entries are sorted by slot: a cursor suffices
        // Seed the list with the query's shared list (see bm25_score_topk_kernel).
        if (__ldcg(a.nhits + q) != 0u) {
            uint32_t* lk = a.qlock + q;
            qlock_acquire(lk, lane);
            const uint32_t n_g = __ldcg(a.nhits + q);
            const uint32_t* gh = reinterpret_cast<const uint32_t*>(a.hits + (size_t)q * k);
            for (uint32_t e = lane; e < n_g; e += 32) {
                ws.top_s[e] = __uint_as_float(__ldcg(gh + 3u * e));
                ws.top_g[e] = __ldcg(gh + 3u * e + 1u) | kForeign;
                ws.top_d[e] = __ldcg(gh + 3u * e + 2u);
            }
            qlock_release(lk, lane);
            ntop = n_g;
            __syncwarp();
            if (ntop == k) thr = ws.top_s[k - 1];
        }

        for (uint32_t slot = 0; slot < a.nseg && ecur < e1; slot++) {
            const uint32_t tb0 = a.tile_base[slot], tb1 = a.tile_base[slot + 1];
            if (tb1 <= g0) continue;
            if (tb0 >= g1) break;
            const uint32_t j0 = (g0 > tb0 ? g0 : tb0) - tb0;
            const uint32_t j1 = (g1 < tb1 ? g1 : tb1) - tb0;

            // advance the cursor to the first entry with slot >= this one
            for (;;) {
                const uint32_t e = ecur + lane;
                const bool lt = (e < e1) && (a.terms[e].slot < slot);
                const uint32_t mk = __ballot_sync(0xffffffffu, lt);
                ecur += __popc(mk);
                if (mk != 0xffffffffu) break;
            }
            // this segment's terms: lane t holds term t
            bool mine = false;
            DevTerm t = {0u, 0u, 0.f, 0.f, 0u, 0u};
            if (ecur + lane < e1) {
                t = a.terms[ecur + lane];
                mine = (t.slot == slot);
            }
            const uint32_t nt = __popc(__ballot_sync(0xffffffffu, mine));  // a prefix of the lanes
            if (nt == 0) continue;
            const uint32_t t_delta = (IMPACT && mine) ? t.delta : 0u;
            const uint32_t t_scr = IMPACT ? t.scratch : 0u;

            const DevSeg seg = a.segs[slot];
            const bool packed = seg.packed != 0u;
            ctx.norm = packed ? seg.lut : seg.norm;
            const uint32_t* to = seg.tileoff + (size_t)t.row * (seg.ntiles + 1);

            if (IMPACT)
                stream_slot<TDW, KCAP, FAST, IMPACT, kPayImpact>(a, ws, ctx, seg, to, t_delta, t_scr, t.idf, t.w, nt, j0, j1, k, ntop,
                                                                 thr, my_found, acc_saddr);
            else if (packed)
                stream_slot<TDW, KCAP, FAST, IMPACT, kPayPacked>(a, ws, ctx, seg, to, t_delta, t_scr, t.idf, t.w, nt, j0, j1, k, ntop,
                                                                 thr, my_found, acc_saddr);
            else
                stream_slot<TDW, KCAP, FAST, IMPACT, kPayRaw>(a, ws, ctx, seg, to, t_delta, t_scr, t.idf, t.w, nt, j0, j1, k, ntop,
                                                              thr, my_found, acc_saddr);
        }

        // ---- merge this item's own hits into the query's shared list ----
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) my_found += __shfl_xor_sync(0xffffffffu, my_found, off);
        if (lane == 0 && my_found != 0u) atomicAdd(a.found + q, (unsigned long long)my_found);
        merge_back<TDW, KCAP>(ws, a.hits + (size_t)q * k, a.nhits + q, a.qlock + q, k, ntop, lane, a.scan_always == 0u);
        __syncwarp();
    }
}

}  // namespace nsb
