"""The memcheck substitute (compute-sanitizer is closed on the development pool): run every kernel variant of the
path against the DEBUG library, whose score kernel checks each index it derives from a posting, a tile table or a
descriptor (NSB_DEBUG_CHECKS in csrc/bm25_kernels.cuh), compare every result with the CPU oracle, then read the
violation counters.  Prints one JSON line {"parity": [...], "violations": {...}, "debug_build": bool}.

    make -C nextsearch-api_b200/csrc debug
    NSB200_LIB=$PWD/nextsearch-api_b200/libnsb200_dbg.so python tools/debug_checks_probe.py

Workloads: 2 segments x 30k docs (15 tiles each, ragged last tile); 512 queries k = 10 and 100; forced item splits
(several warps merge into one query's list under the lock); a two-slot engine (publishing variant, wait, merge);
weighted and negative-weight term lists (generic arithmetic, dense selection); 100- and 250-term queries (NG = 4 / 8);
a foreign idf through the raw ABI (per-batch impact pre-pass); the unpacked posting format."""
import ctypes as C
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nsb200  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def same(res, oi, qs, k):
    _, s, g, d, nh, fo, hf = oi.search_many(qs, k, nthreads=4)
    ok = np.array_equal(res.nhits, nh) and np.array_equal(res.found, fo)
    for q in range(len(qs)):
        n = int(nh[q])
        ok = ok and np.array_equal(res.hits["score"][q, :n].view(np.uint32), s[q, :n].view(np.uint32))
        ok = ok and np.array_equal(res.hits["doc"][q, :n], d[q, :n]) and np.array_equal(res.hits["seg"][q, :n], g[q, :n])
    return bool(ok)


def same_weighted(res, oi, lists, k):
    ok = True
    for q, lst in enumerate(lists):
        want = oi.search_weighted(lst, k)
        n = len(want["results"])
        ok = ok and int(res.nhits[q]) == n
        ok = ok and res.hits["score"][q, :n].view(np.uint32).tolist() == [h["score_bits"] for h in want["results"]]
        ok = ok and res.hits["doc"][q, :n].tolist() == [h["docId"] for h in want["results"]]
    return bool(ok)


def main():
    lib = nsb200._lib.load()
    td = tempfile.mkdtemp(prefix="nsb200_dbg_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    spec = nsb200.CorpusSpec(vocab=20_000)
    nsb200.build_index(td, spec, 60_000, 2)
    oi = orc.OracleIndex(td)
    qs = nsb200.make_queries(spec, 500, 1, 5, seed=5) + ["t3 t3", "the of", "zzzz", "t1 t2 t3 t4 t5 t6 t7 t8"]
    qs += nsb200.make_queries(spec, 8, 1, 3, seed=6, head_ranks=50)
    parity = {}
    eng = nsb200.Engine(td, device=0)
    assert eng.reload(), eng.last_error
    parity["k10"] = same(eng.search_batch(qs, 10), oi, qs, 10)
    parity["k100"] = same(eng.search_batch(qs[:96], 100), oi, qs[:96], 100)
    parity["single_queries"] = all(same(eng.search_batch([q], 10), oi, [q], 10) for q in qs[:8])
    for splits in (2, 7, 64):                      # several items per query: lock + merge_back under contention
        q_off, terms, _ = eng.resolve_batch(qs[:128])
        b = eng.index.prepare(q_off, terms, 10)
        b.set_splits(splits)
        b.launch()
        h, n, f = b.fetch()
        _, s0, g0, d0, nh0, fo0, _ = oi.search_many(qs[:128], 10, nthreads=4)
        ok = np.array_equal(n, nh0) and np.array_equal(f, fo0)
        for q in range(128):
            m = int(nh0[q])
            ok = ok and np.array_equal(h["score"][q, :m].view(np.uint32), s0[q, :m].view(np.uint32)) and np.array_equal(h["doc"][q, :m], d0[q, :m])
        parity[f"splits_{splits}"] = bool(ok)
        b.close()
    rng = np.random.default_rng(3)
    long_qs = [" ".join(f"t{int(x)}" for x in rng.integers(1, 2000, n)) for n in (100, 250, 40, 65)] + qs[:12]
    parity["long_queries"] = same(eng.search_batch(long_qs, 10), oi, long_qs, 10)
    lists = [[(f"t{int(a)}", float(w)) for a, w in zip(rng.integers(1, 300, n), rng.uniform(0.1, 2.0, n))] for n in (1, 3, 7, 20, 38)]
    lists += [[("t1", 0.5), ("t2", -0.25), ("t3", 1.0)], [("t5", -1.0)], []]    # negative weights: dense selection
    for k in (10, 100):
        parity[f"weighted_k{k}"] = same_weighted(eng.search_terms_batch(lists, k), oi, lists, k)
    # a foreign idf through the raw ABI needs the raw postings next to the resident scores: a second, raw-keeping engine
    os.environ["NSB200_KEEP_RAW"] = "1"
    raw = nsb200.Engine(td, device=0)
    assert raw.reload(), raw.last_error
    del os.environ["NSB200_KEEP_RAW"]
    q_off, terms, _ = raw.resolve_batch(qs[:64])
    t2 = terms.copy()
    t2["idf"][::3] = t2["idf"][::3] * np.float32(1.5)      # every third term: idf != resident -> per-batch pre-pass
    h1, n1, f1 = raw.index.search_batch(q_off, terms, 10)
    h2, n2, f2 = raw.index.search_batch(q_off, t2, 10)
    parity["foreign_idf_runs"] = bool(np.array_equal(f1, f2) and np.array_equal(n1, n2))   # same docs touched, other scores
    _, s0, g0, d0, nh0, fo0, _ = oi.search_many(qs[:64], 10, nthreads=4)
    parity["raw_abi"] = bool(np.array_equal(n1, nh0) and np.array_equal(f1, fo0) and all(
        np.array_equal(h1["score"][q, :int(nh0[q])].view(np.uint32), s0[q, :int(nh0[q])].view(np.uint32)) for q in range(64)))
    raw.close()
    multi = nsb200.Engine(td, devices=[0, 0])
    assert multi.reload(), multi.last_error
    parity["two_slots_k10"] = same(multi.search_batch(qs, 10), oi, qs, 10)
    parity["two_slots_k100"] = same(multi.search_batch(qs[:64], 100), oi, qs[:64], 100)
    parity["two_slots_long"] = same(multi.search_batch(long_qs, 10), oi, long_qs, 10)
    multi.close()
    eng.close()
    os.environ["NSB200_NO_PACK"] = "1"              # read when the index handle is created
    unp = nsb200.Engine(td, device=0)
    assert unp.reload(), unp.last_error
    del os.environ["NSB200_NO_PACK"]
    parity["unpacked"] = same(unp.search_batch(qs[:128], 10), oi, qs[:128], 10)
    unp.close()
    counts = (C.c_uint64 * 8)()
    rc = lib.ns_debug_violations(0, counts, 8)
    names = ["item", "term", "tile", "slice", "acc", "doc", "cand", "list"]
    out = {"parity": parity, "parity_ok": all(parity.values()), "debug_build": rc == 0,
           "violations": {n: int(c) for n, c in zip(names, counts)} if rc == 0 else None,
           "lib": os.path.basename(nsb200._lib.LIB_PATH)}
    clean = rc != 0 or sum(counts) == 0
    if rc == 0:   # the counters are live: one deliberate failure shows up as exactly one "list" violation
        after = (C.c_uint64 * 8)()
        out["counters_live"] = bool(lib.ns_debug_selftest(0) == 0 and lib.ns_debug_violations(0, after, 8) == 0
                                    and int(after[7]) == int(counts[7]) + 1 and list(after)[:7] == list(counts)[:7])
    print(json.dumps(out))
    return 0 if out["parity_ok"] and clean and out.get("counters_live", True) else 1


if __name__ == "__main__":
    sys.exit(main())
