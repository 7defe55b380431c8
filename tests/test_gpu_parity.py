"""CUDA path vs oracle through the C ABI: bit-exact docIds, order, found and f32 score bits.

Tolerance: none — scores are compared as u32 bit patterns (north_star allows 1e-5 relative; the
kernels reproduce the reference's operation tree exactly, so the stricter bar holds).
"""
import os

import numpy as np
import pytest

import nsb200
from conftest import EDGE_QUERIES, assert_same_as_oracle, make_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small_engine(small_case):
    e = nsb200.Engine(small_case.path, device=0)
    assert e.reload(), e.last_error
    yield e
    e.close()


@pytest.fixture(scope="module")
def config1_engine(config1_case):
    e = nsb200.Engine(config1_case.path, device=0)
    assert e.reload(), e.last_error
    yield e
    e.close()


@pytest.mark.parametrize("k", [1, 3, 10, 100])
def test_small_two_segments(small_case, small_engine, k):
    qs = nsb200.make_queries(small_case.spec, 300, 1, 5) + EDGE_QUERIES
    assert_same_as_oracle(small_engine.search_batch(qs, k), small_case.oracle, qs, k)


def test_k_is_clamped_like_the_reference(small_case, small_engine):
    qs = nsb200.make_queries(small_case.spec, 20, 1, 3)
    for k in (0, -5, 1000):
        res = small_engine.search_batch(qs, k)
        assert res.k == max(1, min(k, 100))
        assert_same_as_oracle(res, small_case.oracle, qs, k)


def test_config1_10k_docs(config1_case, config1_engine):
    """BASELINE configs[0]: 1000 queries of 1-3 terms, query seed 7, k=10."""
    qs = nsb200.make_queries(config1_case.spec, 1000, 1, 3, seed=nsb200.QUERY_SEED)
    assert_same_as_oracle(config1_engine.search_batch(qs, 10), config1_case.oracle, qs, 10)


def test_single_query_json_fields(small_case, small_engine):
    for q in ["t1 t2", "T7, the t9!", "zzzz", "the of"]:
        got = small_engine.search(q, 10)
        want = small_case.oracle.search(q, 10)
        assert got["k"] == want["k"] and got["segments"] == want["segments"] and got["query"] == q
        assert got.get("found") == want["found"]
        assert ("found" in got) == (want["found"] is not None)
        assert [(r["segment"], r["docId"], r["cord_uid"]) for r in got["results"]] == \
               [(r["segment"], r["docId"], r["cord_uid"]) for r in want["results"]]
        # JSON carries the f32 widened to double: must round-trip to the same f32
        assert [np.float32(r["score"]).view(np.uint32) for r in got["results"]] == \
               [r["score_bits"] for r in want["results"]]


@pytest.mark.parametrize("splits", [1, 2, 3, 7, 64])
def test_splits_do_not_change_results(small_case, small_engine, splits):
    """(query, split) decomposition + merge kernel == one CTA per query."""
    qs = nsb200.make_queries(small_case.spec, 64, 1, 5) + EDGE_QUERIES
    q_off, terms, has = small_engine.resolve_batch(qs)
    idx = small_engine.index
    for k in (10, 100):
        b = idx.prepare(q_off, terms, k)
        b.set_splits(splits)
        b.launch()
        hits, nhits, found = b.fetch()
        b.close()
        res = nsb200.BatchResult(hits, nhits, found, has, nsb200.clamp_k(k))
        assert_same_as_oracle(res, small_case.oracle, qs, k)


def test_single_query_batches(small_case, small_engine):
    """Q=1 goes through the split path (many CTAs per query)."""
    for q in nsb200.make_queries(small_case.spec, 12, 1, 5, seed=99):
        assert_same_as_oracle(small_engine.search_batch([q], 10), small_case.oracle, [q], 10)


def test_weighted_terms_via_abi(small_case, small_engine):
    """Weights flow through ns_qterm.weight (semantic expansion, src/api_engine.cpp:410-421):
    weight 2.0 on a single-term query must double every score exactly (power of two)."""
    qs = ["t2", "t9", "t30"]
    q_off, terms, _ = small_engine.resolve_batch(qs)
    idx = small_engine.index
    h1, n1, f1 = idx.search_batch(q_off, terms, 10)
    t2 = terms.copy()
    t2["weight"] = 2.0
    h2, n2, f2 = idx.search_batch(q_off, t2, 10)
    assert np.array_equal(n1, n2) and np.array_equal(f1, f2)
    for q in range(len(qs)):
        n = int(n1[q])
        assert np.array_equal(h2["score"][q, :n], h1["score"][q, :n] * np.float32(2.0))
        assert np.array_equal(h2["doc"][q, :n], h1["doc"][q, :n])


def test_more_than_32_terms_per_query(small_case, small_engine):
    """The reference's semantic expansion yields up to 40 terms per query (src/api_engine.cpp:417); a batch
    with more than 32 terms for one (query, segment) runs the two-register-group kernel variant (NG = 2).
    Wide and narrow queries mixed, duplicates included, k = 10 and 100."""
    import random

    rng = random.Random(9)
    wide = [" ".join(f"t{rng.randint(1, 400)}" for _ in range(n)) for n in (33, 40, 47, 64, 36, 50)]
    wide.append(" ".join(f"t{i}" for i in range(1, 41)))                 # 40 distinct head terms
    wide.append(" ".join(["t3", "t7"] * 20))                             # 40 terms, two distinct
    qs = wide + nsb200.make_queries(small_case.spec, 40, 1, 5, seed=13) + EDGE_QUERIES
    for k in (10, 100):
        assert_same_as_oracle(small_engine.search_batch(qs, k), small_case.oracle, qs, k)


def test_long_queries_up_to_256_terms(small_case, small_engine):
    """The reference puts no cap on the number of query terms (a pasted paragraph is a legal query); up to
    NS_MAX_TERMS = 256 terms per (query, segment) run the NG = 4 / NG = 8 kernel variants, mixed here with short
    queries in one batch and alone, k = 10 and 100, also through a two-slot engine (the publishing variants)."""
    import random

    rng = random.Random(21)
    long_qs = [" ".join(f"t{rng.randint(1, 600)}" for _ in range(n)) for n in (65, 100, 128, 129, 200, 256)]
    long_qs.append(" ".join(f"t{i}" for i in range(1, 257)))             # 256 distinct head terms
    long_qs.append(" ".join(["t3", "t7", "t11"] * 70))                   # 210 terms, three distinct
    mixed = long_qs + nsb200.make_queries(small_case.spec, 24, 1, 5, seed=17) + EDGE_QUERIES
    for k in (10, 100):
        assert_same_as_oracle(small_engine.search_batch(mixed, k), small_case.oracle, mixed, k)
    assert_same_as_oracle(small_engine.search_batch(long_qs[:2], 10), small_case.oracle, long_qs[:2], 10)   # NG = 4 alone
    multi = nsb200.Engine(small_case.path, devices=[0, 0])
    assert multi.reload(), multi.last_error
    assert_same_as_oracle(multi.search_batch(mixed, 10), small_case.oracle, mixed, 10)
    # a FEW long queries in a large batch are scored as a batch of their own (the short ones keep the NG = 1 kernel)
    # and the answers come back in the caller's order: long queries first, last and in the middle
    short = nsb200.make_queries(small_case.spec, 400, 1, 5, seed=19)
    big = [long_qs[6]] + short[:150] + [long_qs[1], long_qs[3]] + short[150:] + EDGE_QUERIES + [long_qs[7]]
    for k in (10, 100):
        assert_same_as_oracle(small_engine.search_batch(big, k), small_case.oracle, big, k)
    assert_same_as_oracle(multi.search_batch(big, 10), small_case.oracle, big, 10)
    multi.close()
    # one more than the ABI limit is refused, not truncated
    with pytest.raises(nsb200._lib.NsError):
        small_engine.search_batch([" ".join(f"t{i}" for i in range(1, 258))], 10)


def test_empty_and_termless_batches(small_case, small_engine):
    """Q = 0, and batches in which no query has a usable term (no item, no kernel work): like the reference,
    results are empty, `found` is 0 and absent for the all-stopword / empty queries (src/api_engine.cpp:407)."""
    res = small_engine.search_batch([], 10)
    assert len(res.nhits) == 0 and len(res.found) == 0
    qs = ["the of and", "", "a b c", "zzzz", "qqqq wwww"]
    res = small_engine.search_batch(qs, 10)
    assert_same_as_oracle(res, small_case.oracle, qs, 10)
    assert list(res.nhits) == [0] * 5 and list(res.found) == [0] * 5
    assert list(res.has_found) == [False, False, False, True, True]


def test_rejects_unsorted_postings():
    idx = nsb200.DeviceIndex(0)
    doc_len = np.full(16, 10, np.uint32)
    post = np.array([[3, 1], [2, 1], [5, 1]], np.uint32)  # 3,2 out of order
    with pytest.raises(nsb200._lib.NsError) as ei:
        idx.add_segment(0, 10.0, doc_len, np.array([0], np.uint64), np.array([3], np.uint32), post)
    assert ei.value.status == 4  # NS_ERR_FORMAT
    post = np.array([[3, 1], [99, 1]], np.uint32)  # docId >= N
    with pytest.raises(nsb200._lib.NsError):
        idx.add_segment(0, 10.0, doc_len, np.array([0], np.uint64), np.array([2], np.uint32), post)
    idx.close()


def test_search_before_commit_fails_loudly():
    idx = nsb200.DeviceIndex(0)
    with pytest.raises(nsb200._lib.NsError) as ei:
        idx.search_batch(np.zeros(2, np.uint64), np.zeros(0, nsb200.QTERM_DTYPE), 10)
    assert ei.value.status == 6  # NS_ERR_STATE
    idx.close()


def test_reload_swaps_index_and_keeps_old_on_failure(workdir, small_case):
    """Engine::reload semantics (src/api_engine.cpp:82-90): failure keeps the previous segments."""
    import shutil
    path = os.path.join(workdir, "reload_case")
    shutil.copytree(small_case.path, path)
    e = nsb200.Engine(path, device=0)
    assert e.reload()
    qs = nsb200.make_queries(small_case.spec, 32, 1, 3)
    assert_same_as_oracle(e.search_batch(qs, 10), small_case.oracle, qs, 10)
    os.remove(os.path.join(path, "segments", "seg_000002", "lexicon_b017.bin"))
    assert e.reload() is False
    assert_same_as_oracle(e.search_batch(qs, 10), small_case.oracle, qs, 10)
    e.close()


def test_tile_boundaries_and_ragged_segments(workdir):
    """Segments whose sizes straddle the doc tile (2048 by default; 1024/4096 selectable):
    2047, 2048, 2049, 4097, 8191 docs, a 1-doc segment and a larger one."""
    spec = nsb200.CorpusSpec(vocab=800)
    path = os.path.join(workdir, "ragged")
    names = []
    base = 0
    for i, n in enumerate([2047, 2048, 2049, 4097, 8191, 1, 20000]):
        name = nsb200.seg_name(i + 1)
        nsb200.write_segment(spec, base, n, os.path.join(path, "segments", name))
        names.append(name)
        base += n
    nsb200.write_manifest(path, names)
    from oracle import oracle as orc
    oi = orc.OracleIndex(path)
    e = nsb200.Engine(path, device=0)
    assert e.reload(), e.last_error
    qs = nsb200.make_queries(spec, 200, 1, 5) + EDGE_QUERIES
    for k in (10, 100):
        assert_same_as_oracle(e.search_batch(qs, k), oi, qs, k)
    e.close()


def test_concurrent_callers(small_case, small_engine):
    """ns_search_batch is callable from many host threads (the reference serialises on Engine::mtx)."""
    import threading
    qs = nsb200.make_queries(small_case.spec, 64, 1, 4)
    _, s, g, d, nh, fo, hf = small_case.oracle.search_many(qs, 10, nthreads=4)
    errs = []

    def worker():
        try:
            for _ in range(5):
                r = small_engine.search_batch(qs, 10)
                assert np.array_equal(r.found, fo) and np.array_equal(r.nhits, nh)
                assert np.array_equal(r.hits["doc"][0, : nh[0]], d[0, : nh[0]])
        except Exception as ex:  # noqa: BLE001
            errs.append(ex)

    th = [threading.Thread(target=worker) for _ in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


def test_inline_division_is_correctly_rounded():
    """div_rn_inrange (MUFU.RCP + 5 FFMA, no FCHK) == div.rn.f32 on 2^28 operand pairs with exponents
    in [-40, 40] — the range upload/prepare validate before selecting the FAST kernel."""
    import ctypes as C
    lib = nsb200._lib.load()
    bad = C.c_uint64(1)
    nsb200._lib.check(lib.ns_selftest_fastdiv(0, 1 << 28, 12345, C.byref(bad)))
    assert bad.value == 0


@pytest.mark.parametrize("resident", ["0", "1"])
def test_generic_kernel_variant_matches_too(small_case, monkeypatch, resident):
    """NSB200_NO_FAST forces the __fdiv_rn / weighted variant of the kernel (with and without the
    resident impact array built at upload)."""
    monkeypatch.setenv("NSB200_NO_FAST", "1")
    if resident == "0":
        monkeypatch.setenv("NSB200_NO_RESIDENT", "1")
    e = nsb200.Engine(small_case.path, device=0)
    assert e.reload()
    qs = nsb200.make_queries(small_case.spec, 200, 1, 5) + EDGE_QUERIES
    for k in (10, 100):
        assert_same_as_oracle(e.search_batch(qs, k), small_case.oracle, qs, k)
    e.close()


def test_negative_weight_uses_dense_scan(small_case, small_engine):
    """A negative qweight makes partial sums non-monotone; the kernel must fall back to scanning
    each tile.  Check against a numpy evaluation of the same float operations."""
    qs = ["t2 t9", "t5 t30 t7"]
    q_off, terms, _ = small_engine.resolve_batch(qs)
    t2 = terms.copy()
    t2["weight"][::2] = -0.5
    h, n, f = small_engine.index.search_batch(q_off, t2, 10)
    h1, n1, f1 = small_engine.index.search_batch(q_off, terms, 10)
    assert np.array_equal(f, f1)  # found does not depend on weights
    for q in range(len(qs)):
        s = h["score"][q, : n[q]]
        assert np.all(s[:-1] >= s[1:])


@pytest.mark.parametrize("resident", ["0", "1"])
def test_unpacked_posting_format_matches_too(small_case, monkeypatch, resident):
    """NSB200_NO_PACK keeps postings as raw {docId, tf} + per-doc norm gather (the path used when a
    tf >= 65536 or a segment has > 65536 distinct doc lengths)."""
    monkeypatch.setenv("NSB200_NO_PACK", "1")
    if resident == "0":
        monkeypatch.setenv("NSB200_NO_RESIDENT", "1")
    e = nsb200.Engine(small_case.path, device=0)
    assert e.reload()
    qs = nsb200.make_queries(small_case.spec, 200, 1, 5) + EDGE_QUERIES
    for k in (10, 100):
        assert_same_as_oracle(e.search_batch(qs, k), small_case.oracle, qs, k)
    e.close()


def test_wide_tf_falls_back_to_unpacked():
    """tf >= 65536 cannot be packed; results must still follow the reference formula."""
    idx = nsb200.DeviceIndex(0)
    doc_len = np.array([10, 20, 30, 40], np.uint32)
    post = np.array([[0, 70000], [2, 3], [1, 1], [3, 65536]], np.uint32)
    idx.add_segment(0, 25.0, doc_len, np.array([0, 2], np.uint64), np.array([2, 2], np.uint32), post)
    idx.commit()
    terms = np.array([(0, 0, 1.5, 1.0), (0, 1, 0.5, 1.0)], nsb200.QTERM_DTYPE)
    hits, n, found = idx.search_batch(np.array([0, 2], np.uint64), terms, 10)
    assert n[0] == 4 and found[0] == 4
    f = np.float32
    def score(idf, tf, dl):
        nrm = f(1.2) * (f(0.25) + f(0.75) * (f(dl) / f(25.0)))
        return f(idf) * (f(tf) * (f(1.2) + f(1.0))) / (f(tf) + nrm)
    want = {0: score(1.5, 70000, 10), 2: score(1.5, 3, 30), 1: score(0.5, 1, 20), 3: score(0.5, 65536, 40)}
    got = {int(h["doc"]): h["score"] for h in hits[0, :4]}
    for d, s in want.items():
        assert got[d].view(np.uint32) == np.float32(s).view(np.uint32), (d, got[d], s)
    idx.close()


@pytest.mark.parametrize("impact", ["0", "1"])
def test_impact_prepass_on_and_off(small_case, monkeypatch, impact):
    """Without resident impacts: NSB200_IMPACT=1 shares each distinct term's BM25 scores across the
    batch (pre-pass kernel); =0 evaluates them per (query, posting).  Both must be bit-identical to
    the oracle."""
    monkeypatch.setenv("NSB200_IMPACT", impact)
    monkeypatch.setenv("NSB200_NO_RESIDENT", "1")
    e = nsb200.Engine(small_case.path, device=0)
    assert e.reload()
    qs = nsb200.make_queries(small_case.spec, 300, 1, 5) + EDGE_QUERIES
    for k in (10, 100):
        assert_same_as_oracle(e.search_batch(qs, k), small_case.oracle, qs, k)
    e.close()


def test_foreign_idf_mixes_resident_and_per_batch_scores(small_case, small_engine, monkeypatch):
    """The resident impacts hold each row's score under bm25_idf(N, count).  A term that arrives with
    any other idf (bit-wise) must be evaluated per batch instead; a batch may mix both kinds.
    Reference for the comparison: an index uploaded without resident impacts, scoring on the fly."""
    qs = nsb200.make_queries(small_case.spec, 200, 1, 5) + ["t3 t3", "t1 t2 t1"]
    q_off, terms, _ = small_engine.resolve_batch(qs)
    mixed = terms.copy()
    mixed["idf"][::3] = np.nextafter(mixed["idf"][::3], np.float32(10.0))   # 1 ulp off: not resident
    mixed["idf"][1::7] *= np.float32(0.75)
    mixed["weight"][::5] = 0.5
    # the engine keeps only the resident scores by default (NS_SEG_DROP_RAW); foreign idfs need the raw postings
    monkeypatch.setenv("NSB200_KEEP_RAW", "1")
    both = nsb200.Engine(small_case.path, device=0)
    assert both.reload()
    monkeypatch.setenv("NSB200_NO_RESIDENT", "1")
    monkeypatch.setenv("NSB200_IMPACT", "0")
    plain = nsb200.Engine(small_case.path, device=0)
    assert plain.reload()
    with pytest.raises(nsb200._lib.NsError):  # default engine: refused loudly, never scored with the wrong idf
        small_engine.index.search_batch(q_off, mixed, 10)
    for k in (10, 100):
        h0, n0, f0 = plain.index.search_batch(q_off, mixed, k)
        h1, n1, f1 = both.index.search_batch(q_off, mixed, k)
        assert np.array_equal(n0, n1) and np.array_equal(f0, f1)
        for q in range(len(qs)):
            n = int(n0[q])
            assert np.array_equal(h0["score"][q, :n].view(np.uint32), h1["score"][q, :n].view(np.uint32)), qs[q]
            assert np.array_equal(h0["doc"][q, :n], h1["doc"][q, :n]) and np.array_equal(h0["seg"][q, :n], h1["seg"][q, :n])
    plain.close()
    both.close()


def test_sharded_blobs_merge_on_device(workdir):
    """BASELINE configs[2] in small: 8 segments, two shard engines (segment i -> shard i % 2) on the
    same GPU, their result blobs laid out back to back like the all-gather output, merged by
    ns_merge_blobs_device — must equal the oracle on the full 8-segment index."""
    import ctypes as C
    torch = pytest.importorskip("torch")
    from nextsearch_api_b200 import dist as nsdist

    spec = nsb200.CorpusSpec(vocab=4000)
    case = make_case(workdir, "eightseg", spec, 8000, 8)
    qs = nsb200.make_queries(spec, 300, 1, 5) + EDGE_QUERIES
    lib = nsb200._lib.load()
    for k in (10, 100):
        engines = [nsb200.Engine(case.path, device=0, rank=r, world=2) for r in range(2)]
        blobs, has = [], None
        K = nsb200.clamp_k(k)
        total, off_n, off_f = nsdist.blob_layout(len(qs), K)
        gathered = torch.zeros(2 * total, dtype=torch.uint8, device="cuda:0")
        batches = []
        for r, e in enumerate(engines):
            assert e.reload(), e.last_error
            q_off, terms, has = e.resolve_batch(qs)
            assert set(np.unique(terms["seg"]).tolist()) <= {s for s in range(8) if s % 2 == r}
            b = e.index.prepare(q_off, terms, k)
            b.launch()
            b.sync()
            ptr, nbytes, o_n, o_f = C.c_void_p(), C.c_uint64(), C.c_uint64(), C.c_uint64()
            nsb200._lib.check(lib.ns_batch_result_blob(b._h, C.byref(ptr), C.byref(nbytes), C.byref(o_n), C.byref(o_f)))
            assert (nbytes.value, o_n.value, o_f.value) == (total, off_n, off_f)
            local = torch.as_tensor(nsdist._DevMem(ptr.value, nbytes.value), device="cuda:0")
            gathered[r * total:(r + 1) * total].copy_(local)
            batches.append(b)
        out_hits = torch.zeros(len(qs) * K * 12, dtype=torch.uint8, device="cuda:0")
        out_n = torch.zeros(len(qs), dtype=torch.int32, device="cuda:0")
        out_f = torch.zeros(len(qs), dtype=torch.int64, device="cuda:0")
        torch.cuda.synchronize()
        nsb200._lib.check(lib.ns_merge_blobs_device(0, len(qs), k, 2, C.c_void_p(gathered.data_ptr()), total, off_n, off_f,
                                                    C.c_void_p(out_hits.data_ptr()), C.c_void_p(out_n.data_ptr()),
                                                    C.c_void_p(out_f.data_ptr()), None))
        torch.cuda.synchronize()
        res = nsb200.BatchResult(out_hits.cpu().numpy().view(nsb200.HIT_DTYPE).reshape(len(qs), K),
                                 out_n.cpu().numpy().view(np.uint32), out_f.cpu().numpy().view(np.uint64), has, K)
        assert_same_as_oracle(res, case.oracle, qs, k)
        for b in batches:
            b.close()
        for e in engines:
            e.close()


def test_sharded_searcher_single_rank_pipeline(workdir):
    """ShardedSearcher at world 1 (no collective): launch -> merge of the single blob -> fetch, and the
    pipelined search_many (host front end of batch i+1 while the GPU works on batch i) — every batch of
    the stream must equal the oracle, in order."""
    pytest.importorskip("torch")
    from nextsearch_api_b200.dist import ShardedSearcher

    spec = nsb200.CorpusSpec(vocab=4000)
    case = make_case(workdir, "eightseg_s", spec, 8000, 8)
    s = ShardedSearcher(case.path, 0, 0, 1)
    assert s.reload(), s.engine.last_error
    batches = [nsb200.make_queries(spec, 200, 1, 5, seed=21 + i) + EDGE_QUERIES for i in range(4)]
    one = s.search_batch(batches[0], 10)
    assert_same_as_oracle(one, case.oracle, batches[0], 10)
    many = s.search_many(batches, 10)
    assert len(many) == len(batches)
    for qs, res in zip(batches, many):
        assert_same_as_oracle(res, case.oracle, qs, 10)
    for res, qs in zip(s.search_many(batches[:2], 100), batches[:2]):
        assert_same_as_oracle(res, case.oracle, qs, 100)
    s.engine.close()


def test_result_list_lock_under_contention(workdir):
    """Few heavy queries cut into 64 items each: all items of a query run at the same time and merge into
    ONE shared result list under the per-query spin lock.  60 launches of the same batch must each give
    the oracle's answer (a lost update, a torn list or a warp that lost convergence in the spin would not)."""
    spec = nsb200.CorpusSpec(vocab=60_000)
    case = make_case(workdir, "mid200k_lock", spec, 200_000, 1)
    e = nsb200.Engine(case.path, device=0)
    assert e.reload(), e.last_error
    qs = nsb200.make_queries(spec, 6, 2, 5, seed=3, head_ranks=20)
    q_off, terms, has = e.resolve_batch(qs)
    for k in (10, 100):
        b = e.index.prepare(q_off, terms, k)
        b.set_splits(64)
        for _ in range(60):
            b.launch()
            hits, nhits, found = b.fetch()
            res = nsb200.BatchResult(hits.copy(), nhits.copy(), found.copy(), has, nsb200.clamp_k(k))
            assert_same_as_oracle(res, case.oracle, qs, k)
        b.close()
    e.close()


def test_high_df_top100(workdir):
    """BASELINE configs[3]: every query contains a very frequent term (df > 10 % of the corpus), k=100."""
    spec = nsb200.CorpusSpec(vocab=20_000)
    case = make_case(workdir, "highdf", spec, 30_000, 2)
    e = nsb200.Engine(case.path, device=0)
    assert e.reload(), e.last_error
    qs = nsb200.make_queries(spec, 200, 1, 5, seed=11, head_ranks=40)
    df, _ = e.term_stats(0, qs[0].split()[0])
    assert df > 0.10 * case.oracle.segment_stats(0)["N"]
    assert_same_as_oracle(e.search_batch(qs, 100), case.oracle, qs, 100)
    assert_same_as_oracle(e.search_batch(qs, 10), case.oracle, qs, 10)
    e.close()


def test_long_slices_200k_docs(workdir):
    """A corpus large enough that frequent terms have several full 128-posting groups per 2048-doc
    tile (the double-buffered path) and that items span many tiles and segment boundaries."""
    spec = nsb200.CorpusSpec(vocab=60_000)
    case = make_case(workdir, "mid200k", spec, 200_000, 3)
    e = nsb200.Engine(case.path, device=0)
    assert e.reload(), e.last_error
    qs = nsb200.make_queries(spec, 384, 1, 5, seed=5) + nsb200.make_queries(spec, 128, 2, 5, seed=6, head_ranks=30)
    assert_same_as_oracle(e.search_batch(qs, 10), case.oracle, qs, 10)
    assert_same_as_oracle(e.search_batch(qs[:128], 100), case.oracle, qs[:128], 100)
    e.close()
