"""One engine spanning several device slots (ns_engine_create_multi), the request coalescer, reload under load,
and the upload-time defences — through the C ABI.  Device slots may name the same GPU twice, which is how the
multi-device path (fan-out, peer publish, root merge) runs on a single-GPU box; test_real_two_gpus uses two."""
import os
import shutil
import threading

import numpy as np
import pytest

import nsb200
from conftest import EDGE_QUERIES, assert_same_as_oracle, make_case
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def five_seg_case(workdir):
    return make_case(workdir, "multi5", nsb200.CorpusSpec(vocab=5000), 10_000, 5)


@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0], [0]])
@pytest.mark.parametrize("k", [10, 100])
def test_multi_slot_engine_equals_oracle(five_seg_case, devices, k):
    eng = nsb200.Engine(five_seg_case.path, devices=devices)
    assert eng.reload(), eng.last_error
    queries = nsb200.make_queries(five_seg_case.spec, 400, 1, 5, seed=41) + EDGE_QUERIES
    for _ in range(3):  # the exchange group is reused: steps 0, 1, 2
        assert_same_as_oracle(eng.search_batch(queries, k), five_seg_case.oracle, queries, k)
    one = eng.search("t1 t2", 10)
    ref = five_seg_case.oracle.search("t1 t2", 10)
    assert one["found"] == ref["found"]
    assert [(r["segment"], r["docId"]) for r in one["results"]] == [(r["segment"], r["docId"]) for r in ref["results"]]
    st = eng.reload_stats()
    assert st["device_bytes"] <= 2.0 * st["posting_bytes"] + (1 << 20) * len(devices) * 5, st
    eng.close()


def test_real_two_gpus(five_seg_case):
    if nsb200._lib.load().ns_device_count() < 2:
        pytest.skip("needs two GPUs")
    eng = nsb200.Engine(five_seg_case.path, devices=[0, 1])
    assert eng.reload(), eng.last_error
    queries = nsb200.make_queries(five_seg_case.spec, 1000, 1, 5, seed=42) + EDGE_QUERIES
    for k in (10, 100):
        assert_same_as_oracle(eng.search_batch(queries, k), five_seg_case.oracle, queries, k)
    eng.close()


def test_multi_slot_engine_concurrent_callers(five_seg_case):
    eng = nsb200.Engine(five_seg_case.path, devices=[0, 0])
    assert eng.reload()
    batches = [nsb200.make_queries(five_seg_case.spec, 200, 1, 4, seed=50 + i) for i in range(6)]
    errors = []

    def worker(i):
        try:
            for _ in range(3):
                assert_same_as_oracle(eng.search_batch(batches[i], 10), five_seg_case.oracle, batches[i], 10)
        except Exception as ex:  # noqa: BLE001
            errors.append(repr(ex))

    ths = [threading.Thread(target=worker, args=(i,)) for i in range(6)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors[:2]
    eng.close()


def test_coalescer_batches_single_requests(five_seg_case):
    """64 threads each issue single queries (the reference's one engine.search per HTTP request,
    src/api_server.cpp:117-178); the dispatcher gathers them into GPU batches; every answer equals the oracle's."""
    eng = nsb200.Engine(five_seg_case.path, device=0)
    assert eng.reload()
    eng.coalescer_start(max_batch=256, max_wait_us=2000, dispatchers=2)
    queries = nsb200.make_queries(five_seg_case.spec, 64 * 8, 1, 4, seed=60)
    ks = [10, 3, 100, 1]
    got, errors = {}, []

    def worker(t):
        try:
            for j in range(8):
                i = t * 8 + j
                got[i] = eng.search_one(queries[i], ks[i % 4])
        except Exception as ex:  # noqa: BLE001
            errors.append(repr(ex))

    ths = [threading.Thread(target=worker, args=(t,)) for t in range(64)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors[:2]
    stats = eng.coalescer_stats()
    assert stats["queries"] == len(queries) and stats["max_batch"] > 1, stats
    oi = five_seg_case.oracle
    for i, q in enumerate(queries):
        want = oi.search(q, ks[i % 4])
        hits, found = got[i]
        assert found == want["found"], q
        assert hits["score"].view(np.uint32).tolist() == [h["score_bits"] for h in want["results"]], q
        assert hits["doc"].tolist() == [h["docId"] for h in want["results"]]
        assert hits["seg"].tolist() == [h["seg"] for h in want["results"]]
    # the JSON entry point goes through the same queue
    assert eng.search("t1 t2", 10)["found"] == oi.search("t1 t2", 10)["found"]
    eng.coalescer_stop()
    assert eng.search("t1 t2", 10)["found"] == oi.search("t1 t2", 10)["found"]  # direct path again
    eng.close()


def test_coalescer_isolates_a_refused_request(five_seg_case):
    """A request the device layer refuses (more than NS_MAX_TERMS terms for one segment) fails alone: the strangers
    that shared its batch get their answers."""
    eng = nsb200.Engine(five_seg_case.path, device=0)
    assert eng.reload()
    eng.coalescer_start(max_batch=64, max_wait_us=200000, dispatchers=1)   # long gather window: one batch for all
    too_long = " ".join(f"t{i}" for i in range(1, 400))
    queries = nsb200.make_queries(five_seg_case.spec, 15, 1, 4, seed=61)
    got, errors, refused = {}, [], []

    def worker(i):
        try:
            if i == 15:
                try:
                    eng.search_one(too_long, 10)
                except nsb200._lib.NsError as ex:
                    refused.append(ex.status)
            else:
                got[i] = eng.search_one(queries[i], 10)
        except Exception as ex:  # noqa: BLE001
            errors.append(repr(ex))

    ths = [threading.Thread(target=worker, args=(i,)) for i in range(16)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errors, errors[:2]
    assert refused == [1]                                   # NS_ERR_INVALID for the long query only
    oi = five_seg_case.oracle
    for i, q in enumerate(queries):
        want = oi.search(q, 10)
        hits, found = got[i]
        assert found == want["found"], q
        assert hits["score"].view(np.uint32).tolist() == [h["score_bits"] for h in want["results"]], q
        assert hits["doc"].tolist() == [h["docId"] for h in want["results"]]
    eng.coalescer_stop()
    eng.close()


def test_reload_swaps_generation_under_load(workdir):
    """Searches run while the index directory is switched to a DIFFERENT corpus and reloaded: every batch must
    equal the oracle on corpus A or on corpus B as a whole — never old lexicon rows against new device arrays."""
    a = make_case(workdir, "gen_a", nsb200.CorpusSpec(vocab=2500, seed=5), 6000, 3)
    b = make_case(workdir, "gen_b", nsb200.CorpusSpec(vocab=1800, seed=6), 4000, 2)
    link = os.path.join(workdir, "gen_cur")
    if os.path.lexists(link):
        os.remove(link)
    os.symlink(a.path, link)
    eng = nsb200.Engine(link, devices=[0, 0])
    assert eng.reload()
    queries = nsb200.make_queries(a.spec, 300, 1, 4, seed=70) + ["t1 t2", "t3"]

    def expect(case):
        _, s, g, d, nh, fo, hf = case.oracle.search_many(queries, 10, nthreads=4)
        return (s.view(np.uint32), g, d, nh, fo)

    want = {"a": expect(a), "b": expect(b)}
    stop, errors, seen = threading.Event(), [], set()

    def matches(res, w):
        s, g, d, nh, fo = w
        if not (np.array_equal(res.nhits, nh) and np.array_equal(res.found, fo)):
            return False
        for q in range(len(queries)):
            n = int(nh[q])
            if not (np.array_equal(res.hits["score"][q, :n].view(np.uint32), s[q, :n]) and np.array_equal(res.hits["doc"][q, :n], d[q, :n])
                    and np.array_equal(res.hits["seg"][q, :n], g[q, :n])):
                return False
        return True

    def searcher():
        try:
            while not stop.is_set():
                res = eng.search_batch(queries, 10)
                which = [nm for nm, w in want.items() if matches(res, w)]
                if not which:
                    errors.append("a batch matches neither generation")
                    return
                seen.update(which)
        except Exception as ex:  # noqa: BLE001
            errors.append(repr(ex))

    ths = [threading.Thread(target=searcher) for _ in range(4)]
    for t in ths:
        t.start()
    for target in (b, a, b):
        os.remove(link)
        os.symlink(target.path, link)
        assert eng.reload(), eng.last_error
    stop.set()
    for t in ths:
        t.join()
    assert not errors, errors[:2]
    assert "b" in seen
    assert matches(eng.search_batch(queries, 10), want["b"])
    eng.close()


def _tiny_segment(bad_doc=None, n=16):
    doc_len = np.full(n, 10, np.uint32)
    postings = np.array([[0, 1], [3, 2], [7, 1], [2, 1], [9, 3]], np.uint32)
    if bad_doc is not None:
        postings[2, 0] = bad_doc
    term_begin = np.array([0, 3], np.uint64)
    term_count = np.array([3, 2], np.uint32)
    return doc_len, term_begin, term_count, postings


@pytest.mark.parametrize("bad_doc", [0xFFFFFFF0, 1 << 30, 16])
def test_corrupt_segment_is_rejected_before_any_docid_indexed_kernel(bad_doc):
    """docId far outside N: NS_ERR_FORMAT, no illegal address, the previous index stays live and searchable."""
    idx = nsb200.DeviceIndex(0)
    doc_len, tb, tc, post = _tiny_segment()
    idx.add_segment(0, 10.0, doc_len, tb, tc, post)
    idx.commit()
    doc_len, tb, tc, bad = _tiny_segment(bad_doc)
    with pytest.raises(nsb200._lib.NsError) as ei:
        idx.add_segment(1, 10.0, doc_len, tb, tc, bad)
    assert ei.value.status == 4
    # row range wrapping in u64
    with pytest.raises(nsb200._lib.NsError) as ei:
        idx.add_segment(1, 10.0, doc_len, np.array([0, 0xFFFFFFFFFFFFFFFE], np.uint64), tc, post)
    assert ei.value.status == 4
    terms = np.array([(0, 0, 1.0, 1.0)], dtype=nsb200.QTERM_DTYPE)
    hits, nhits, found = idx.search_batch(np.array([0, 1], np.uint64), terms, 10)
    assert int(found[0]) == 3 and int(nhits[0]) == 3  # the CUDA context is alive, the old index answers
    idx.close()


def test_drop_raw_segment_refuses_foreign_idf():
    lib = nsb200._lib.load()
    import ctypes as C
    idx = nsb200.DeviceIndex(0)
    doc_len, tb, tc, post = _tiny_segment()
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    post = np.ascontiguousarray(post)
    row_idf = np.array([1.5, 2.5], np.float32)  # the caller's idf per row: what the resident scores are built with
    nsb200._lib.check(lib.ns_index_add_segment_ex(idx._h, 0, 16, C.c_float(10.0), P(doc_len), 2, P(tb), P(tc), P(row_idf), P(post),
                                                   5, nsb200._lib.NS_SEG_DROP_RAW))
    idx.commit()
    ok_terms = np.array([(0, 0, 1.5, 1.0)], dtype=nsb200.QTERM_DTYPE)
    hits, nhits, found = idx.search_batch(np.array([0, 1], np.uint64), ok_terms, 10)
    assert int(found[0]) == 3
    with pytest.raises(nsb200._lib.NsError) as ei:
        idx.search_batch(np.array([0, 1], np.uint64), np.array([(0, 0, 0.123, 1.0)], dtype=nsb200.QTERM_DTYPE), 10)
    assert ei.value.status == 6
    idx.close()


def test_negative_avgdl_segment_uses_dense_selection(workdir):
    """stats.bin with avgdl < 0 (the reference reads it verbatim, src/api_segment.cpp:110-115): the doc-length factors
    and with them some term scores are negative, partial sums are not monotone, so the threshold-crossing shortcut
    must not be used.  Results still equal the oracle's, bit for bit."""
    import struct

    import fmt
    idx = os.path.join(workdir, "neg_avgdl")
    seg = os.path.join(idx, "segments", "seg_000001")
    fmt.write_segment(seg, fmt.semantic_docs(0) + fmt.semantic_docs(1))
    fmt.write_manifest(idx, ["seg_000001"])
    with open(os.path.join(seg, "stats.bin"), "r+b") as f:
        n = struct.unpack("<I", f.read(4))[0]
        f.seek(0)
        f.write(struct.pack("<If", n, -37.5))
    eng = nsb200.Engine(idx, device=0)
    assert eng.reload(), eng.last_error
    oi = orc.OracleIndex(idx)
    queries = ["virus", "covid vaccine", "bat virus spike", "fever cough rna lung", "mask masks ppe"]
    for k in (3, 10, 100):
        assert_same_as_oracle(eng.search_batch(queries, k), oi, queries, k)
    eng.close()
