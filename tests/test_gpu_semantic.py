"""GPU side of tests/test_semantic_golden.py: weighted scoring through the C ABI against the oracle, and the
product's JSON text against the reference's own j.dump() (tests/golden/semantic.json)."""
import json
import os

import numpy as np
import pytest

import fmt
import nsb200
from oracle import oracle as orc
from test_semantic_golden import GOLD, build_sem_index, qterms_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sem_index(workdir):
    return build_sem_index(os.path.join(workdir, "sem_idx_gpu"))


@pytest.fixture(scope="module")
def plain_index(workdir):
    return build_sem_index(os.path.join(workdir, "sem_idx_gpu_plain"), embeddings=False)


@pytest.mark.parametrize("k", [10, 100, 3])
def test_weighted_terms_bit_exact_vs_oracle(plain_index, k):
    """qweight != 1 and up to 20 terms per query: CUDA (ns_engine_search_terms_batch) vs the oracle's weighted
    entry, which tests/test_semantic_golden.py pins to the reference."""
    eng = nsb200.Engine(plain_index, device=0)
    assert eng.reload(), eng.last_error
    oi = orc.OracleIndex(plain_index)
    lists = [qterms_of(r) for r in GOLD["expanded"]["10"]]
    lists.append([("virus", 0.25), ("virus", 0.5), ("nosuchterm", 1.0), ("bats", 2.0)])   # duplicates, unknown, weight > 1
    lists.append([])                                                                       # no terms: no "found"
    res = eng.search_terms_batch(lists, k)
    for q, lst in enumerate(lists):
        want = oi.search_weighted(lst, k)
        n = len(want["results"])
        assert int(res.nhits[q]) == n
        assert bool(res.has_found[q]) == (want["found"] is not None)
        if want["found"] is not None:
            assert int(res.found[q]) == want["found"]
        assert res.hits["score"][q, :n].view(np.uint32).tolist() == [h["score_bits"] for h in want["results"]], lst
        assert res.hits["seg"][q, :n].tolist() == [h["seg"] for h in want["results"]]
        assert res.hits["doc"][q, :n].tolist() == [h["docId"] for h in want["results"]]
    eng.close()


@pytest.mark.parametrize("k", ["10", "100"])
def test_expanded_search_json_text_equals_reference_dump(sem_index, k):
    """Engine::search with embeddings loaded, end to end: expansion, weighted scoring on the GPU, decoration,
    serialisation — the text is the reference's, byte for byte."""
    eng = nsb200.Engine(sem_index, device=0)
    assert eng.reload(), eng.last_error
    for row in GOLD["expanded"][k]:
        assert eng.search_json_text(row["query"], int(k)) == row["text"], row["query"]
    eng.close()


@pytest.mark.parametrize("k", ["10", "3"])
def test_plain_search_json_text_equals_reference_dump(plain_index, k):
    eng = nsb200.Engine(plain_index, device=0)
    assert eng.reload(), eng.last_error
    for row in GOLD["plain"][k]:
        assert eng.search_json_text(row["query"], int(k)) == row["text"], row["query"]
    eng.close()


def test_handmade_json_text_equals_reference_dump(workdir):
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "handmade.json")))
    idx = os.path.join(workdir, "hand_idx_json")
    fmt.write_segment(os.path.join(idx, "segments", "seg_000001"), fmt.handmade_docs())
    fmt.write_manifest(idx, ["seg_000001"])
    eng = nsb200.Engine(idx, device=0)
    assert eng.reload()
    oi = orc.OracleIndex(idx)
    for row in gold["search"]["10"]:
        bits = [h[2] for h in row["hits"]]
        if len(set(bits)) != len(bits):
            continue  # the reference's order inside a tie group is a hash-map artefact: not a text-level target
        assert eng.search_json_text(row["query"], 10) == row["text"], row["query"]
    eng.close()


def test_invalid_utf8_query_is_refused_like_the_reference_throws(plain_index):
    eng = nsb200.Engine(plain_index, device=0)
    assert eng.reload()
    with pytest.raises(nsb200._lib.NsError) as ei:
        eng.search_json_text(b"virus \xff\xfe", 10)
    assert ei.value.status == 1 and "UTF-8" in str(ei.value)
    eng.close()


@pytest.mark.skipif(not orc.have_shim(), reason="oracle/_ref/shim_engine not built (needs /root/reference at build time)")
@pytest.mark.parametrize("mode,k", [("expanded", 10), ("plain", 10), ("plain", 3)])
def test_compiled_reference_side_shim_returns_the_reference_text(sem_index, plain_index, mode, k):
    """INTEGRATION.md §1 as a real build: cord19::Engine::reload/search compiled against the REFERENCE's own
    include/api_engine.hpp, delegating to the C ABI.  The nlohmann::json it returns, dumped by the reference's json
    library, equals what the unmodified reference produced for the same index and queries."""
    rows = GOLD[mode][str(k)]
    texts = orc.shim_search(sem_index if mode == "expanded" else plain_index, [r["query"] for r in rows], k)
    assert len(texts) == len(rows)
    for row, text in zip(rows, texts):
        assert text == row["text"], row["query"]


def _write_random_embeddings(path, vocab, dim, nclusters, seed):
    """Clustered vectors for terms t1..t<vocab>: members of a cluster are close (cosine ~0.7-0.95), clusters are far apart."""
    rng = np.random.default_rng(seed)
    centres = rng.normal(size=(nclusters, dim))
    with open(path, "w", newline="") as f:
        f.write(f"{vocab} {dim}\n")
        for r in range(1, vocab + 1):
            v = centres[r % nclusters] + rng.normal(scale=0.45, size=dim)
            f.write(f"t{r} " + " ".join(f"{x:.5f}" for x in v) + "\n")


def test_device_similarity_scan_gives_the_host_expansion_bit_for_bit(workdir):
    """the expansion's similarity scans on the GPU (cosine_scan_kernel).  Same corpus and embeddings,
    one engine with the scans on the device, one host-only engine with the host loop (which
    tests/test_semantic_golden.py pins to the reference's expand): terms, weight bits and order must be identical,
    and so must the searches that score those lists (CUDA vs the oracle's weighted entry)."""
    from conftest import make_case
    case = make_case(workdir, "sem_rand", nsb200.CorpusSpec(vocab=2500), 4000, 2)
    emb = os.path.join(case.path, "embeddings.vec")
    if not os.path.exists(emb):
        _write_random_embeddings(emb, 2400, 24, 160, 5)
    gpu = nsb200.Engine(case.path, device=0)
    assert gpu.reload(), gpu.last_error
    host = nsb200.Engine(case.path, device=None)
    assert host.reload(), host.last_error
    queries = nsb200.make_queries(case.spec, 150, 1, 4, seed=91) + ["t1", "t7 t7", "t2400 t3", "nosuch t5"]
    expanded = 0
    lists = []
    for q in queries:
        a, b = gpu.expand(q), host.expand(q)
        assert a is not None and b is not None
        assert [(t, np.float32(w).view(np.uint32)) for t, w in a] == [(t, np.float32(w).view(np.uint32)) for t, w in b], q
        expanded += len(a) > len(set(q.split()))
        lists.append(a)
    assert expanded > 20  # the fixture really produces neighbours
    oi = orc.OracleIndex(case.path)
    res = gpu.search_batch(queries, 10)  # expansion (device scans) + weighted scoring, end to end
    for i, q in enumerate(queries):
        want = oi.search_weighted(lists[i], 10)
        n = len(want["results"])
        assert int(res.nhits[i]) == n and int(res.found[i]) == (want["found"] or 0), q
        assert res.hits["score"][i, :n].view(np.uint32).tolist() == [h["score_bits"] for h in want["results"]], q
        assert res.hits["doc"][i, :n].tolist() == [h["docId"] for h in want["results"]], q
    gpu.close()
    host.close()
