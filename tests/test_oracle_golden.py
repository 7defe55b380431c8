"""The oracle restatement (oracle/bm25_oracle.c) against outputs of the reference itself
(tests/golden/*.json, produced by tests/golden/make_golden.py from oracle/_ref/ref_engine)."""
import hashlib
import json
import os

import numpy as np
import pytest

import fmt
import nsb200
from oracle import oracle as orc
from refcmp import check_against_reference, golden_result, oracle_result

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return json.load(open(os.path.join(GOLD, name)))


def sha_dir(d):
    return {f: hashlib.sha256(open(os.path.join(d, f), "rb").read()).hexdigest() for f in sorted(os.listdir(d))}


@pytest.fixture(scope="module")
def gold_small():
    return load("small_index.json")


def test_product_writer_matches_reference_segment_writer_bytes(workdir, gold_small):
    """ns_corpus_write_segment == SegmentWriter::write_segment (include/segment_writer.hpp:65-168),
    all 133 files per segment, via the sha256 the reference-written files had."""
    path = os.path.join(workdir, "gold_small_fwd")
    spec = nsb200.CorpusSpec(**{k: gold_small["spec"][k] for k in ("vocab", "seed", "zipf_s", "zipf_q", "len_lo", "len_hi")})
    names = nsb200.build_index(path, spec, gold_small["ndocs"], gold_small["nseg"], write_forward=True)
    for name in names:
        assert sha_dir(os.path.join(path, "segments", name)) == gold_small["segment_sha256"][name]
    got = hashlib.sha256(open(os.path.join(path, "manifest.bin"), "rb").read()).hexdigest()
    assert got == gold_small["manifest_sha256"]  # save_manifest, src/api_segment.cpp:29-35


def test_python_format_writer_matches_reference_bytes(workdir):
    gold = load("handmade.json")
    seg = os.path.join(workdir, "hand_fmt", "segments", "seg_000001")
    fmt.write_segment(seg, fmt.handmade_docs())
    assert sha_dir(seg) == gold["segment_sha256"]
    ties = load("ties.json")
    for s in range(2):
        seg = os.path.join(workdir, "ties_fmt", "segments", nsb200.seg_name(s + 1))
        fmt.write_segment(seg, fmt.tie_docs(1000, s * 1000))
        assert sha_dir(seg) == ties["segment_sha256"][nsb200.seg_name(s + 1)]


@pytest.mark.parametrize("k", ["10", "1", "3", "100", "1000", "0"])
def test_oracle_matches_reference_search_generated_corpus(small_case, gold_small, k):
    """found, score bits, tie-aware doc sets — Engine::search as shipped vs the restatement."""
    oi = small_case.oracle
    seg_index = {oi.segment_name(i): i for i in range(oi.num_segments)}
    for row in gold_small["search"][k]:
        ref = golden_result(row)
        ours = oracle_result(oi, row["query"], int(k))
        assert row["segments"] == oi.num_segments
        check_against_reference(ours, ref, oi, seg_index)
        # cord_uid of every returned doc
        for h in row["hits"]:
            assert oi.cord_uid(seg_index[h[0]], h[1]) == h[3]


def test_oracle_matches_reference_handmade_and_legacy(workdir):
    gold = load("handmade.json")
    for legacy, key in ((False, None), (True, "search_legacy_10")):
        idx = os.path.join(workdir, f"hand_idx_{int(legacy)}")
        fmt.write_segment(os.path.join(idx, "segments", "seg_000001"), fmt.handmade_docs(), legacy=legacy)
        fmt.write_manifest(idx, ["seg_000001"])
        oi = orc.OracleIndex(idx)
        runs = {"10": gold["search_legacy_10"]} if legacy else gold["search"]
        for k, rows in runs.items():
            for row in rows:
                check_against_reference(oracle_result(oi, row["query"], int(k)), golden_result(row), oi, {"seg_000001": 0})


def test_oracle_matches_reference_on_massive_ties(workdir):
    """2 x 1000 identical docs: found = 2000, every score identical, and only the EARLIER segment
    appears (strict '>' replacement, src/api_engine.cpp:488) — the stated total order reproduces that."""
    gold = load("ties.json")
    idx = os.path.join(workdir, "ties_idx")
    for s in range(2):
        fmt.write_segment(os.path.join(idx, "segments", nsb200.seg_name(s + 1)), fmt.tie_docs(1000, s * 1000))
    fmt.write_manifest(idx, [nsb200.seg_name(1), nsb200.seg_name(2)])
    oi = orc.OracleIndex(idx)
    seg_index = {nsb200.seg_name(1): 0, nsb200.seg_name(2): 1}
    for k, rows in gold["search"].items():
        for row in rows:
            ours = oracle_result(oi, row["query"], int(k))
            check_against_reference(ours, golden_result(row), oi, seg_index)
            assert row["found"] == 2000
            assert {h[0] for h in row["hits"]} == {"seg_000001"}          # the reference
            assert [h[0] for h in ours["hits"]] == ["seg_000001"] * len(ours["hits"])
            assert [h[1] for h in ours["hits"]] == list(range(len(ours["hits"])))  # docId asc on ties


@pytest.mark.ref
def test_oracle_matches_live_reference(config1_case):
    """BASELINE configs[0] against the reference binary itself (only where /root/reference exists)."""
    qs = nsb200.make_queries(config1_case.spec, 200, 1, 3, seed=nsb200.QUERY_SEED)
    _, res = orc.ref_search(config1_case.path, qs, 10)
    oi = config1_case.oracle
    for q, r in zip(qs, res):
        ref = {"query": q, "found": r.get("found"), "k": r["k"],
               "hits": [(h["segment"], h["docId"], h["score_bits"]) for h in r["results"]]}
        check_against_reference(oracle_result(oi, q, 10), ref, oi, {"seg_000001": 0})


@pytest.mark.ref
@pytest.mark.parametrize("vocab,seed,base,ndocs", [
    (1500, 99, 700, 900),
    (10, 3, 0, 40),          # T < 64: terms_per_barrel = 1, most barrels empty (include/barrels.hpp:26-30)
    (64, 4, 5, 17),          # T == barrel count
    (65, 5, 0, 1),           # a single document
    (5000, 6, 123456, 300),  # T > 64 with a ragged last barrel
])
def test_product_writer_matches_live_reference_writer(workdir, vocab, seed, base, ndocs):
    spec = nsb200.CorpusSpec(vocab=vocab, seed=seed)
    mine = os.path.join(workdir, f"live_mine_{vocab}_{seed}")
    dump = os.path.join(workdir, f"live_dump_{vocab}_{seed}.bin")
    nsb200.write_segment(spec, base, ndocs, mine, True, dump)
    ref = os.path.join(workdir, f"live_ref_{vocab}_{seed}")
    orc.ref_write_segment(dump, ref)
    assert sha_dir(mine) == sha_dir(ref)


@pytest.mark.ref
def test_oracle_matches_live_reference_on_messy_query_strings(small_case):
    """Tokenisation and filtering (include/textutil.hpp:13-37, src/api_engine.cpp:388-397) through the whole search:
    random mixtures of case, punctuation, digits, stopwords, one-letter tokens, tabs and valid multi-byte UTF-8."""
    import random

    rng = random.Random(77)
    pieces = ["t1", "T2", "t3,", "(t4)", "t5-t6", "t7_t8", "the", "The", "a", "x", "of", "t9.", "t10!", "é", "t11é", "naïve", "t12\tt13",
              "  ", "t14;t15", "zzzz", "9", "42", "t16/t17", "IS", "t18's", "\"t19\"", "ｔ２０", "t21 t22", "an", "t23?", "#t24", "t25+t26"]
    queries = []
    for _ in range(150):
        q = " ".join(rng.choice(pieces) for _ in range(rng.randint(1, 7)))
        queries.append(q)
    queries = list(dict.fromkeys(queries))
    oi = small_case.oracle
    seg_index = {oi.segment_name(i): i for i in range(oi.num_segments)}
    _, res = orc.ref_search(small_case.path, queries, 10)
    assert len(res) == len(queries)
    for q, r in zip(queries, res):
        ref = {"query": q, "found": r.get("found"), "k": r["k"],
               "hits": [(h["segment"], h["docId"], h["score_bits"]) for h in r["results"]]}
        check_against_reference(oracle_result(oi, q, 10), ref, oi, seg_index)
        assert nsb200.query_terms(q) == orc.query_terms(q), q          # the product's tokenizer gives the oracle's terms
