"""Hand-derived BM25 values (formulas at src/api_engine.cpp:45-47 and :477-480) against the oracle,
evaluated here with numpy float32 one operation at a time."""
import os

import numpy as np
import pytest

import fmt
from oracle import oracle as orc

import ctypes

f = np.float32
_libm = ctypes.CDLL("libm.so.6")
_libm.logf.restype = ctypes.c_float
_libm.logf.argtypes = [ctypes.c_float]


def idf(N, df):
    """bm25_idf: the reference calls std::log(float) = glibc logf (numpy's own float32 log may differ
    in the last bit, so call the same libm)."""
    return f(_libm.logf(f(((f(np.uint32((N - df) & 0xFFFFFFFF)) + f(0.5)) / (f(df) + f(0.5))) + f(1.0))))


def term_score(idf_v, tf, dl, avgdl):
    k1, b = f(1.2), f(0.75)
    denom = f(tf) + k1 * (f(1.0) - b + b * (f(dl) / f(avgdl)))
    return f(idf_v) * (f(tf) * (k1 + f(1.0))) / denom


@pytest.fixture(scope="module")
def hand(workdir):
    idx = os.path.join(workdir, "known_answer")
    docs = fmt.handmade_docs()
    fmt.write_segment(os.path.join(idx, "segments", "seg_000001"), docs)
    fmt.write_manifest(idx, ["seg_000001"])
    return orc.OracleIndex(idx), docs


def test_stats_and_avgdl(hand):
    oi, docs = hand
    st = oi.segment_stats(0)
    assert st["N"] == 12 and st["T"] == 8
    assert f(st["avgdl"]) == f(sum(d[1] for d in docs)) / f(12)
    assert oi.term_stats(0, "alpha") == (6, 6) and oi.term_stats(0, "nosuch") == (0, 0)


def test_single_term_scores_and_order(hand):
    oi, docs = hand
    avgdl = f(oi.segment_stats(0)["avgdl"])
    r = oi.search("alpha", 10)
    want = []
    for doc_id, (_, dl, tfs) in enumerate(docs):
        for t, tf in tfs:
            if t == "alpha":
                want.append((term_score(idf(12, 6), tf, dl, avgdl), doc_id))
    want.sort(key=lambda x: (-x[0], x[1]))  # score desc, docId asc
    assert r["found"] == 6
    assert [(h["docId"]) for h in r["results"]] == [d for _, d in want]
    assert [h["score_bits"] for h in r["results"]] == [int(f(s).view(np.uint32)) for s, _ in want]


def test_multi_term_accumulates_in_query_order(hand):
    oi, docs = hand
    avgdl = f(oi.segment_stats(0)["avgdl"])
    df = {"alpha": 6, "beta": 5, "gamma": 4}
    for q in (["alpha", "beta", "gamma"], ["gamma", "alpha", "beta"]):
        r = oi.search(" ".join(q), 10)
        want = {}
        for t in q:  # term at a time, in query order: score[doc] += 1.0f * s
            for doc_id, (_, dl, tfs) in enumerate(docs):
                for tt, tf in tfs:
                    if tt == t:
                        want[doc_id] = f(want.get(doc_id, f(0.0)) + f(1.0) * term_score(idf(12, df[t]), tf, dl, avgdl))
        assert r["found"] == len(want)
        for h in r["results"]:
            assert h["score_bits"] == int(f(want[h["docId"]]).view(np.uint32)), (q, h)


def test_duplicate_terms_score_twice(hand):
    oi, _ = hand
    one = {h["docId"]: h["score"] for h in oi.search("eps", 10)["results"]}
    three = {h["docId"]: h["score"] for h in oi.search("eps eps eps", 10)["results"]}
    for d, s in one.items():
        assert f(three[d]) == f(f(f(0.0) + f(s)) + f(s)) + f(s)


def test_found_key_rules_and_k_clamp(hand):
    oi, _ = hand
    assert oi.search("the of and", 10)["found"] is None      # all stopwords: key omitted
    assert oi.search("", 10)["found"] is None
    assert oi.search("x y", 10)["found"] is None              # len < 2
    assert oi.search("nosuchterm", 10)["found"] == 0          # usable term, no match
    assert oi.search("alpha", 1000)["k"] == 100 and oi.search("alpha", 0)["k"] == 1
    assert len(oi.search("alpha", 2)["results"]) == 2
