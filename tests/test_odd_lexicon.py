"""Lexicon entries no writer of the reference produces but its loader and scoring loop accept (tests/fmt.py
ODD_ENTRIES): df != count, df == 0 with postings, df > N (u32 wrap in bm25_idf), an empty posting list, a term
listed twice (first entry wins), stats.bin N != docs.bin count, a stored avgdl that is not the mean, tf >= 65536.
tests/golden/odd.json holds what the reference itself (oracle/_ref/ref_engine) returned for them."""
import json
import os
import struct

import numpy as np
import pytest

import fmt
import nsb200
from oracle import oracle as orc
from refcmp import check_against_reference, golden_result, oracle_result

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "odd.json")))
SEGS = {"seg_000001": 0, "seg_000002": 1}


def build(workdir, legacy, alias=False):
    idx = os.path.join(workdir, f"odd_idx_{int(legacy)}_{int(alias)}")
    if not os.path.isdir(idx):
        fmt.write_odd_index(idx, legacy, alias)
    return idx


def f32_bits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


@pytest.mark.parametrize("layout", ["barrels", "legacy"])
def test_oracle_matches_reference_on_odd_lexicon(workdir, layout):
    idx = build(workdir, layout == "legacy")
    oi = orc.OracleIndex(idx)
    for k in (10, 2):
        for row in GOLD[f"{layout}_{k}"]:
            check_against_reference(oracle_result(oi, row["query"], k), golden_result(row), oi, SEGS)


def test_oracle_matches_reference_when_entries_share_a_posting_list(workdir):
    """Two lexicon entries with the same (offset, count) and different dfs: the same postings under two idfs."""
    oi = orc.OracleIndex(build(workdir, False, alias=True))
    for row in GOLD["alias_10"]:
        check_against_reference(oracle_result(oi, row["query"], 10), golden_result(row), oi, SEGS)
    rows = {r["query"]: r for r in GOLD["alias_10"]}
    assert rows["hh"]["found"] == 3 and rows["aa hh"]["found"] == 5          # "hh" reads the list of "aa"
    assert [h[1] for h in rows["hh"]["hits"]] == [h[1] for h in rows["aa"]["hits"] if h[0] == "seg_000001"]
    assert [h[2] for h in rows["hh"]["hits"]] != [h[2] for h in rows["aa"]["hits"] if h[0] == "seg_000001"]  # other idf


def test_reference_semantics_pinned_by_the_fixture():
    """What the fixture is for, read off the reference's own answers."""
    rows = {r["query"]: r for r in GOLD["barrels_10"]}
    # df == 0: the entry is skipped although it has two postings — only the ordinary second segment answers
    assert rows["cc"]["found"] == 2 and {h[0] for h in rows["cc"]["hits"]} == {"seg_000002"}
    # an empty list: the term is usable ("found" present) and matches nothing in the odd segment
    assert rows["ee"]["found"] == 1
    # the term listed twice: the first entry (3 postings) is the one that is scored, not the later (doc 4, tf 9)
    assert rows["aa"]["found"] == 5 and ("seg_000001", 4) not in {(h[0], h[1]) for h in rows["aa"]["hits"]}
    # df != count: two postings are streamed for "bb"
    assert rows["bb"]["found"] == 3


@pytest.mark.parametrize("layout", ["barrels", "legacy"])
def test_host_front_end_resolves_odd_rows_like_the_reference(workdir, layout):
    """Host-only engine (no GPU): which rows a query names and with which idf — (:454-461)."""
    from test_known_answer import idf

    idx = build(workdir, layout == "legacy")
    e = nsb200.Engine(idx, device=None)
    assert e.reload(), e.last_error
    st = e.segment_stats(0)
    assert st["N"] == fmt.ODD_STATS_N and np.float32(st["avgdl"]) == np.float32(fmt.ODD_AVGDL)
    assert e.term_stats(0, "aa") == (3, 3)      # first entry wins
    assert e.term_stats(0, "bb") == (5, 2)
    assert e.term_stats(0, "dd") == (12, 2)
    q_off, terms, has = e.resolve_batch(["aa", "bb", "cc", "dd", "ee", "cc ee", "nosuch"])
    per_q = [terms[q_off[i]:q_off[i + 1]] for i in range(7)]
    assert list(has) == [True] * 7
    # "cc": df == 0 in the odd segment -> only segment 1 is named
    assert per_q[2]["seg"].tolist() == [1]
    # "dd": idf of the odd segment is bm25_idf(10, 12) with the u32 wrap of N - df
    dd0 = per_q[3][per_q[3]["seg"] == 0]
    assert len(dd0) == 1 and int(dd0["idf"].view(np.uint32)[0]) == f32_bits(idf(fmt.ODD_STATS_N, 12))
    assert idf(fmt.ODD_STATS_N, 12) > 15.0      # the wrap really happened: (2^32 - 2 + 0.5) / 12.5
    # "bb": idf from df (5), not from count (2)
    bb0 = per_q[1][per_q[1]["seg"] == 0]
    assert int(bb0["idf"].view(np.uint32)[0]) == f32_bits(idf(fmt.ODD_STATS_N, 5))
    assert per_q[6].size == 0
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("layout", ["barrels", "legacy"])
def test_gpu_matches_oracle_and_reference_text_on_odd_lexicon(workdir, layout):
    """The same index through the CUDA path: bit-exact against the oracle for every query and k, and the JSON text
    equal to the reference's own dump for the tie-free lists."""
    from conftest import assert_same_as_oracle

    idx = build(workdir, layout == "legacy")
    eng = nsb200.Engine(idx, device=0)
    assert eng.reload(), eng.last_error
    oi = orc.OracleIndex(idx)
    for k in (10, 2, 1, 100):
        assert_same_as_oracle(eng.search_batch(fmt.ODD_QUERIES, k), oi, fmt.ODD_QUERIES, k)
    for row in GOLD[f"{layout}_10"]:
        bits = [h[2] for h in row["hits"]]
        if len(set(bits)) != len(bits):
            continue  # order inside a tie group is a hash-map artefact of the reference
        assert eng.search_json_text(row["query"], 10) == row["text"], row["query"]
    eng.close()


@pytest.mark.gpu
def test_gpu_scores_entries_that_share_a_posting_list(workdir):
    """Overlapping rows cannot carry one resident score per posting: the segment stays on the raw-posting path
    (per-batch pre-pass, each row under its own idf) next to an ordinary segment with resident scores — and the
    results are the oracle's, bit for bit, and the reference's text."""
    from conftest import assert_same_as_oracle

    idx = build(workdir, False, alias=True)
    eng = nsb200.Engine(idx, device=0)
    assert eng.reload(), eng.last_error
    oi = orc.OracleIndex(idx)
    for k in (10, 3, 100):
        assert_same_as_oracle(eng.search_batch(fmt.ALIAS_QUERIES, k), oi, fmt.ALIAS_QUERIES, k)
    for row in GOLD["alias_10"]:
        bits = [h[2] for h in row["hits"]]
        if len(set(bits)) != len(bits):
            continue
        assert eng.search_json_text(row["query"], 10) == row["text"], row["query"]
    eng.close()
    multi = nsb200.Engine(idx, devices=[0, 0])      # the same through a two-slot engine (segment 0 / segment 1 apart)
    assert multi.reload(), multi.last_error
    assert_same_as_oracle(multi.search_batch(fmt.ALIAS_QUERIES, 10), oi, fmt.ALIAS_QUERIES, 10)
    multi.close()
