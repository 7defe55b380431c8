"""Reconciling results with the AS-IS reference (tests/golden/*.json).

The reference's order among equal scores is an artefact of std::unordered_map iteration and heap
order (src/api_engine.cpp:445,485-504), so parity with it is defined as (SURVEY.md §7 "Ties"):
  * identical `found` (and identical presence/absence of the key),
  * identical descending f32 score sequence, bit for bit,
  * identical (segment, docId) SETS for every score group that lies wholly inside the list,
  * for the last group of a full list (ties cut by k): every document the reference returned
    really has that score (checked with the oracle's per-document scorer).
Our own output additionally follows the stated total order, which tests check against the oracle.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np


def check_against_reference(ours, ref, oracle_index=None, seg_index=None):
    """ours/ref: dicts with keys found (None if absent), k, hits=[(segment_name, docId, score_bits)]."""
    assert ours["found"] == ref["found"], (ref["query"], ours["found"], ref["found"])
    assert ours["k"] == ref["k"]
    o_bits = [h[2] for h in ours["hits"]]
    r_bits = [h[2] for h in ref["hits"]]
    assert o_bits == r_bits, (ref["query"], o_bits[:5], r_bits[:5])
    scores = [np.array([b], np.uint32).view(np.float32)[0] for b in r_bits]
    assert all(scores[i] >= scores[i + 1] for i in range(len(scores) - 1))
    og, rg = defaultdict(set), defaultdict(set)
    for h in ours["hits"]:
        og[h[2]].add((h[0], h[1]))
    for h in ref["hits"]:
        rg[h[2]].add((h[0], h[1]))
    full = len(r_bits) == ref["k"]
    last = r_bits[-1] if r_bits else None
    for bits, docs in rg.items():
        if full and bits == last:
            if oracle_index is not None:
                for seg_name, doc in docs:
                    s, matched = oracle_index.score_doc(ref["query"], seg_index[seg_name], doc)
                    assert matched and int(np.float32(s).view(np.uint32)) == bits, (ref["query"], seg_name, doc)
        else:
            assert og[bits] == docs, (ref["query"], bits, og[bits], docs)


def oracle_result(oracle_index, query, k):
    r = oracle_index.search(query, k)
    return {"query": query, "found": r["found"], "k": r["k"],
            "hits": [(h["segment"], h["docId"], h["score_bits"]) for h in r["results"]]}


def golden_result(row):
    return {"query": row["query"], "found": row["found"], "k": row["k"],
            "hits": [(h[0], h[1], h[2]) for h in row["hits"]]}
