"""Randomised lexicon files (seeded): entries with arbitrary df, repeated terms, entries that share a posting list,
empty lists, stats N unrelated to the number of docs — whatever the reference's loader accepts.
  * the product's host loader / front end against the oracle (always);
  * the oracle against the LIVE reference (`ref`: where oracle/_ref/ref_engine exists) — a randomised extension of
    tests/golden/odd.json, which pins the same semantics with committed vectors."""
import os
import random

import numpy as np
import pytest

import fmt
import nsb200
from oracle import oracle as orc
from refcmp import check_against_reference, oracle_result
from test_known_answer import idf
from test_odd_lexicon import f32_bits

VOCAB = [f"w{i}" for i in range(24)]


def random_segment(rng, ndocs):
    """(stats_n, avgdl, docs, entries) for fmt.write_segment_raw."""
    docs = [(f"d{rng.randrange(10**6)}", rng.randint(1, 40)) for _ in range(ndocs)]
    entries, nplaced = [], 0
    terms = rng.sample(VOCAB, rng.randint(3, len(VOCAB)))
    for t in terms + rng.sample(terms, rng.randint(0, 3)):          # a few terms listed twice (first entry wins)
        barrel = rng.choice([0, 0, 1, 2, 17, 63])
        if nplaced and rng.random() < 0.15:                         # share an earlier entry's posting list
            entries.append((barrel, t, len(entries), rng.randint(0, 2 * ndocs), rng.randrange(nplaced)))
        else:
            n = rng.choice([0, 1, 1, 2, 3, ndocs // 2, ndocs])
            plist = [(d, rng.choice([1, 1, 2, 5, 70000])) for d in sorted(rng.sample(range(ndocs), min(n, ndocs)))]
            df = rng.choice([len(plist), len(plist), 0, rng.randint(0, 2 * ndocs)])
            entries.append((barrel, t, len(entries), df, plist))
        nplaced += 1
    stats_n = rng.choice([ndocs, ndocs, ndocs + 7, max(1, ndocs - 3)])
    avgdl = rng.choice([float(np.float32(sum(dl for _, dl in docs)) / np.float32(ndocs)), 12.5, 1.0])
    return stats_n, avgdl, docs, entries


def build(workdir, seed):
    rng = random.Random(seed)
    idx = os.path.join(workdir, f"fuzz_lex_{seed}")
    nseg = rng.randint(1, 3)
    names = [nsb200.seg_name(i + 1) for i in range(nseg)]
    if not os.path.isdir(idx):
        for name in names:
            stats_n, avgdl, docs, entries = random_segment(rng, rng.randint(4, 30))
            fmt.write_segment_raw(os.path.join(idx, "segments", name), stats_n, avgdl, docs, entries, legacy=rng.random() < 0.3)
        fmt.write_manifest(idx, names)
    queries = [" ".join(rng.choice(VOCAB) for _ in range(rng.randint(1, 4))) for _ in range(12)]
    return idx, nseg, list(dict.fromkeys(queries))                   # no query twice (the reference's LRU cache)


@pytest.mark.parametrize("seed", range(40))
def test_host_front_end_equals_oracle_on_random_lexicons(workdir, seed):
    idx, nseg, queries = build(workdir, seed)
    e = nsb200.Engine(idx, device=None)
    assert e.reload(), e.last_error
    oi = orc.OracleIndex(idx)
    assert e.num_segments == oi.num_segments == nseg
    for i in range(nseg):
        a, b = e.segment_stats(i), oi.segment_stats(i)
        # (P, the sum of the rows' counts, is not compared: with a term listed twice the two loaders count the shadowed row differently)
        assert (a["N"], a["T"]) == (b["N"], b["T"]) and np.float32(a["avgdl"]) == np.float32(b["avgdl"])
        for t in VOCAB:
            assert e.term_stats(i, t) == oi.term_stats(i, t), (seed, i, t)
    q_off, terms, has = e.resolve_batch(queries)
    for qi, q in enumerate(queries):
        want = []                                                      # (segment, idf bits) in (segment, query order)
        for i in range(nseg):
            N = oi.segment_stats(i)["N"]
            for t in q.split():
                df, cnt = oi.term_stats(i, t)
                if df == 0:
                    continue                                           # absent, or df == 0 (src/api_engine.cpp:455-458)
                want.append((i, f32_bits(idf(N, df))))
        got = [(int(x["seg"]), int(np.float32(x["idf"]).view(np.uint32))) for x in terms[q_off[qi]:q_off[qi + 1]]]
        assert got == want, (seed, q)
    e.close()


@pytest.mark.ref
@pytest.mark.parametrize("seed", range(0, 40, 4))
def test_oracle_equals_live_reference_on_random_lexicons(workdir, seed):
    idx, nseg, queries = build(workdir, seed)
    oi = orc.OracleIndex(idx)
    seg_index = {nsb200.seg_name(i + 1): i for i in range(nseg)}
    for k in (10, 2):
        _, res = orc.ref_search(idx, queries, k)
        for q, r in zip(queries, res):
            ref = {"query": q, "found": r.get("found"), "k": r["k"],
                   "hits": [(h["segment"], h["docId"], h["score_bits"]) for h in r["results"]]}
            check_against_reference(oracle_result(oi, q, k), ref, oi, seg_index)
