"""Randomised embeddings files and corpora (seeded) against the LIVE reference (`ref`: where oracle/_ref/ref_engine
exists): the product's expansion (host/semantic.hpp) must give the reference's qterms_w — terms, f32 weight bits and
order — and the oracle's weighted entry the reference's results for those lists.  A randomised extension of
tests/golden/semantic.json (committed vectors, checked everywhere)."""
import os
import random

import numpy as np
import pytest

import fmt
import nsb200
from oracle import oracle as orc

WORDS = [f"{p}{i}" for p in ("vir", "vac", "lun", "bat", "rna", "ppe") for i in range(9)]


def build(workdir, seed):
    rng = random.Random(1000 + seed)
    idx = os.path.join(workdir, f"fuzz_sem_{seed}")
    dim = rng.choice([10, 12, 16, 25])
    if not os.path.isdir(idx):
        for s in range(2):
            docs = []
            for d in range(rng.randint(15, 40)):
                picked = rng.sample(WORDS, rng.randint(1, 6))
                tfs = [(w, rng.randint(1, 5)) for w in picked]
                docs.append((f"u{s}_{d}", 10 + 3 * d + s + sum(tf for _, tf in tfs), tfs))
            fmt.write_segment(os.path.join(idx, "segments", nsb200.seg_name(s + 1)), docs)
        fmt.write_manifest(idx, [nsb200.seg_name(1), nsb200.seg_name(2)])
        nclusters = 6
        centres = [[rng.gauss(0, 1) for _ in range(dim)] for _ in range(nclusters)]
        lines = [f"{len(WORDS) + 3} {dim}"]
        order = WORDS[:]
        rng.shuffle(order)
        for w in order:
            if rng.random() < 0.1:
                continue                                              # a word without a vector
            c = centres[WORDS.index(w) // 9]
            spread = rng.choice([0.15, 0.3, 0.6])
            lines.append(w + " " + " ".join(f"{x + rng.gauss(0, spread):.4f}" for x in c))
        lines.append("notindexed " + " ".join("0.5" for _ in range(dim)))     # not in any lexicon: filtered
        lines.append(order[0] + " 0.1 0.2 0.3")                               # too short: skipped
        lines.append(order[1] + " " + " ".join("0.25" for _ in range(dim + 2)))  # other dimension: skipped
        lines.append(order[2] + " " + " ".join("0.3" for _ in range(dim)))    # a word again
        with open(os.path.join(idx, "embeddings.vec"), "w", newline="") as f:
            f.write("\n".join(lines) + "\n")
    queries = [" ".join(rng.choice(WORDS + ["unknownword", "the"]) for _ in range(rng.randint(1, 3))) for _ in range(14)]
    return idx, list(dict.fromkeys(queries))


@pytest.mark.ref
@pytest.mark.parametrize("seed", range(16))
def test_expansion_and_weighted_scoring_equal_the_live_reference(workdir, seed):
    idx, queries = build(workdir, seed)
    _, res = orc.ref_search(idx, queries, 10)
    assert any("_qterms" in r for r in res), "the reference did not load the embeddings"
    eng = nsb200.Engine(idx, device=None)          # host-only: the expansion needs no GPU
    assert eng.reload(), eng.last_error
    oi = orc.OracleIndex(idx)
    expanded = 0
    for q, r in zip(queries, res):
        got = eng.expand(q)
        if "_qterms" not in r:                      # no usable token, or the expansion came back empty (:407, :424)
            assert not got, q
            continue
        want = [tuple(x) for x in r["_qterms"]]
        assert got is not None, q
        assert [(t, int(np.float32(w).view(np.uint32))) for t, w in got] == want, (seed, q)
        expanded += len(want) > len(set(q.split()))
        lst = [(t, float(np.array([b], np.uint32).view(np.float32)[0])) for t, b in want]
        o = oi.search_weighted(lst, 10)
        assert o["found"] == r.get("found"), (seed, q)
        ref_bits = [h["score_bits"] for h in r["results"]]
        assert [h["score_bits"] for h in o["results"]] == ref_bits, (seed, q)
        if len(set(ref_bits)) == len(ref_bits):     # tie-free: the documents must be the same, in the same order
            assert [(h["segment"], h["docId"]) for h in o["results"]] == [(h["segment"], h["docId"]) for h in r["results"]], (seed, q)
    assert expanded > 0                             # the fixture really produces neighbours
    eng.close()
