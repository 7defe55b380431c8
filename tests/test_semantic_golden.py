"""Weighted scoring, semantic expansion and result decoration against outputs of the reference itself
(tests/golden/semantic.json: the reference run on the hand-made corpus of tests/fmt.py with an
embeddings.vec and a metadata.csv — generating script tests/golden/make_golden.py).

  * the oracle's weighted entry (qweight != 1, > 5 terms per query) reproduces the reference's results
    bit for bit when it is fed the reference's own qterms_w  -> pins the weighted path of the oracle;
  * the product's expansion (host/semantic.hpp) reproduces that qterms_w: terms, f32 weights AND order;
  * (GPU, tests/test_gpu_semantic.py) the CUDA path matches the oracle on those lists, and the product's JSON
    text equals the reference's j.dump() byte for byte, decoration included."""
import json
import os

import numpy as np
import pytest

import fmt
import nsb200
from oracle import oracle as orc

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "semantic.json")))


def build_sem_index(path, embeddings=True, metadata=True):
    for s in range(2):
        fmt.write_segment(os.path.join(path, "segments", nsb200.seg_name(s + 1)), fmt.semantic_docs(s))
    fmt.write_manifest(path, [nsb200.seg_name(1), nsb200.seg_name(2)])
    if metadata:
        with open(os.path.join(path, "metadata.csv"), "w", newline="") as f:
            f.write(fmt.metadata_csv_text())
    if embeddings:
        with open(os.path.join(path, "embeddings.vec"), "w", newline="") as f:
            f.write(fmt.semantic_embeddings_text())
    return path


@pytest.fixture(scope="module")
def sem_index(workdir):
    return build_sem_index(os.path.join(workdir, "sem_idx"))


def qterms_of(row):
    return [(t, float(np.array([b], np.uint32).view(np.float32)[0])) for t, b in row["qterms"]]


def test_fixture_has_real_weights():
    rows = GOLD["expanded"]["10"]
    assert any(len(r["qterms"]) > 5 for r in rows)
    assert any(b != 0x3F800000 for r in rows for _, b in r["qterms"])


@pytest.mark.parametrize("k", ["10", "100"])
def test_oracle_weighted_matches_reference(sem_index, k):
    oi = orc.OracleIndex(sem_index)
    for row in GOLD["expanded"][k]:
        got = oi.search_weighted(qterms_of(row), int(k))
        assert got["found"] == row["found"], row["query"]
        assert [(h["segment"], h["docId"], h["score_bits"], h["cord_uid"]) for h in got["results"]] == \
               [tuple(h) for h in row["hits"]], row["query"]


@pytest.mark.parametrize("k", ["10", "3"])
def test_oracle_plain_matches_reference_on_the_same_corpus(workdir, k):
    idx = build_sem_index(os.path.join(workdir, "sem_idx_plain"), embeddings=False)
    oi = orc.OracleIndex(idx)
    for row in GOLD["plain"][k]:
        got = oi.search(row["query"], int(k))
        assert got["found"] == row["found"], row["query"]
        assert [(h["segment"], h["docId"], h["score_bits"]) for h in got["results"]] == [tuple(h[:3]) for h in row["hits"]]


def test_product_expansion_equals_reference_qterms(sem_index):
    """terms, weight BITS and order of SemanticIndex::expand (src/semantic_embedding.cpp:148-229) — the order is
    the order of float additions in the scoring loop, so it is part of the result."""
    eng = nsb200.Engine(sem_index, device=None)  # host-only: expansion needs no GPU
    assert eng.reload(), eng.last_error
    for row in GOLD["expanded"]["10"]:
        got = eng.expand(row["query"])
        assert got is not None
        got_bits = [(t, int(np.float32(w).view(np.uint32))) for t, w in got]
        assert got_bits == [tuple(x) for x in row["qterms"]], row["query"]
    eng.close()


def test_no_embeddings_means_no_expansion(workdir):
    idx = build_sem_index(os.path.join(workdir, "sem_idx_noemb"), embeddings=False)
    eng = nsb200.Engine(idx, device=None)
    assert eng.reload()
    assert eng.expand("virus") is None
    eng.close()
