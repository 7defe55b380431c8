#!/usr/bin/env python
"""Generates tests/golden/*.json by running the REFERENCE ITSELF in this container.

    python tests/golden/make_golden.py        (needs /root/reference -> oracle/_ref/ref_engine)

oracle/_ref/ref_engine is the reference's unmodified src/api_engine.cpp, src/api_segment.cpp (+
autocomplete/metadata/semantic objects) and include/segment_writer.hpp compiled by oracle/Makefile.
The reference ships no tests or golden vectors of its own, so these files are the
pin: what cord19::Engine::search returns, and the bytes SegmentWriter::write_segment emits, for
corpora that the test-suite can rebuild without the reference.

Outputs
  small_index.json     sha256 of every file SegmentWriter wrote for the generated 2x1500-doc corpus
                       + Engine::search results for generated and edge-case queries at several k
  handmade.json        the same for the 12-doc hand-made corpus of tests/fmt.py (barrels and legacy)
  ties.json            2 segments x 1000 identical docs: found, scores, which segment wins
  semantic.json        2 x 24-doc corpus WITH embeddings.vec and metadata.csv (tests/fmt.py): for every query the
                       reference's expanded (term, weight bits) list (SemanticIndex::expand), its results, and the
                       exact text of its JSON (j.dump()) including title / url / publish_time / author; the same
                       queries on the same index without the embeddings file (decoration only)
  odd.json             a hand-written lexicon no writer produces but the reference's loader and scoring loop accept
                       (tests/fmt.py: df != count, df == 0 with postings, df > N, an empty list, a term listed twice,
                       stats N != docs, a stored avgdl that is not the mean, tf >= 65536), barrels and legacy layout
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import nsb200  # noqa: E402
import fmt  # noqa: E402
from conftest import EDGE_QUERIES  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def sha_dir(d):
    return {f: hashlib.sha256(open(os.path.join(d, f), "rb").read()).hexdigest() for f in sorted(os.listdir(d))}


def compact(results, text=False):
    """Engine::search JSON -> compact rows (text=True keeps the reference's own j.dump() and qterms_w)."""
    out = []
    for r in results:
        row = {
            "query": r["query"], "k": r["k"], "segments": r["segments"], "found": r.get("found"),
            "hits": [[h["segment"], h["docId"], h["score_bits"], h["cord_uid"]] for h in r["results"]],
        }
        if text:
            row["text"] = r["_text"]
            if "_qterms" in r:
                row["qterms"] = r["_qterms"]
        out.append(row)
    return out


def main():
    if not orc.have_ref():
        raise SystemExit("oracle/_ref/ref_engine missing: run `make -C oracle` where /root/reference exists")
    td = tempfile.mkdtemp(prefix="nsb200_golden_")
    try:
        # ---- generated corpus: 2 segments x 1500 docs, V=3000 (tests' small_case) ----
        spec = nsb200.CorpusSpec(vocab=3000)
        ref_idx = os.path.join(td, "ref_small")
        hashes = {}
        names = []
        for s in range(2):
            name = nsb200.seg_name(s + 1)
            mine = os.path.join(td, "mine_small", "segments", name)
            dump = os.path.join(td, f"dump{s}.bin")
            nsb200.write_segment(spec, s * 1500, 1500, mine, True, dump)
            seg = os.path.join(ref_idx, "segments", name)
            orc.ref_write_segment(dump, seg)         # the reference's SegmentWriter
            hashes[name] = sha_dir(seg)
            names.append(name)
        orc.ref_manifest(ref_idx, names)              # the reference's save_manifest
        manifest_sha = hashlib.sha256(open(os.path.join(ref_idx, "manifest.bin"), "rb").read()).hexdigest()
        queries = nsb200.make_queries(spec, 150, 1, 5) + EDGE_QUERIES
        runs = {}
        for k in (10, 1, 3, 100, 1000, 0):
            qs = queries if k == 10 else queries[:40] + EDGE_QUERIES
            _, res = orc.ref_search(ref_idx, qs, k)
            runs[str(k)] = compact(res)
        json.dump({"spec": {"vocab": 3000, "seed": spec.seed, "zipf_s": spec.zipf_s, "zipf_q": spec.zipf_q,
                            "len_lo": spec.len_lo, "len_hi": spec.len_hi},
                   "ndocs": 3000, "nseg": 2, "segment_sha256": hashes, "manifest_sha256": manifest_sha,
                   "search": runs}, open(os.path.join(HERE, "small_index.json"), "w"), indent=0)

        # ---- hand-made corpus (barrels + legacy layouts) ----
        docs = fmt.handmade_docs()
        dump = os.path.join(td, "hand.bin")
        fmt.write_dump(dump, docs)
        hand_idx = os.path.join(td, "ref_hand")
        seg = os.path.join(hand_idx, "segments", "seg_000001")
        orc.ref_write_segment(dump, seg)
        orc.ref_manifest(hand_idx, ["seg_000001"])
        hand = {"segment_sha256": sha_dir(seg), "search": {}}
        for k in (10, 2):
            _, res = orc.ref_search(hand_idx, fmt.HANDMADE_QUERIES, k)
            hand["search"][str(k)] = compact(res, text=(k == 10))
        # legacy layout: the reference has no writer for multi-doc legacy segments (only
        # src/AddDocument.cpp, one doc); write it with tests/fmt.py and let the REFERENCE read it
        leg_idx = os.path.join(td, "leg_hand")
        fmt.write_segment(os.path.join(leg_idx, "segments", "seg_000001"), docs, legacy=True)
        fmt.write_manifest(leg_idx, ["seg_000001"])
        _, res = orc.ref_search(leg_idx, fmt.HANDMADE_QUERIES, 10)
        hand["search_legacy_10"] = compact(res)
        json.dump(hand, open(os.path.join(HERE, "handmade.json"), "w"), indent=0)

        # ---- ties: 2 segments x 1000 identical docs ----
        tie_idx = os.path.join(td, "ref_ties")
        tie_hash = {}
        for s in range(2):
            name = nsb200.seg_name(s + 1)
            dump = os.path.join(td, f"tie{s}.bin")
            fmt.write_dump(dump, fmt.tie_docs(1000, s * 1000))
            seg = os.path.join(tie_idx, "segments", name)
            orc.ref_write_segment(dump, seg)
            tie_hash[name] = sha_dir(seg)
        orc.ref_manifest(tie_idx, [nsb200.seg_name(1), nsb200.seg_name(2)])
        ties = {"segment_sha256": tie_hash, "search": {}}
        for k in (10, 100):
            _, res = orc.ref_search(tie_idx, ["aa", "aa aa"], k)
            ties["search"][str(k)] = compact(res)
        json.dump(ties, open(os.path.join(HERE, "ties.json"), "w"), indent=0)
        # ---- semantic expansion + metadata decoration ----
        sem_idx = os.path.join(td, "ref_sem")
        sem_hash = {}
        for s in range(2):
            name = nsb200.seg_name(s + 1)
            dump = os.path.join(td, f"sem{s}.bin")
            fmt.write_dump(dump, fmt.semantic_docs(s))
            seg = os.path.join(sem_idx, "segments", name)
            orc.ref_write_segment(dump, seg)
            sem_hash[name] = sha_dir(seg)
        orc.ref_manifest(sem_idx, [nsb200.seg_name(1), nsb200.seg_name(2)])
        with open(os.path.join(sem_idx, "metadata.csv"), "w", newline="") as f:
            f.write(fmt.metadata_csv_text())
        sem = {"segment_sha256": sem_hash, "plain": {}, "expanded": {}}
        for k in (10, 3):
            _, res = orc.ref_search(sem_idx, fmt.SEM_QUERIES, k)      # no embeddings file yet: decoration only
            sem["plain"][str(k)] = compact(res, text=True)
        with open(os.path.join(sem_idx, "embeddings.vec"), "w", newline="") as f:
            f.write(fmt.semantic_embeddings_text())
        for k in (10, 100):
            _, res = orc.ref_search(sem_idx, fmt.SEM_QUERIES, k)
            rows = compact(res, text=True)
            assert all("qterms" in r for r in rows), "embeddings were not loaded by the reference"
            sem["expanded"][str(k)] = rows
        for mode in ("plain", "expanded"):                            # byte-level text parity needs tie-free lists
            for rows in sem[mode].values():
                for r in rows:
                    bits = [h[2] for h in r["hits"]]
                    assert len(set(bits)) == len(bits), ("score tie in the semantic fixture", r["query"])
        json.dump(sem, open(os.path.join(HERE, "semantic.json"), "w"), indent=0)
        # ---- odd lexicon entries (hand-written files, read by the reference) ----
        odd = {}
        for legacy in (False, True):
            odd_idx = os.path.join(td, f"odd_{int(legacy)}")
            fmt.write_odd_index(odd_idx, legacy)
            for k in (10, 2):
                _, res = orc.ref_search(odd_idx, fmt.ODD_QUERIES, k)
                odd[f"{'legacy' if legacy else 'barrels'}_{k}"] = compact(res, text=(k == 10))
        alias_idx = os.path.join(td, "odd_alias")
        fmt.write_odd_index(alias_idx, alias=True)      # two entries sharing one posting list
        _, res = orc.ref_search(alias_idx, fmt.ALIAS_QUERIES, 10)
        odd["alias_10"] = compact(res, text=True)
        json.dump(odd, open(os.path.join(HERE, "odd.json"), "w"), indent=0)
        print("golden fixtures written to", HERE)
    finally:
        shutil.rmtree(td, ignore_errors=True)


if __name__ == "__main__":
    main()
