"""Pure-Python writer of NextSearch's on-disk segment format for HAND-MADE corpora (tests only).

A third, independent statement of the format (Appendix A of SURVEY.md) next to the product's C++
writer and the oracle's C reader; tests/golden/*.json hold the sha256 of the files the reference's
own SegmentWriter (include/segment_writer.hpp:48-168) produced for the same documents.
"""
from __future__ import annotations

import os
import struct
from typing import List, Sequence, Tuple

Doc = Tuple[str, int, Sequence[Tuple[str, int]]]  # (cord_uid, doc_len, [(term, tf), ...])

BARRELS = 64


def _s(b: str) -> bytes:
    raw = b.encode("utf-8")
    return struct.pack("<I", len(raw)) + raw


def write_manifest(index_dir: str, names: Sequence[str]) -> None:
    os.makedirs(index_dir, exist_ok=True)
    with open(os.path.join(index_dir, "manifest.bin"), "wb") as f:
        f.write(struct.pack("<I", len(names)) + b"".join(_s(n) for n in names))


def write_segment(segdir: str, docs: Sequence[Doc], legacy: bool = False, write_forward: bool = True) -> None:
    """legacy=True writes lexicon.bin + inverted.bin instead of the 64 barrels
    (the layout src/api_segment.cpp:45-67 reads when barrels.bin is absent)."""
    os.makedirs(segdir, exist_ok=True)
    term_id, terms, inverted = {}, [], []
    total_len = 0
    forward = []
    for doc_id, (_uid, dl, tfs) in enumerate(docs):
        total_len += dl
        row = []
        for term, tf in tfs:  # intern in first-seen order (segment_writer.hpp:38-46)
            if term not in term_id:
                term_id[term] = len(terms)
                terms.append(term)
                inverted.append([])
            tid = term_id[term]
            row.append((tid, tf))
            inverted[tid].append((doc_id, tf))
        forward.append(sorted(row))
    n = len(docs)
    # avgdl = (float)total_len / (float)ndocs  (segment_writer.hpp:68)
    import numpy as np

    avgdl = np.float32(0.0) if n == 0 else np.float32(total_len) / np.float32(n)
    with open(os.path.join(segdir, "stats.bin"), "wb") as f:
        f.write(struct.pack("<I", n) + np.float32(avgdl).tobytes())
    with open(os.path.join(segdir, "docs.bin"), "wb") as f:
        f.write(struct.pack("<I", n))
        for uid, dl, _ in docs:
            f.write(_s(uid) + _s("") + _s("") + struct.pack("<I", dl))
    if write_forward:
        with open(os.path.join(segdir, "forward.bin"), "wb") as f:
            f.write(struct.pack("<I", n))
            for row in forward:
                f.write(struct.pack("<I", len(row)) + b"".join(struct.pack("<II", t, tf) for t, tf in row))
        with open(os.path.join(segdir, "terms.bin"), "wb") as f:
            f.write(struct.pack("<I", len(terms)) + b"".join(_s(t) for t in terms))

    def entry(tid: int, offset: int) -> bytes:
        df = len(inverted[tid])
        return _s(terms[tid]) + struct.pack("<IIQI", tid, df, offset, df)

    def plist(tid: int) -> bytes:
        return b"".join(struct.pack("<II", d, tf) for d, tf in sorted(inverted[tid]))

    T = len(terms)
    if legacy:
        lex, inv, off = [], [], 0
        for tid in range(T):
            lex.append(entry(tid, off))
            inv.append(plist(tid))
            off += 8 * len(inverted[tid])
        with open(os.path.join(segdir, "lexicon.bin"), "wb") as f:
            f.write(struct.pack("<I", T) + b"".join(lex))
        with open(os.path.join(segdir, "inverted.bin"), "wb") as f:
            f.write(b"".join(inv))
        return
    tpb = max(1, (T + BARRELS - 1) // BARRELS)
    with open(os.path.join(segdir, "barrels.bin"), "wb") as f:
        f.write(struct.pack("<II", BARRELS, tpb))
    lex: List[List[bytes]] = [[] for _ in range(BARRELS)]
    inv: List[List[bytes]] = [[] for _ in range(BARRELS)]
    offs = [0] * BARRELS
    for tid in range(T):
        b = min(tid // tpb, BARRELS - 1)
        lex[b].append(entry(tid, offs[b]))
        inv[b].append(plist(tid))
        offs[b] += 8 * len(inverted[tid])
    for b in range(BARRELS):
        with open(os.path.join(segdir, f"lexicon_b{b:03d}.bin"), "wb") as f:
            f.write(struct.pack("<I", len(lex[b])) + b"".join(lex[b]))
        with open(os.path.join(segdir, f"inverted_b{b:03d}.bin"), "wb") as f:
            f.write(b"".join(inv[b]))


def write_dump(path: str, docs: Sequence[Doc]) -> None:
    """The flat corpus dump oracle/ref_driver `write` feeds to the reference SegmentWriter."""
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(docs)))
        for uid, dl, tfs in docs:
            f.write(_s(uid) + struct.pack("<II", dl, len(tfs)))
            for term, tf in tfs:
                f.write(_s(term) + struct.pack("<I", tf))


# ---- hand-made corpora shared by the golden generator and the tests ----

def handmade_docs() -> List[Doc]:
    """12 docs, 9 terms; several docs share (tf, dl) for a term => exact score ties."""
    return [
        ("h0", 5, [("alpha", 2), ("beta", 1), ("gamma", 2)]),
        ("h1", 3, [("alpha", 1), ("delta", 2)]),
        ("h2", 8, [("beta", 4), ("gamma", 1), ("delta", 1), ("eps", 2)]),
        ("h3", 3, [("alpha", 1), ("delta", 2)]),           # same (tf, dl) as h1 for alpha and delta
        ("h4", 1, [("zeta9", 1)]),
        ("h5", 6, [("alpha", 3), ("beta", 3)]),
        ("h6", 2, [("t2", 1), ("ab", 1)]),
        ("h7", 10, [("alpha", 1), ("beta", 1), ("gamma", 1), ("delta", 1), ("eps", 1), ("t2", 5)]),
        ("h8", 4, [("gamma", 4)]),
        ("h9", 3, [("alpha", 1), ("delta", 2)]),           # third member of the tie group
        ("h10", 7, [("eps", 7)]),
        ("h11", 5, [("ab", 2), ("beta", 3)]),
    ]


HANDMADE_QUERIES = [
    "alpha", "alpha delta", "delta alpha", "alpha alpha", "beta gamma eps", "gamma", "t2 ab", "zeta9",
    "Alpha, the BETA!", "the of and", "", "x y", "nosuchterm", "nosuchterm alpha", "eps eps eps", "alpha beta gamma delta eps t2 ab zeta9",
]


def tie_docs(n: int, base: int) -> List[Doc]:
    """n identical docs: term 'aa' tf=1, doc_len 10 (SURVEY.md Appendix B probe)."""
    return [(f"tie{base + i}", 10, [("aa", 1)]) for i in range(n)]
