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


# ---- corpus with embeddings and metadata.csv: semantic expansion (weights != 1, > 5 terms) and result
# ---- decoration, both as the reference itself produces them (tests/golden/semantic.json) ----

_SEM_CLUSTERS = [
    ["virus", "viral", "viruses", "corona", "covid", "sars"],
    ["vaccine", "vaccines", "immunity", "antibody", "antibodies"],
    ["mask", "masks", "respirator", "ppe"],
    ["lung", "lungs", "pneumonia", "fever", "cough"],
    ["bat", "bats", "pangolin", "spillover"],
    ["spike", "protein", "rna", "genome", "gene"],
]
SEM_VOCAB = [w for c in _SEM_CLUSTERS for w in c]
SEM_DIM = 12


def _lcg(seed: int):
    x = seed & 0xFFFFFFFF
    while True:
        x = (x * 1664525 + 1013904223) & 0xFFFFFFFF
        yield x


def semantic_docs(seg: int) -> List[Doc]:
    """Segment `seg` (0 or 1): 24 docs over SEM_VOCAB, every doc with its own length and tf pattern
    (no two docs share a (tf, dl) pair for a term, so scores do not tie)."""
    rng = _lcg(1000 + seg)
    docs = []
    for d in range(24):
        nterms = 2 + next(rng) % 5
        picked, tfs = [], []
        for _ in range(nterms):
            w = SEM_VOCAB[next(rng) % len(SEM_VOCAB)]
            if w in picked:
                continue
            picked.append(w)
            tfs.append((w, 1 + next(rng) % 4))
        dl = 20 + 7 * d + seg * 3 + sum(tf for _, tf in tfs)
        docs.append((f"uid{seg}_{d:02d}", dl, tfs))
    return docs


def semantic_embeddings_text() -> str:
    """embeddings.vec: header, clustered vectors (cosine ~0.9 inside a cluster, ~0 across), plus lines the loader
    must skip: a word outside the lexicon, a short vector, a vector of another dimension, a repeated word."""
    rng = _lcg(77)
    lines = [f"{len(SEM_VOCAB) + 4} {SEM_DIM}"]
    for ci, cluster in enumerate(_SEM_CLUSTERS):
        for wi, w in enumerate(cluster):
            v = []
            for j in range(SEM_DIM):
                base = 1.0 if j == 2 * ci or j == 2 * ci + 1 else 0.0
                noise = ((next(rng) % 2001) - 1000) / 1000.0 * (0.18 + 0.05 * wi)
                v.append(base + noise)
            lines.append(w + " " + " ".join(f"{x:.4f}" for x in v))
    lines.append("notindexed " + " ".join("0.5" for _ in range(SEM_DIM)))      # filtered: not in any lexicon
    lines.append("rna 0.1 0.2 0.3")                                            # < 10 values: skipped
    lines.append("gene " + " ".join("0.25" for _ in range(SEM_DIM + 3)))      # other dimension: skipped
    lines.append("virus " + " ".join("0.3" for _ in range(SEM_DIM)))          # repeated word: first row stays the word's row
    lines.append("")
    return "\n".join(lines) + "\n"


SEM_QUERIES = [
    "virus", "covid vaccine", "masks", "bat virus spike", "the lung", "unknownword", "virus virus",
    "fever cough rna", "antibody", "Spike PROTEIN, of the genome!", "pangolin spillover bats", "ppe respirator mask gene",
]


def metadata_csv_text() -> str:
    """metadata.csv for the semantic corpus: quoted commas, several authors, a romanised name in parentheses,
    empty fields, a repeated uid (first row wins), a short row, a row ending in CR LF, url lists with ';'."""
    hdr = "cord_uid,sha,source_x,title,doi,publish_time,authors,journal,url"
    rows = [
        'uid0_00,s0,PMC,"Coronaviruses, bats and spillover",10.1/a,2020-01-15,"Zhou, Peng; Yang, Xing-Lou; Wang, Xian-Guang",Nature,https://a.example/1; https://b.example/1',
        'uid0_01,s1,PMC,Plain title,10.1/b,2019,"Smith John",J Virol,https://a.example/2',
        'uid0_02,s2,WHO,"Quoted ""title"" here",,2020-03,"(Li Wei) 李伟; Chen, Q",,',
        'uid0_03,s3,PMC,,10.1/d,,"  ,  ",Lancet,https://a.example/4',
        'uid0_04,s4,PMC,Title four,10.1/e,2021-05-06,"Garcia-Lopez, Maria,",BMJ,;https://second.example',
        'uid0_05,s5',
        'uid0_06,s6,PMC,Tab\tin title,10.1/f,2020,"van der Waals, J; Other, A",Cell,https://a.example/6\r',
        'uid0_00,dup,PMC,SHOULD NOT APPEAR,x,1999,"Nobody, N",None,https://dup.example',
        'uid1_00,t0,PMC,Second segment doc,10.2/a,2020-12-31,"Müller, Jürgen",Science,https://c.example/0',
        'uid1_03,t3,PMC,"Only, title",,,,,',
        ',nouid,PMC,row without uid,,,,,',
        'uid1_07,t7,PMC,Single name,10.2/h,2018-07,"Aristotle",Mind,https://c.example/7;',
    ]
    return hdr + "\n" + "\n".join(rows) + "\n"


# ---- "odd" lexicon: entries no writer of the reference produces but its loader and scoring loop accept
# ---- (tests/golden/odd.json holds what the reference itself returns for them) ----

def write_segment_raw(segdir: str, stats_n: int, avgdl: float, docs: Sequence[Tuple[str, int]],
                      entries: Sequence[Tuple[int, str, int, int, Sequence[Tuple[int, int]]]], legacy: bool = False) -> None:
    """A segment whose stats.bin, docs.bin and lexicon entries are given verbatim.
    docs: (cord_uid, doc_len); entries: (barrel, term, termId, df, postings) in file order — LexEntry.count is
    len(postings) while df is whatever the caller says (src/api_engine.cpp:458-461 read df, :473 reads count).
    `postings` may be the index of an earlier entry instead: the new entry then shares that entry's (offset, count)."""
    import numpy as np

    os.makedirs(segdir, exist_ok=True)
    with open(os.path.join(segdir, "stats.bin"), "wb") as f:
        f.write(struct.pack("<I", stats_n) + np.float32(avgdl).tobytes())
    with open(os.path.join(segdir, "docs.bin"), "wb") as f:
        f.write(struct.pack("<I", len(docs)))
        for uid, dl in docs:
            f.write(_s(uid) + _s("") + _s("") + struct.pack("<I", dl))
    nb = 1 if legacy else BARRELS
    lex: List[List[bytes]] = [[] for _ in range(nb)]
    inv: List[List[bytes]] = [[] for _ in range(nb)]
    offs = [0] * nb
    placed = []  # per entry: (file, byte offset, count)
    for barrel, term, tid, df, plist in entries:
        if isinstance(plist, int):  # an ALIAS: the entry points at the posting list of the earlier entry `plist`
            b, off, cnt = placed[plist]
            lex[b].append(_s(term) + struct.pack("<IIQI", tid, df, off, cnt))
            placed.append((b, off, cnt))
            continue
        b = 0 if legacy else barrel
        lex[b].append(_s(term) + struct.pack("<IIQI", tid, df, offs[b], len(plist)))
        inv[b].append(b"".join(struct.pack("<II", d, tf) for d, tf in plist))
        placed.append((b, offs[b], len(plist)))
        offs[b] += 8 * len(plist)
    if legacy:
        with open(os.path.join(segdir, "lexicon.bin"), "wb") as f:
            f.write(struct.pack("<I", len(lex[0])) + b"".join(lex[0]))
        with open(os.path.join(segdir, "inverted.bin"), "wb") as f:
            f.write(b"".join(inv[0]))
        return
    with open(os.path.join(segdir, "barrels.bin"), "wb") as f:
        f.write(struct.pack("<II", BARRELS, 1))
    for b in range(BARRELS):
        with open(os.path.join(segdir, f"lexicon_b{b:03d}.bin"), "wb") as f:
            f.write(struct.pack("<I", len(lex[b])) + b"".join(lex[b]))
        with open(os.path.join(segdir, f"inverted_b{b:03d}.bin"), "wb") as f:
            f.write(b"".join(inv[b]))


ODD_DOCS = [("o0", 4), ("o1", 9), ("o2", 2), ("o3", 7), ("o4", 5), ("o5", 3)]
ODD_STATS_N = 10       # stats.bin N != number of docs in docs.bin: idf uses stats N (src/api_engine.cpp:461)
ODD_AVGDL = 5.25       # stored verbatim, not the mean of the doc lengths (:478 reads seg.avgdl)
ODD_ENTRIES = [
    (0, "aa", 0, 3, [(0, 1), (2, 2), (5, 1)]),
    (0, "bb", 1, 5, [(1, 3), (4, 1)]),          # df != count: idf from df, two postings streamed
    (1, "cc", 2, 0, [(0, 1), (3, 1)]),          # df == 0: the term is skipped although it has postings (:458)
    (1, "dd", 3, 12, [(2, 1), (3, 4)]),         # df > N: (N - df) wraps in u32 before the float conversion (:46)
    (2, "ee", 4, 2, []),                        # usable term without postings
    (3, "aa", 5, 1, [(4, 9)]),                  # the term again, later file: lex.emplace keeps the FIRST (api_segment.cpp:99)
    (63, "ff", 6, 1, [(5, 2)]),
    (63, "gg", 7, 6, [(0, 2), (1, 1), (2, 1), (3, 1), (4, 1), (5, 70000)]),   # tf >= 65536: unpacked device format
]
# (no query twice: the reference answers a repeated (query, k) from its LRU cache and marks it "from_cache")
ODD_QUERIES = ["aa", "bb", "cc", "dd", "ee", "ff", "gg", "aa bb cc dd ee ff gg", "cc ee", "dd dd", "ff aa", "bb aa dd",
               "gg aa", "cc cc dd"]


def odd_second_segment_docs() -> List[Doc]:
    """An ordinary second segment holding the same terms, so that the odd rows also meet the cross-segment top-k."""
    return [
        ("p0", 6, [("aa", 1), ("cc", 2)]),
        ("p1", 3, [("bb", 1), ("dd", 1)]),
        ("p2", 8, [("cc", 1), ("ee", 3), ("ff", 1)]),
        ("p3", 5, [("aa", 2), ("dd", 2), ("gg", 1)]),
    ]


# Two more entries that SHARE the posting lists of "aa" and "gg" under other names and dfs: the reference only follows
# (offset, count), so the same postings are scored with two different idfs.  (A separate index: on the device such
# a segment cannot carry one resident score per posting and stays on the raw-posting path.)
ALIAS_ENTRIES = ODD_ENTRIES + [
    (0, "hh", 8, 2, 0),      # the list of "aa" (first entry), df 2 instead of 3
    (63, "kk", 9, 1, 7),     # the list of "gg", df 1 instead of 6
]
ALIAS_QUERIES = ["hh", "aa hh", "hh aa", "kk gg", "kk", "aa bb hh kk", "hh hh"] + ODD_QUERIES[:8]


def write_odd_index(index_dir: str, legacy: bool = False, alias: bool = False) -> None:
    write_segment_raw(os.path.join(index_dir, "segments", "seg_000001"), ODD_STATS_N, ODD_AVGDL, ODD_DOCS,
                      ALIAS_ENTRIES if alias else ODD_ENTRIES, legacy)
    write_segment(os.path.join(index_dir, "segments", "seg_000002"), odd_second_segment_docs())
    write_manifest(index_dir, ["seg_000001", "seg_000002"])
