"""Synthetic CORD-19-shaped corpus in the reference's on-disk format (BASELINE.json configs).

The generator and writer live in libnsb200.so (csrc/host/corpus.cpp); this module only names the
workloads.  The reference's include/segment_writer.hpp:23-169 produces the same bytes for the same
documents (byte-for-byte test under tests/) but needs ~2.5 min and ~2 GB per 1M docs.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional

from . import _lib
from ._lib import CorpusSpecC, check


@dataclass(frozen=True)
class CorpusSpec:
    seed: int = 20260101
    vocab: int = 400_000
    zipf_s: float = 1.0
    zipf_q: float = 25.0
    len_lo: int = 100
    len_hi: int = 250

    def c(self) -> CorpusSpecC:
        return CorpusSpecC(self.seed, self.vocab, self.zipf_s, self.zipf_q, self.len_lo, self.len_hi)


# BASELINE.json configs: V=50k at 10k docs, V=400k at 1M docs
SPEC_10K = CorpusSpec(vocab=50_000)
SPEC_1M = CorpusSpec(vocab=400_000)
QUERY_SEED = 7


def seg_name(i: int) -> str:
    """seg_name() of src/api_segment.cpp:38-42; the reference numbers segments from 1."""
    return f"seg_{i:06d}"


def write_segment(spec: CorpusSpec, doc_base: int, ndocs: int, segdir: str, write_forward: bool = False,
                  dump_path: Optional[str] = None, nthreads: int = 0) -> None:
    lib = _lib.load()
    cs = spec.c()
    check(lib.ns_corpus_write_segment(C.byref(cs), int(doc_base), int(ndocs), segdir.encode(), int(write_forward),
                                      dump_path.encode() if dump_path else None, int(nthreads)))


def write_manifest(index_dir: str, names: List[str]) -> None:
    lib = _lib.load()
    arr = (C.c_char_p * max(1, len(names)))(*[n.encode() for n in names])
    check(lib.ns_corpus_write_manifest(index_dir.encode(), len(names), arr))


def build_index(index_dir: str, spec: CorpusSpec, ndocs: int, nseg: int = 1, write_forward: bool = False,
                nthreads: int = 0) -> List[str]:
    """ndocs documents split by contiguous doc ranges into nseg segments seg_000001.. + manifest.bin."""
    names = []
    per = ndocs // nseg
    for s in range(nseg):
        base = s * per
        n = per if s < nseg - 1 else ndocs - base
        name = seg_name(s + 1)
        write_segment(spec, base, n, os.path.join(index_dir, "segments", name), write_forward, None, nthreads)
        names.append(name)
    write_manifest(index_dir, names)
    return names


def make_queries(spec: CorpusSpec, nq: int, min_terms: int = 1, max_terms: int = 5, seed: int = QUERY_SEED,
                 head_ranks: int = 0) -> List[str]:
    lib = _lib.load()
    cs = spec.c()
    need = C.c_size_t()
    check(lib.ns_corpus_make_queries(C.byref(cs), seed, nq, min_terms, max_terms, head_ranks, None, 0, C.byref(need)))
    buf = C.create_string_buffer(max(1, need.value))
    check(lib.ns_corpus_make_queries(C.byref(cs), seed, nq, min_terms, max_terms, head_ranks, buf, need.value,
                                     C.byref(need)))
    return [s.decode() for s in buf.raw[: need.value].split(b"\0")[:nq]]
