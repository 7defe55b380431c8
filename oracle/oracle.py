"""ctypes wrapper of oracle/liboracle.so and runner of oracle/_ref/ref_engine.

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under nextsearch-api_b200/ may import this.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess
import tempfile
from typing import List, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_ENGINE = os.path.join(HERE, "_ref", "ref_engine")
SHIM_ENGINE = os.path.join(HERE, "_ref", "shim_engine")  # INTEGRATION.md §1 compiled against the reference's header

_lib = None


def build(quiet: bool = True) -> None:
    """make -C oracle: liboracle.so always; _ref/ref_engine when /root/reference is present."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True)


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    lib.orc_last_error.restype = C.c_char_p
    lib.orc_open.restype = P
    lib.orc_open.argtypes = [C.c_char_p]
    lib.orc_close.argtypes = [P]
    lib.orc_num_segments.restype = C.c_int
    lib.orc_num_segments.argtypes = [P]
    lib.orc_segment_name.restype = C.c_char_p
    lib.orc_segment_name.argtypes = [P, C.c_int]
    lib.orc_cord_uid.restype = C.c_char_p
    lib.orc_cord_uid.argtypes = [P, C.c_uint32, C.c_uint32]
    lib.orc_segment_stats.argtypes = [P, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.POINTER(C.c_uint32),
                                      C.POINTER(C.c_uint64)]
    lib.orc_term_stats.restype = C.c_int
    lib.orc_term_stats.argtypes = [P, C.c_int, C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.orc_query_terms.restype = C.c_int
    lib.orc_query_terms.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int]
    lib.orc_search.restype = C.c_int
    lib.orc_search.argtypes = [P, C.c_char_p, C.c_int, P, P, P, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
    lib.orc_search_weighted.restype = C.c_int
    lib.orc_search_weighted.argtypes = [P, C.c_int, C.POINTER(C.c_char_p), P, C.c_int, P, P, P, C.POINTER(C.c_uint64),
                                        C.POINTER(C.c_int)]
    lib.orc_score_doc.argtypes = [P, C.c_char_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    lib.orc_search_many.restype = C.c_double
    lib.orc_search_many.argtypes = [P, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, P, P, P, P, P, P]
    lib.orc_query_postings.restype = C.c_uint64
    lib.orc_query_postings.argtypes = [P, C.c_char_p]
    _lib = lib
    return lib


def _p(a: Optional[np.ndarray]):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleIndex:
    def __init__(self, index_dir: str):
        self.lib = load()
        self.h = self.lib.orc_open(str(index_dir).encode())
        if not self.h:
            raise RuntimeError("oracle: " + (self.lib.orc_last_error() or b"").decode())

    @property
    def num_segments(self) -> int:
        return self.lib.orc_num_segments(self.h)

    def segment_name(self, i: int) -> str:
        return self.lib.orc_segment_name(self.h, i).decode()

    def cord_uid(self, seg: int, doc: int) -> str:
        return self.lib.orc_cord_uid(self.h, seg, doc).decode("utf-8", "replace")

    def segment_stats(self, i: int) -> dict:
        N, T, P, a = C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_float()
        self.lib.orc_segment_stats(self.h, i, C.byref(N), C.byref(a), C.byref(T), C.byref(P))
        return {"N": N.value, "avgdl": a.value, "T": T.value, "P": P.value}

    def term_stats(self, i: int, term: str):
        df, cnt = C.c_uint32(), C.c_uint32()
        self.lib.orc_term_stats(self.h, i, term.encode(), C.byref(df), C.byref(cnt))
        return df.value, cnt.value

    def search(self, query: str, k: int = 10) -> dict:
        """Same fields as the reference's JSON: found (None when omitted), k, results[...]."""
        K = max(1, min(int(k), 100))
        s = np.zeros(100, np.float32)
        g = np.zeros(100, np.uint32)
        d = np.zeros(100, np.uint32)
        found, has = C.c_uint64(), C.c_int()
        n = self.lib.orc_search(self.h, query.encode("utf-8"), int(k), _p(s), _p(g), _p(d), C.byref(found), C.byref(has))
        res = [{"score": float(s[i]), "score_bits": int(s[i:i + 1].view(np.uint32)[0]), "seg": int(g[i]),
                "segment": self.segment_name(int(g[i])), "docId": int(d[i]), "cord_uid": self.cord_uid(int(g[i]), int(d[i]))}
               for i in range(n)]
        return {"query": query, "k": K, "segments": self.num_segments, "found": found.value if has.value else None,
                "results": res}

    def search_weighted(self, qterms, k: int = 10) -> dict:
        """The scoring loop over an explicit [(term, weight)] list — the reference's qterms_w."""
        K = max(1, min(int(k), 100))
        n = len(qterms)
        arr = (C.c_char_p * max(1, n))(*[t.encode("utf-8") for t, _ in qterms])
        w = np.ascontiguousarray([x for _, x in qterms] if n else [0.0], dtype=np.float32)
        s = np.zeros(100, np.float32)
        g = np.zeros(100, np.uint32)
        d = np.zeros(100, np.uint32)
        found, has = C.c_uint64(), C.c_int()
        nh = self.lib.orc_search_weighted(self.h, n, arr, _p(w), int(k), _p(s), _p(g), _p(d), C.byref(found), C.byref(has))
        if nh < 0:
            raise ValueError("too many terms")
        res = [{"score": float(s[i]), "score_bits": int(s[i:i + 1].view(np.uint32)[0]), "seg": int(g[i]),
                "segment": self.segment_name(int(g[i])), "docId": int(d[i]), "cord_uid": self.cord_uid(int(g[i]), int(d[i]))}
               for i in range(nh)]
        return {"k": K, "segments": self.num_segments, "found": found.value if has.value else None, "results": res}

    def score_doc(self, query: str, seg: int, doc: int):
        sc, m = C.c_float(), C.c_int()
        self.lib.orc_score_doc(self.h, query.encode("utf-8"), seg, doc, C.byref(sc), C.byref(m))
        return (np.float32(sc.value), bool(m.value))

    def search_many(self, queries: Sequence[str], k: int = 10, nthreads: int = 1, want_results: bool = True):
        """Returns (seconds, scores[Q,K] f32, segs, docs, nhits, found, has_found)."""
        Q = len(queries)
        K = max(1, min(int(k), 100))
        arr = (C.c_char_p * max(1, Q))(*[q.encode("utf-8") for q in queries])
        if want_results:
            s = np.zeros((Q, K), np.float32)
            g = np.zeros((Q, K), np.uint32)
            d = np.zeros((Q, K), np.uint32)
        else:
            s = g = d = None
        nh = np.zeros(Q, np.uint32)
        fo = np.zeros(Q, np.uint64)
        hf = np.zeros(Q, np.uint8)
        sec = self.lib.orc_search_many(self.h, arr, Q, int(k), int(nthreads), _p(s), _p(g), _p(d), _p(nh), _p(fo), _p(hf))
        return sec, s, g, d, nh, fo, hf.astype(bool)

    def query_postings(self, query: str) -> int:
        return int(self.lib.orc_query_postings(self.h, query.encode("utf-8")))

    def close(self):
        if self.h:
            self.lib.orc_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def query_terms(query: str) -> List[str]:
    lib = load()
    out = (C.c_char_p * 256)()
    n = lib.orc_query_terms(query.encode("utf-8"), out, 256)
    return [out[i].decode("ascii") for i in range(n)]  # (leaks n small strings: test code)


# ---- the reference itself -------------------------------------------------------------------

def have_ref() -> bool:
    return os.path.exists(REF_ENGINE) and os.access(REF_ENGINE, os.X_OK)


def have_shim() -> bool:
    return os.path.exists(SHIM_ENGINE) and os.access(SHIM_ENGINE, os.X_OK)


def shim_search(index_dir: str, queries: Sequence[str], k: int, timeout: float = 300) -> List[str]:
    """cord19::Engine::reload()/search() of the compiled shim (reference header -> C ABI -> CUDA): the j.dump() texts."""
    with tempfile.TemporaryDirectory() as td:
        qf = os.path.join(td, "q.txt")
        with open(qf, "w") as f:
            for q in queries:
                f.write(q.replace("\\", "\\\\").replace("\n", "\\n") + "\n")
        out = os.path.join(td, "out.jsonl")
        subprocess.run([SHIM_ENGINE, "search", index_dir, qf, str(k), out], check=True, timeout=timeout)
        with open(out) as f:
            return [json.loads(line)["text"] for line in f]


def ref_write_segment(dump_path: str, segdir: str) -> None:
    subprocess.run([REF_ENGINE, "write", dump_path, segdir], check=True)


def ref_manifest(index_dir: str, names: Sequence[str]) -> None:
    subprocess.run([REF_ENGINE, "manifest", index_dir, *names], check=True)


def ref_search(index_dir: str, queries: Sequence[str], k: int, want_results: bool = True, timeout: float = 3600):
    """Runs cord19::Engine::reload() + search() for each query.  Returns (summary, results|None)."""
    with tempfile.TemporaryDirectory() as td:
        qf = os.path.join(td, "q.txt")
        with open(qf, "w") as f:
            for q in queries:  # one per line; the driver unescapes \\n and \\\\
                f.write(q.replace("\\", "\\\\").replace("\n", "\\n") + "\n")
        out = os.path.join(td, "out.jsonl") if want_results else "-"
        r = subprocess.run([REF_ENGINE, "search", index_dir, qf, str(k), out], check=True, capture_output=True,
                           text=True, timeout=timeout)
        summary = json.loads(r.stdout.strip().splitlines()[-1])
        results = None
        if want_results:
            with open(out) as f:
                results = [json.loads(line) for line in f]
        return summary, results
