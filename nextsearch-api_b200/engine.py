"""Python mirror of the reference's search surface (cord19::Engine, include/api_engine.hpp:23-91)
over the C ABI.  Names and argument meaning follow the reference: ``Engine.reload()`` returns a
bool like Engine::reload (src/api_engine.cpp:50), ``Engine.search(query, k)`` returns the same
JSON object as Engine::search (src/api_engine.cpp:369-542) minus cache flags and metadata.csv
decoration.  All scoring happens in libnsb200.so on the GPU; this file only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
import json
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import NS_MAX_K, check

HIT_DTYPE = np.dtype([("score", "<f4"), ("seg", "<u4"), ("doc", "<u4")])
QTERM_DTYPE = np.dtype([("seg", "<u4"), ("row", "<u4"), ("idf", "<f4"), ("weight", "<f4")])


def clamp_k(k: int) -> int:
    """K = max(1, min(k, 100)) — src/api_engine.cpp:377."""
    return max(1, min(int(k), NS_MAX_K))


def _cstr_array(strings: Sequence[str]):
    arr = (C.c_char_p * max(1, len(strings)))()
    for i, s in enumerate(strings):
        arr[i] = s.encode("utf-8") if isinstance(s, str) else s
    return arr


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class BatchResult:
    hits: np.ndarray       # [Q, K] HIT_DTYPE, best first; entries >= nhits[q] are undefined
    nhits: np.ndarray      # [Q] u32
    found: np.ndarray      # [Q] u64
    has_found: np.ndarray  # [Q] bool — False where the reference omits "found" (no usable terms)
    k: int


class Batch:
    """ns_batch: descriptors resident on the device; launch() enqueues the kernels only."""

    def __init__(self, index: "DeviceIndex", q_off: np.ndarray, terms: np.ndarray, k: int):
        self._lib = _lib.load()
        self.Q = int(len(q_off) - 1)
        self.k = clamp_k(k)
        q_off = np.ascontiguousarray(q_off, dtype=np.uint64)
        terms = np.ascontiguousarray(terms, dtype=QTERM_DTYPE)
        h = C.c_void_p()
        check(self._lib.ns_batch_prepare(index._h, self.Q, int(k), _ptr(q_off), _ptr(terms), C.byref(h)))
        self._h = h
        self._index = index  # keep the index alive

    @classmethod
    def _adopt(cls, handle, Q: int, k: int, keep_alive) -> "Batch":
        b = cls.__new__(cls)
        b._lib = _lib.load()
        b.Q, b.k, b._h, b._index = int(Q), clamp_k(k), handle, keep_alive
        return b

    def set_splits(self, splits: int) -> None:
        check(self._lib.ns_batch_set_splits(self._h, int(splits)))

    def launch(self, stream: Optional[int] = None) -> None:
        check(self._lib.ns_batch_launch(self._h, C.c_void_p(stream) if stream else None))

    def sync(self) -> None:
        check(self._lib.ns_batch_sync(self._h))

    def fetch(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        hits = np.zeros((self.Q, self.k), dtype=HIT_DTYPE)
        nhits = np.zeros(self.Q, dtype=np.uint32)
        found = np.zeros(self.Q, dtype=np.uint64)
        check(self._lib.ns_batch_fetch(self._h, _ptr(hits), _ptr(nhits), _ptr(found)))
        return hits, nhits, found

    def device_results(self) -> Tuple[int, int, int]:
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self._lib.ns_batch_device_results(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    @property
    def posting_count(self) -> int:
        return int(self._lib.ns_batch_posting_count(self._h))

    @property
    def stream(self) -> int:
        """The batch's own CUDA stream (launch(None) uses it)."""
        return int(self._lib.ns_batch_stream(self._h) or 0)

    @property
    def num_launches(self) -> int:
        return int(self._lib.ns_batch_num_launches(self._h))

    @property
    def upload_bytes(self) -> int:
        return int(self._lib.ns_batch_upload_bytes(self._h))

    @property
    def result_bytes(self) -> int:
        return int(self._lib.ns_batch_result_bytes(self._h))

    def kernel_ms(self, which: int = 0) -> float:
        return float(self._lib.ns_batch_last_kernel_ms(self._h, which))

    def close(self) -> None:
        if self._h:
            self._lib.ns_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceIndex:
    """ns_index: the device-resident CSR form of the engine's segments on one GPU."""

    def __init__(self, device: int = 0, _borrowed: Optional[int] = None, _owner=None):
        self._lib = _lib.load()
        self._owner = _owner
        if _borrowed is not None:
            self._h = C.c_void_p(_borrowed)
            self._own = False
        else:
            h = C.c_void_p()
            check(self._lib.ns_index_create(int(device), C.byref(h)))
            self._h = h
            self._own = True
        self.device = device

    def add_segment(self, global_seg: int, avgdl: float, doc_len: np.ndarray, term_begin: np.ndarray,
                    term_count: np.ndarray, postings: np.ndarray) -> None:
        doc_len = np.ascontiguousarray(doc_len, dtype=np.uint32)
        term_begin = np.ascontiguousarray(term_begin, dtype=np.uint64)
        term_count = np.ascontiguousarray(term_count, dtype=np.uint32)
        postings = np.ascontiguousarray(postings, dtype=np.uint32).reshape(-1, 2)
        check(self._lib.ns_index_add_segment(self._h, int(global_seg), len(doc_len), C.c_float(avgdl), _ptr(doc_len),
                                             len(term_begin), _ptr(term_begin), _ptr(term_count), _ptr(postings),
                                             postings.shape[0]))

    def commit(self) -> None:
        check(self._lib.ns_index_commit(self._h))

    def abort(self) -> None:
        check(self._lib.ns_index_abort(self._h))

    @property
    def num_segments(self) -> int:
        return int(self._lib.ns_index_num_segments(self._h))

    @property
    def device_bytes(self) -> int:
        return int(self._lib.ns_index_device_bytes(self._h))

    def prepare(self, q_off: np.ndarray, terms: np.ndarray, k: int) -> Batch:
        return Batch(self, q_off, terms, k)

    def search_batch(self, q_off: np.ndarray, terms: np.ndarray, k: int):
        """ns_search_batch: host buffers in, host buffers out (H2D + kernels + D2H)."""
        Q = len(q_off) - 1
        K = clamp_k(k)
        q_off = np.ascontiguousarray(q_off, dtype=np.uint64)
        terms = np.ascontiguousarray(terms, dtype=QTERM_DTYPE)
        hits = np.zeros((Q, K), dtype=HIT_DTYPE)
        nhits = np.zeros(Q, dtype=np.uint32)
        found = np.zeros(Q, dtype=np.uint64)
        check(self._lib.ns_search_batch(self._h, Q, int(k), _ptr(q_off), _ptr(terms), _ptr(hits), _ptr(nhits),
                                        _ptr(found)))
        return hits, nhits, found

    def close(self) -> None:
        if self._h and self._own:
            self._lib.ns_index_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Exchange:
    """ns_exchange: peer exchange of per-GPU result blobs, published by the score kernel itself (P2P stores).

    One per rank.  ``attach(peer)`` adds an exchange of this process as a destination (``attach(self)`` makes this
    rank a receiver: it waits for all ``world`` blobs of a step and merges them); ``attach_ipc`` adds a rank
    living in another process through its 64-byte handle."""

    def __init__(self, device: int, world: int, rank: int, max_queries: int, slots: int = 2):
        self._lib = _lib.load()
        h = C.c_void_p()
        check(self._lib.ns_exchange_create(int(device), int(world), int(rank), int(max_queries), int(slots), C.byref(h)))
        self._h = h
        self.device, self.world, self.rank, self.slots = device, world, rank, slots
        self._peers = []  # keep attached local peers alive

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(_lib.NS_IPC_HANDLE_BYTES)
        check(self._lib.ns_exchange_ipc_handle(self._h, buf))
        return buf.raw

    def attach(self, peer: "Exchange") -> None:
        check(self._lib.ns_exchange_attach_local(self._h, peer._h))
        if peer is not self:
            self._peers.append(peer)

    def attach_ipc(self, peer_rank: int, handle: bytes) -> None:
        check(self._lib.ns_exchange_attach_ipc(self._h, int(peer_rank), C.c_char_p(handle)))

    def launch(self, batch: "Batch", step: int, stream: Optional[int] = None) -> None:
        check(self._lib.ns_batch_launch_exchange(batch._h, self._h, int(step), C.c_void_p(stream) if stream else None))

    def merge(self, step: int, Q: int, k: int, spin: bool = True, stream: Optional[int] = None) -> None:
        check(self._lib.ns_exchange_merge(self._h, int(step), int(Q), int(k), 1 if spin else 0,
                                          C.c_void_p(stream) if stream else None))

    def fetch(self, step: int, Q: int, k: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        K = clamp_k(k)
        hits = np.zeros((Q, K), dtype=HIT_DTYPE)
        nhits = np.zeros(Q, dtype=np.uint32)
        found = np.zeros(Q, dtype=np.uint64)
        check(self._lib.ns_exchange_fetch(self._h, int(step), int(Q), int(k), _ptr(hits), _ptr(nhits), _ptr(found)))
        return hits, nhits, found

    def close(self) -> None:
        if self._h:
            self._lib.ns_exchange_destroy(self._h)
            self._h = None
        self._peers = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def merge_device(device: int, Q: int, k: int, nlists: int, d_hits: int, d_nhits: int, d_found: int, d_out_hits: int,
                 d_out_nhits: int, d_out_found: int, stream: Optional[int] = None) -> None:
    """ns_merge_device on raw device pointers ([nlists][Q][k] lists -> [Q][k])."""
    lib = _lib.load()
    check(lib.ns_merge_device(int(device), int(Q), int(k), int(nlists), C.c_void_p(d_hits), C.c_void_p(d_nhits),
                              C.c_void_p(d_found), C.c_void_p(d_out_hits), C.c_void_p(d_out_nhits),
                              C.c_void_p(d_out_found), C.c_void_p(stream) if stream else None))


def query_terms(query: str) -> List[str]:
    """tokenize + len<2/stopword filter (include/textutil.hpp:13-37, src/api_engine.cpp:391-397)."""
    lib = _lib.load()
    cap = 2 * len(query.encode("utf-8")) + 16
    buf = C.create_string_buffer(cap)
    n = lib.ns_text_query_terms(query.encode("utf-8"), buf, cap)
    if n < 0:
        raise RuntimeError("ns_text_query_terms: buffer too small")
    raw = buf.raw
    out, at = [], 0
    for _ in range(n):
        end = raw.index(b"\0", at)
        out.append(raw[at:end].decode("ascii"))
        at = end + 1
    return out


class Engine:
    """Host mirror of cord19::Engine for the search path.

    device=None builds a host-only engine (lexicon, tokenizer, resolve); its search() raises.
    rank/world select this engine's segment shard (segment i -> rank i % world).
    """

    def __init__(self, index_dir: str, device: Optional[int] = 0, rank: int = 0, world: int = 1,
                 devices: Optional[Sequence[int]] = None):
        """devices=[d0, d1, ...] builds ONE engine spanning several GPUs (ns_engine_create_multi): segment j of the
        engine's share lives on devices[j % len(devices)], results are merged on devices[0]."""
        self._lib = _lib.load()
        h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * max(1, len(devices)))(*[int(d) for d in devices])
            check(self._lib.ns_engine_create_multi(str(index_dir).encode(), len(devices), arr, C.byref(h)))
            device = devices[0] if devices else None
        else:
            check(self._lib.ns_engine_create(str(index_dir).encode(), -1 if device is None else int(device), C.byref(h)))
        self._h = h
        self.index_dir = str(index_dir)
        self.device = device
        self.devices = list(devices) if devices is not None else ([] if device is None else [device])
        self.rank, self.world = rank, world
        check(self._lib.ns_engine_set_shard(self._h, rank, world))

    # -- Engine::reload ------------------------------------------------------------------------
    def reload(self) -> bool:
        rc = self._lib.ns_engine_reload(self._h)
        if rc == _lib.NS_OK:
            return True
        if rc == 3:  # NS_ERR_IO: the reference's `return false`
            return False
        check(rc)
        return False

    @property
    def last_error(self) -> str:
        return (self._lib.ns_last_error() or b"").decode("utf-8", "replace")

    @property
    def num_segments(self) -> int:
        return int(self._lib.ns_engine_num_segments(self._h))

    def segment_name(self, i: int) -> str:
        buf = C.create_string_buffer(512)
        n = self._lib.ns_engine_segment_name(self._h, i, buf, 512)
        if n < 0:
            raise IndexError(i)
        return buf.value.decode()

    @property
    def seg_names(self) -> List[str]:
        return [self.segment_name(i) for i in range(self.num_segments)]

    def owns(self, seg: int) -> bool:
        return seg % self.world == self.rank

    def segment_stats(self, i: int) -> dict:
        N, T, P, avg = C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_float()
        check(self._lib.ns_engine_segment_stats(self._h, i, C.byref(N), C.byref(avg), C.byref(T), C.byref(P)))
        return {"N": N.value, "avgdl": avg.value, "T": T.value, "P": P.value}

    def term_stats(self, i: int, term: str) -> Tuple[int, int]:
        df, cnt = C.c_uint32(), C.c_uint32()
        check(self._lib.ns_engine_term_stats(self._h, i, term.encode(), C.byref(df), C.byref(cnt)))
        return df.value, cnt.value

    def cord_uid(self, seg: int, doc: int) -> str:
        buf = C.create_string_buffer(4096)
        n = self._lib.ns_engine_cord_uid(self._h, int(seg), int(doc), buf, 4096)
        return buf.value.decode("utf-8", "replace") if n >= 0 else ""

    @property
    def index(self) -> DeviceIndex:
        h = self._lib.ns_engine_index(self._h)
        if not h:
            raise RuntimeError("engine has no device index (created with device=None)")
        return DeviceIndex(self.device, _borrowed=h, _owner=self)

    # -- Engine::search ------------------------------------------------------------------------
    def search(self, query: str, k: int = 10) -> dict:
        q = query.encode("utf-8")
        cap = 1 << 16
        while True:
            buf = C.create_string_buffer(cap)
            need = C.c_size_t()
            check(self._lib.ns_engine_search_json(self._h, q, int(k), buf, cap, C.byref(need)))
            if need.value < cap:
                return json.loads(buf.value.decode("utf-8"))
            cap = need.value + 1

    def resolve_batch(self, queries: Sequence[str]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Host front end only: (q_off[Q+1] u64, terms QTERM_DTYPE, has_terms[Q] bool)."""
        Q = len(queries)
        q_off = np.zeros(Q + 1, dtype=np.uint64)
        has = np.zeros(max(1, Q), dtype=np.uint8)
        n = C.c_uint64()
        packed = not any("\0" in q for q in queries)
        if packed:
            z = ("\0".join(queries) + "\0").encode("utf-8") if Q else b""

            def call(terms, cap):
                return self._lib.ns_engine_resolve_batch_packed(self._h, Q, z, len(z), _ptr(q_off), terms, cap, C.byref(n),
                                                                _ptr(has))
        else:
            arr = _cstr_array(queries)

            def call(terms, cap):
                return self._lib.ns_engine_resolve_batch(self._h, Q, arr, _ptr(q_off), terms, cap, C.byref(n), _ptr(has))

        # one pass with a generous buffer; the exact count comes back in n if it was too small
        cap = max(64, 8 * Q * max(1, self.num_segments // max(1, self.world) + 1))
        terms = np.empty(cap, dtype=QTERM_DTYPE)
        rc = call(_ptr(terms), cap)
        if rc != _lib.NS_OK and n.value > cap:
            cap = n.value
            terms = np.empty(cap, dtype=QTERM_DTYPE)
            rc = call(_ptr(terms), cap)
        check(rc)
        return q_off, terms[: n.value], has[:Q].astype(bool)

    def search_batch(self, queries: Sequence[str], k: int = 10) -> BatchResult:
        """Q query strings through tokenizer, lexicon, H2D, kernels, D2H — the e2e call."""
        Q = len(queries)
        K = clamp_k(k)
        hits = np.empty((Q, K), dtype=HIT_DTYPE)
        nhits = np.empty(Q, dtype=np.uint32)
        found = np.empty(Q, dtype=np.uint64)
        has = np.zeros(max(1, Q), dtype=np.uint8)
        if any("\0" in q for q in queries):  # C strings end at NUL: keep that behaviour
            arr = _cstr_array(queries)
            check(self._lib.ns_engine_search_batch(self._h, Q, arr, int(k), _ptr(hits), _ptr(nhits), _ptr(found),
                                                   _ptr(has)))
        else:
            z = ("\0".join(queries) + "\0").encode("utf-8") if Q else b""
            check(self._lib.ns_engine_search_batch_packed(self._h, Q, z, len(z), int(k), _ptr(hits), _ptr(nhits),
                                                          _ptr(found), _ptr(has)))
        return BatchResult(hits, nhits, found, has[:Q].astype(bool), K)

    @staticmethod
    def pack_queries(queries: Sequence[str]) -> bytes:
        """The Q query strings back to back, each NUL-terminated: the host buffer ns_engine_search_batch_packed takes
        (what a request-coalescing front end accumulates)."""
        return ("\0".join(queries) + "\0").encode("utf-8") if len(queries) else b""

    @staticmethod
    def result_buffers(Q: int, k: int = 10):
        """Caller-owned output buffers for search_batch_packed(out=...): a serving thread allocates them once and
        reuses them for every call (fresh half-megabyte arrays per call are mmap'ed and unmapped by the allocator)."""
        K = clamp_k(k)
        return (np.empty((Q, K), dtype=HIT_DTYPE), np.empty(Q, dtype=np.uint32), np.empty(Q, dtype=np.uint64),
                np.zeros(max(1, Q), dtype=np.uint8))

    def search_batch_packed(self, zqueries: bytes, Q: int, k: int = 10, out=None) -> BatchResult:
        """ns_engine_search_batch_packed on an already packed host buffer; `out` = result_buffers(Q, k) to reuse."""
        K = clamp_k(k)
        hits, nhits, found, has = out if out is not None else self.result_buffers(Q, k)
        check(self._lib.ns_engine_search_batch_packed(self._h, Q, zqueries, len(zqueries), int(k), _ptr(hits), _ptr(nhits),
                                                      _ptr(found), _ptr(has)))
        return BatchResult(hits, nhits, found, has[:Q].view(np.bool_), K)

    def prepare_batch_packed(self, zqueries: bytes, Q: int, k: int = 10) -> Tuple["Batch", np.ndarray]:
        """ns_engine_prepare_batch_packed: front end + descriptors + H2D, no launch.  Returns (batch, has_found)."""
        has = np.zeros(max(1, Q), dtype=np.uint8)
        h = C.c_void_p()
        check(self._lib.ns_engine_prepare_batch_packed(self._h, int(Q), zqueries, len(zqueries), int(k), C.byref(h), _ptr(has)))
        return Batch._adopt(h, Q, k, self), has[:Q].astype(bool)

    def search_terms_batch(self, term_lists: Sequence[Sequence[Tuple[str, float]]], k: int = 10) -> BatchResult:
        """Explicit (term, qweight) lists — the reference's qterms_w — one per query (ns_engine_search_terms_batch)."""
        Q = len(term_lists)
        K = clamp_k(k)
        t_off = np.zeros(Q + 1, dtype=np.uint64)
        flat_t, flat_w = [], []
        for q, lst in enumerate(term_lists):
            for t, w in lst:
                flat_t.append(t)
                flat_w.append(w)
            t_off[q + 1] = len(flat_t)
        terms = _cstr_array(flat_t)
        weights = np.ascontiguousarray(flat_w if flat_w else [0.0], dtype=np.float32)
        hits = np.zeros((Q, K), dtype=HIT_DTYPE)
        nhits = np.zeros(Q, dtype=np.uint32)
        found = np.zeros(Q, dtype=np.uint64)
        has = np.zeros(max(1, Q), dtype=np.uint8)
        check(self._lib.ns_engine_search_terms_batch(self._h, Q, _ptr(t_off), terms, _ptr(weights), int(k), _ptr(hits),
                                                     _ptr(nhits), _ptr(found), _ptr(has)))
        return BatchResult(hits, nhits, found, has[:Q].astype(bool), K)

    def expand(self, query: str) -> Optional[List[Tuple[str, float]]]:
        """SemanticIndex::expand of the query's kept tokens; None when no embeddings are loaded."""
        cap = 1 << 16
        buf = C.create_string_buffer(cap)
        w = np.zeros(64, dtype=np.float32)
        en = C.c_int()
        n = self._lib.ns_engine_expand(self._h, query.encode("utf-8"), buf, cap, _ptr(w), 64, C.byref(en))
        if not en.value:
            return None
        if n < 0:
            raise RuntimeError("ns_engine_expand: buffer too small")
        raw, out, at = buf.raw, [], 0
        for i in range(n):
            end = raw.index(b"\0", at)
            out.append((raw[at:end].decode("utf-8", "replace"), float(w[i])))
            at = end + 1
        return out

    def search_one(self, query: str, k: int = 10):
        """One query, blocking; coalesced with other threads' queries when the coalescer runs.
        Returns (hits[:n], found or None)."""
        K = clamp_k(k)
        hits = np.zeros(K, dtype=HIT_DTYPE)
        n, f, h = C.c_uint32(), C.c_uint64(), C.c_uint8()
        check(self._lib.ns_engine_search_one(self._h, query.encode("utf-8"), int(k), _ptr(hits), C.byref(n), C.byref(f),
                                             C.byref(h)))
        return hits[: n.value], (f.value if h.value else None)

    def coalescer_start(self, max_batch: int = 4096, max_wait_us: int = 200, dispatchers: int = 2) -> None:
        check(self._lib.ns_engine_coalescer_start(self._h, int(max_batch), int(max_wait_us), int(dispatchers)))

    def coalescer_stop(self) -> None:
        check(self._lib.ns_engine_coalescer_stop(self._h))

    def coalescer_stats(self) -> dict:
        b, q, m = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(self._lib.ns_engine_coalescer_stats(self._h, C.byref(b), C.byref(q), C.byref(m)))
        return {"batches": b.value, "queries": q.value, "max_batch": m.value}

    def load_test(self, queries: Sequence[str], nthreads: int, per_thread: int, k: int = 10) -> dict:
        """ns_engine_load_test: nthreads native threads x per_thread blocking single-query calls."""
        z = ("\0".join(queries) + "\0").encode("utf-8")
        qps, p50, p99 = C.c_double(), C.c_double(), C.c_double()
        check(self._lib.ns_engine_load_test(self._h, int(nthreads), int(per_thread), len(queries), z, len(z), int(k),
                                            C.byref(qps), C.byref(p50), C.byref(p99)))
        return {"qps": qps.value, "p50_us": p50.value, "p99_us": p99.value, "threads": nthreads, "calls": nthreads * per_thread}

    def reload_stats(self) -> dict:
        t, r, d = C.c_double(), C.c_double(), C.c_double()
        pb, db = C.c_uint64(), C.c_uint64()
        check(self._lib.ns_engine_reload_stats(self._h, C.byref(t), C.byref(r), C.byref(d), C.byref(pb), C.byref(db)))
        return {"total_s": t.value, "read_upload_s": r.value, "dict_s": d.value, "posting_bytes": pb.value,
                "device_bytes": db.value}

    def search_json_text(self, query: str, k: int = 10) -> str:
        """The raw text ns_engine_search_json returns (byte-compatible with the reference's j.dump())."""
        q = query.encode("utf-8") if isinstance(query, str) else query
        cap = 1 << 16
        while True:
            buf = C.create_string_buffer(cap)
            need = C.c_size_t()
            check(self._lib.ns_engine_search_json(self._h, q, int(k), buf, cap, C.byref(need)))
            if need.value < cap:
                return buf.raw[: need.value].decode("utf-8")
            cap = need.value + 1

    def close(self) -> None:
        if self._h:
            self._lib.ns_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
