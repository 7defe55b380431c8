"""Segment-sharded search over the GPUs of one box.

One process per GPU (torch.distributed, NCCL).  Segment i lives on rank i % world; every rank
receives the same query batch, scores its own segments, and the per-rank result blobs
(hits | nhits | found) are exchanged with ONE all-gather per batch and merged on every rank by
ns_merge_blobs_device under the total order (score desc, segment asc, docId asc).  Scores are
segment-local in the reference (src/api_engine.cpp:461,478), so no other exchange exists.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check
from .engine import HIT_DTYPE, Batch, BatchResult, Engine, clamp_k

import ctypes as C

_ALIGN = 256


def _up(n: int) -> int:
    return (n + _ALIGN - 1) // _ALIGN * _ALIGN


def blob_layout(Q: int, k: int) -> Tuple[int, int, int]:
    """(bytes, off_nhits, off_found) of one rank's result blob — mirrors ns_batch_prepare."""
    k = clamp_k(k)
    sz_hits = _up(max(1, Q * k) * HIT_DTYPE.itemsize)
    sz_n = _up(max(1, Q) * 4)
    sz_f = _up(max(1, Q) * 8)
    return sz_hits + sz_n + sz_f, sz_hits, sz_hits + sz_n


def pack_blob(hits: np.ndarray, nhits: np.ndarray, found: np.ndarray, k: int) -> np.ndarray:
    Q = len(nhits)
    total, off_n, off_f = blob_layout(Q, k)
    blob = np.zeros(total, np.uint8)
    blob[: Q * clamp_k(k) * HIT_DTYPE.itemsize] = np.ascontiguousarray(hits, HIT_DTYPE).view(np.uint8).reshape(-1)
    blob[off_n: off_n + Q * 4] = np.ascontiguousarray(nhits, np.uint32).view(np.uint8)
    blob[off_f: off_f + Q * 8] = np.ascontiguousarray(found, np.uint64).view(np.uint8)
    return blob


def unpack_blob(blob: np.ndarray, Q: int, k: int):
    k = clamp_k(k)
    _, off_n, off_f = blob_layout(Q, k)
    hits = blob[: Q * k * HIT_DTYPE.itemsize].view(HIT_DTYPE).reshape(Q, k)
    nhits = blob[off_n: off_n + Q * 4].view(np.uint32)
    found = blob[off_f: off_f + Q * 8].view(np.uint64)
    return hits, nhits, found


def owner_of_segment(seg: int, world: int) -> int:
    return seg % world


class _DevMem:
    """Expose raw device memory to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


@dataclass
class ShardedBatch:
    batch: Batch
    has_found: np.ndarray
    Q: int
    k: int
    local_blob: "object"      # torch uint8 view of the batch's device result blob
    gathered: "object"        # torch uint8 [world * blob_bytes]
    out: "object"             # torch uint8 [blob_bytes]: merged hits | nhits | found (same layout as a rank blob)
    blob_bytes: int
    off_n: int
    off_f: int
    comm_done: "object" = None  # torch event: this batch's all-gather + merge have finished (exchange stream)


class ShardedSearcher:
    def __init__(self, index_dir: str, device: int, rank: int, world: int, group=None):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank, self.world, self.device, self.group = rank, world, device, group
        self.engine = Engine(index_dir, device=device, rank=rank, world=world)
        self.comm_stream = None  # exchange stream (all-gather + merge), created on first sharded launch
        self._last_comm = None

    def reload(self) -> bool:
        return self.engine.reload()

    def prepare(self, queries: Sequence[str], k: int) -> ShardedBatch:
        torch = self.torch
        q_off, terms, has = self.engine.resolve_batch(queries)
        b = self.engine.index.prepare(q_off, terms, k)
        lib = _lib.load()
        ptr, nbytes, off_n, off_f = C.c_void_p(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib.ns_batch_result_blob(b._h, C.byref(ptr), C.byref(nbytes), C.byref(off_n), C.byref(off_f)))
        dev = torch.device("cuda", self.device)
        local = torch.as_tensor(_DevMem(ptr.value, nbytes.value), device=dev)
        Q, K = len(queries), clamp_k(k)
        return ShardedBatch(
            batch=b, has_found=has, Q=Q, k=K, local_blob=local,
            gathered=torch.empty(self.world * nbytes.value, dtype=torch.uint8, device=dev),
            out=torch.empty(nbytes.value, dtype=torch.uint8, device=dev),
            blob_bytes=nbytes.value, off_n=off_n.value, off_f=off_f.value)

    def launch(self, sb: ShardedBatch) -> None:
        """score+top-k on this rank's segments (torch's current stream) -> all-gather -> merge.

        With more than one rank the exchange (NCCL all-gather of the result blobs + device merge) runs on
        its own stream behind an event, so the NEXT batch's score kernel — which does not depend on it —
        is not held back by the latency-bound collective: the exchange of batch i lands in the tail of the
        score kernel of batch i+1, where SMs are idle anyway.  `fetch` / `drain` order later work after it."""
        torch, dist = self.torch, self.dist
        cur = torch.cuda.current_stream(self.device)
        # torch's default stream has handle 0, which ns_batch_launch reads as "the batch's own stream";
        # name it explicitly (cudaStreamLegacy == 0x1) so that everything is ordered on torch's streams
        stream = cur.cuda_stream or 1
        if sb.comm_done is not None:
            cur.wait_event(sb.comm_done)  # the previous exchange of this batch still reads its result blob
        sb.batch.launch(stream)
        lib = _lib.load()
        base = sb.out.data_ptr()
        if self.world == 1:
            check(lib.ns_merge_blobs_device(self.device, sb.Q, sb.k, 1, C.c_void_p(sb.local_blob.data_ptr()),
                                            sb.blob_bytes, sb.off_n, sb.off_f, C.c_void_p(base),
                                            C.c_void_p(base + sb.off_n), C.c_void_p(base + sb.off_f), C.c_void_p(stream)))
            return
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream(device=self.device)
        cs = self.comm_stream
        scored = torch.cuda.Event()
        scored.record(cur)
        cs.wait_event(scored)
        with torch.cuda.stream(cs):
            dist.all_gather_into_tensor(sb.gathered, sb.local_blob, group=self.group)
            check(lib.ns_merge_blobs_device(self.device, sb.Q, sb.k, self.world, C.c_void_p(sb.gathered.data_ptr()),
                                            sb.blob_bytes, sb.off_n, sb.off_f, C.c_void_p(base),
                                            C.c_void_p(base + sb.off_n), C.c_void_p(base + sb.off_f),
                                            C.c_void_p(cs.cuda_stream)))
            sb.comm_done = torch.cuda.Event()
            sb.comm_done.record(cs)
        sb.gathered.record_stream(cs)
        sb.out.record_stream(cs)
        self._last_comm = sb.comm_done

    def drain(self) -> None:
        """Order torch's current stream after every exchange launched so far."""
        if self._last_comm is not None:
            self.torch.cuda.current_stream(self.device).wait_event(self._last_comm)

    def fetch(self, sb: ShardedBatch) -> BatchResult:
        if sb.comm_done is not None:
            self.torch.cuda.current_stream(self.device).wait_event(sb.comm_done)
        hits, nhits, found = unpack_blob(sb.out.cpu().numpy(), sb.Q, sb.k)  # one D2H copy
        return BatchResult(hits, nhits, found, sb.has_found, sb.k)

    def search_many(self, batches: Sequence[Sequence[str]], k: int = 10):
        """Several query batches, one result each, with the host front end of batch i+1 (tokenise, lexicon,
        prepare, H2D) running while the GPU works on batch i.  Every rank must call it with the same batches
        (one collective per batch, issued in the same order everywhere)."""
        out, prev = [], None
        for qs in batches:
            sb = self.prepare(qs, k)
            self.launch(sb)
            if prev is not None:
                out.append(self.fetch(prev))
                prev.batch.close()
            prev = sb
        if prev is not None:
            out.append(self.fetch(prev))
            prev.batch.close()
        return out

    def search_batch(self, queries: Sequence[str], k: int = 10) -> BatchResult:
        sb = self.prepare(queries, k)
        self.launch(sb)
        res = self.fetch(sb)
        sb.batch.close()
        return res
