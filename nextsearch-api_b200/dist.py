"""Segment-sharded search with ONE PROCESS PER GPU (torchrun / torch.distributed).

Segment i lives on rank i % world; every rank receives the same query batch and scores its own segments.
What crosses GPUs is only what the reference keeps across segments — the global top-K and the found sum
(src/api_engine.cpp:435,495).  The exchange is fused into the score kernel (ns_exchange, C ABI): each rank's
kernel stores every finished query's list straight into all peers' gather buffers over NVLink (CUDA IPC
mappings of the peers' buffers), raises a flag, and every rank merges the `world` blobs under the total
order (score desc, segment asc, docId asc).  torch.distributed is plumbing here: rendezvous, the one-time
exchange of the 64-byte IPC handles, barriers.  If the IPC mapping cannot be set up (e.g. no peer access
between two GPUs) the ranks agree to fall back to one NCCL all-gather of the result blobs per batch followed
by the same merge kernel.

The single-process alternative — one engine handle spanning all GPUs of the box, which is what a C++
api_server would link — is ``Engine(index_dir, devices=[...])`` (ns_engine_create_multi).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check
from .engine import HIT_DTYPE, Batch, BatchResult, Engine, Exchange, clamp_k

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
    step: int = -1                 # exchange step of the last launch
    # NCCL fallback only
    local_blob: "object" = None    # torch uint8 view of the batch's device result blob
    gathered: "object" = None      # torch uint8 [world * blob_bytes]
    out: "object" = None           # torch uint8 [blob_bytes]: merged hits | nhits | found
    blob_bytes: int = 0
    off_n: int = 0
    off_f: int = 0


class ShardedSearcher:
    """mode: "peer" (default; falls back to "nccl" if the IPC set-up fails on any rank) or "nccl"."""

    def __init__(self, index_dir: str, device: int, rank: int, world: int, group=None, max_queries: int = 4096,
                 mode: str = "peer"):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank, self.world, self.device, self.group = rank, world, device, group
        self.engine = Engine(index_dir, device=device, rank=rank, world=world)
        self.exchange: Optional[Exchange] = None
        self.step = 0
        self.mode = "single" if world == 1 else mode
        if world > 1 and mode == "peer":
            ok = 1
            try:
                self.exchange = Exchange(device, world, rank, max_queries, slots=2)
                mine = self.exchange.ipc_handle()
            except Exception:  # noqa: BLE001
                ok, mine = 0, b""
            handles = [None] * world
            dist.all_gather_object(handles, mine, group=group)
            if ok and all(handles):
                try:
                    for p in range(world):
                        if p == rank:
                            self.exchange.attach(self.exchange)
                        else:
                            self.exchange.attach_ipc(p, handles[p])
                except Exception:  # noqa: BLE001
                    ok = 0
            else:
                ok = 0
            oks = [None] * world
            dist.all_gather_object(oks, ok, group=group)  # every rank takes the same path
            if not all(oks):
                self.exchange = None
                self.mode = "nccl"

    def reload(self) -> bool:
        return self.engine.reload()

    def prepare(self, queries, k: int, Q: Optional[int] = None) -> ShardedBatch:
        """queries: a sequence of strings, or (with Q given) the already packed NUL-separated byte buffer."""
        if isinstance(queries, (bytes, bytearray)):
            z, nq = bytes(queries), int(Q)
        else:
            z, nq = Engine.pack_queries(queries), len(queries)
        b, has = self.engine.prepare_batch_packed(z, nq, k)  # front end + descriptors + H2D in one C call
        sb = ShardedBatch(batch=b, has_found=has, Q=nq, k=clamp_k(k))
        if self.mode == "nccl":
            torch = self.torch
            lib = _lib.load()
            ptr, nbytes, off_n, off_f = C.c_void_p(), C.c_uint64(), C.c_uint64(), C.c_uint64()
            check(lib.ns_batch_result_blob(b._h, C.byref(ptr), C.byref(nbytes), C.byref(off_n), C.byref(off_f)))
            dev = torch.device("cuda", self.device)
            sb.local_blob = torch.as_tensor(_DevMem(ptr.value, nbytes.value), device=dev)
            sb.gathered = torch.empty(self.world * nbytes.value, dtype=torch.uint8, device=dev)
            sb.out = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
            sb.blob_bytes, sb.off_n, sb.off_f = nbytes.value, off_n.value, off_f.value
        return sb

    def launch(self, sb: ShardedBatch, stream: Optional[int] = None) -> None:
        """Score this rank's segments and merge with the other ranks' results; everything is enqueued on ONE
        stream (`stream`, default torch's current one), which is what makes two exchange slots sufficient.
        Collective: every rank must launch the same batches in the same order.  A batch's result must be
        fetched before the batch two launches later is launched (two slots)."""
        torch = self.torch
        if stream is None:
            # torch's default stream has handle 0, which the C ABI reads as "the batch's own stream";
            # name it explicitly (cudaStreamLegacy == 0x1)
            stream = torch.cuda.current_stream(self.device).cuda_stream or 1
        if self.mode == "single":
            sb.batch.launch(stream)
            return
        if self.mode == "peer":
            sb.step = self.step
            self.step += 1
            self.exchange.launch(sb.batch, sb.step, stream)
            self.exchange.merge(sb.step, sb.Q, sb.k, spin=True, stream=stream)
            return
        # NCCL fallback: all-gather of the blobs on the launching stream, then the merge kernel
        lib = _lib.load()
        sb.batch.launch(stream)
        ext = torch.cuda.ExternalStream(stream, device=self.device) if stream not in (0, 1) else torch.cuda.default_stream(self.device)
        with torch.cuda.stream(ext):
            self.dist.all_gather_into_tensor(sb.gathered, sb.local_blob, group=self.group)
        base = sb.out.data_ptr()
        check(lib.ns_merge_blobs_device(self.device, sb.Q, sb.k, self.world, C.c_void_p(sb.gathered.data_ptr()), sb.blob_bytes,
                                        sb.off_n, sb.off_f, C.c_void_p(base), C.c_void_p(base + sb.off_n),
                                        C.c_void_p(base + sb.off_f), C.c_void_p(stream)))

    def fetch(self, sb: ShardedBatch) -> BatchResult:
        if self.mode == "single":
            hits, nhits, found = sb.batch.fetch()
        elif self.mode == "peer":
            hits, nhits, found = self.exchange.fetch(sb.step, sb.Q, sb.k)
        else:
            self.torch.cuda.current_stream(self.device).synchronize()
            hits, nhits, found = unpack_blob(sb.out.cpu().numpy(), sb.Q, sb.k)
        return BatchResult(hits, nhits, found, sb.has_found, sb.k)

    def search_many(self, batches: Sequence[Sequence[str]], k: int = 10):
        """Several query batches, one result each, with the host front end of batch i+1 (tokenise, lexicon,
        prepare, H2D) running while the GPUs work on batch i.  Every rank must call it with the same batches."""
        out, prev = [], None
        for qs in batches:
            sb = self.prepare(qs[0], k, qs[1]) if isinstance(qs, tuple) else self.prepare(qs, k)  # (packed bytes, Q) or strings
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

    def close(self) -> None:
        if self.exchange is not None:
            self.exchange.close()
            self.exchange = None
        self.engine.close()
