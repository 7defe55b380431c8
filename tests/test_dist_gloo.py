"""world_size-2 test of the segment-sharded path on CPU (gloo): shard ownership, the per-rank
result blob layout and its exchange with ONE all-gather, and the merge order.

There is no CPU scoring path in the product, so each rank's local top-k lists come from the oracle
run over that rank's own segments (a rank-local index view); what is under test is the host logic of
nextsearch-api_b200/dist.py + ns_engine_set_shard/ns_engine_resolve_batch.  The CUDA merge kernel
itself is covered by the -m gpu tests and the multi-GPU bench.
"""
import os
import socket

import numpy as np
import pytest

import nsb200
from nextsearch_api_b200 import dist as nsdist
from nextsearch_api_b200.engine import HIT_DTYPE
from oracle import oracle as orc

WORLD = 2
K = 10


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_view(full_dir, view_dir, names):
    """An index directory holding only `names` (symlinked), with its own manifest."""
    os.makedirs(os.path.join(view_dir, "segments"), exist_ok=True)
    for n in names:
        dst = os.path.join(view_dir, "segments", n)
        if not os.path.exists(dst):
            os.symlink(os.path.join(full_dir, "segments", n), dst)
    nsb200.write_manifest(view_dir, list(names))


def merge_reference(lists, k):
    """lists: [(score, seg, doc)] per rank, each best first -> global best-first under
    (score desc, seg asc, doc asc)."""
    allh = [h for l in lists for h in l]
    allh.sort(key=lambda h: (-float(h[0]), int(h[1]), int(h[2])))
    return allh[:k]


def _worker(rank, world, port, full_dir, work, queries, out_q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- shard ownership through the product's host engine (no device) ---
        eng = nsb200.Engine(full_dir, device=None, rank=rank, world=world)
        assert eng.reload(), eng.last_error
        nseg = eng.num_segments
        own = [i for i in range(nseg) if nsdist.owner_of_segment(i, world) == rank]
        assert [i for i in range(nseg) if eng.owns(i)] == own
        q_off, terms, has = eng.resolve_batch(queries)
        assert set(np.unique(terms["seg"]).tolist()) <= set(own)

        # --- local lists from the oracle over this rank's segments, global segment ids restored ---
        names = [eng.segment_name(i) for i in own]
        view = os.path.join(work, f"view{rank}")
        _rank_view(full_dir, view, names)
        oi = orc.OracleIndex(view)
        _, s, g, d, nh, fo, hf = oi.search_many(queries, K, nthreads=2)
        hits = np.zeros((len(queries), K), HIT_DTYPE)
        hits["score"], hits["doc"] = s, d
        hits["seg"] = np.asarray(own, np.uint32)[g]
        blob = nsdist.pack_blob(hits, nh, fo, K)
        total, off_n, off_f = nsdist.blob_layout(len(queries), K)
        assert blob.nbytes == total

        # --- ONE all-gather of the blobs, as ShardedSearcher.launch does on the device ---
        local = torch.from_numpy(blob)
        gathered = torch.empty(world * total, dtype=torch.uint8)
        dist.all_gather_into_tensor(gathered, local)
        parts = [nsdist.unpack_blob(gathered[r * total:(r + 1) * total].numpy(), len(queries), K) for r in range(world)]
        merged = []
        for q in range(len(queries)):
            lists = [[tuple(p[0][q, i]) for i in range(int(p[1][q]))] for p in parts]
            merged.append((merge_reference(lists, K), int(sum(int(p[2][q]) for p in parts))))
        if rank == 0:
            out_q.put(merged)
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_exchange_and_merge(workdir):
    import torch.multiprocessing as mp

    spec = nsb200.CorpusSpec(vocab=3000)
    full = os.path.join(workdir, "dist4")
    if not os.path.isdir(full):
        nsb200.build_index(full, spec, 4000, 4)
    queries = nsb200.make_queries(spec, 64, 1, 4) + ["t3 t3", "zzzz", "the of"]
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    work = os.path.join(workdir, "dist_views")
    os.makedirs(work, exist_ok=True)
    procs = [ctx.Process(target=_worker, args=(r, WORLD, port, full, work, queries, out_q)) for r in range(WORLD)]
    for p in procs:
        p.start()
    merged = out_q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the global answer: the oracle over all four segments
    oi = orc.OracleIndex(full)
    _, s, g, d, nh, fo, hf = oi.search_many(queries, K, nthreads=2)
    for q, (hits, found) in enumerate(merged):
        assert len(hits) == int(nh[q]), queries[q]
        assert found == int(fo[q]), queries[q]
        for i, (sc, sg, dc) in enumerate(hits):
            assert np.float32(sc).view(np.uint32) == s[q, i].view(np.uint32), (queries[q], i)
            assert (int(sg), int(dc)) == (int(g[q, i]), int(d[q, i])), (queries[q], i)


def test_blob_roundtrip():
    Q = 5
    hits = np.zeros((Q, K), HIT_DTYPE)
    hits["score"] = np.arange(Q * K, dtype=np.float32).reshape(Q, K)
    hits["seg"] = 3
    hits["doc"] = np.arange(Q * K, dtype=np.uint32).reshape(Q, K)
    nh = np.arange(Q, dtype=np.uint32)
    fo = np.arange(Q, dtype=np.uint64) * 1000
    h2, n2, f2 = nsdist.unpack_blob(nsdist.pack_blob(hits, nh, fo, K), Q, K)
    assert np.array_equal(h2, hits) and np.array_equal(n2, nh) and np.array_equal(f2, fo)
