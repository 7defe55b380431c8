"""One process per rank over CUDA IPC (nextsearch-api_b200/dist.py, mode "peer"): two processes, each with its own
engine share, publish their per-query results into each other's gather buffers from inside the score kernel
and merge.  Both ranks use cuda:0 when the box has one GPU (two contexts time-slice the device), cuda:rank when
it has two.  Rendezvous over gloo on 127.0.0.1."""
import os
import socket

import numpy as np
import pytest

import nsb200
from conftest import make_case

pytestmark = pytest.mark.gpu
K = 10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, path, queries, ndev, out_q):
    import torch.distributed as dist

    from nextsearch_api_b200.dist import ShardedSearcher

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = ShardedSearcher(path, rank % ndev, rank, world, max_queries=len(queries))
        assert s.reload(), s.engine.last_error
        assert s.mode == "peer", s.mode
        last = None
        for _ in range(5):  # steps 0..4 over two slots
            last = s.search_batch(queries, K)
        res = s.search_many([queries, queries[:7], queries], K)
        if rank == 0:
            out_q.put((last.hits.tobytes(), last.nhits.tobytes(), last.found.tobytes(), res[1].nhits.tobytes(),
                       res[2].hits.tobytes()))
        dist.barrier()
        s.close()
    finally:
        dist.destroy_process_group()


def test_two_processes_exchange_over_cuda_ipc(workdir):
    import torch.multiprocessing as mp

    case = make_case(workdir, "ipc4", nsb200.CorpusSpec(vocab=3000), 6000, 4)
    queries = nsb200.make_queries(case.spec, 200, 1, 4, seed=81) + ["t3 t3", "zzzz", "the of"]
    ndev = min(2, nsb200._lib.load().ns_device_count())
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case.path, queries, ndev, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    got = out_q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    hits = np.frombuffer(got[0], dtype=nsb200.HIT_DTYPE).reshape(len(queries), K)
    nhits = np.frombuffer(got[1], dtype=np.uint32)
    found = np.frombuffer(got[2], dtype=np.uint64)
    _, s, g, d, nh, fo, hf = case.oracle.search_many(queries, K, nthreads=4)
    assert np.array_equal(nhits, nh) and np.array_equal(found, fo)
    for q in range(len(queries)):
        n = int(nh[q])
        assert np.array_equal(hits["score"][q, :n].view(np.uint32), s[q, :n].view(np.uint32)), queries[q]
        assert np.array_equal(hits["doc"][q, :n], d[q, :n]) and np.array_equal(hits["seg"][q, :n], g[q, :n])
    assert np.array_equal(np.frombuffer(got[3], dtype=np.uint32), nh[:7])
    assert got[4] == got[0] or np.array_equal(np.frombuffer(got[4], dtype=nsb200.HIT_DTYPE).reshape(len(queries), K)["doc"][0, :int(nh[0])], d[0, :int(nh[0])])
