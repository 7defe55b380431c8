"""Peer exchange of result blobs fused into the score kernel, through the C ABI.

World 2 on ONE device: rank 0 owns the even segments and receives, rank 1 owns the odd ones and publishes
into rank 0's gather buffer — the same kernel path as two GPUs, minus the NVLink hop, so it runs on a
single-GPU box.  test_two_devices covers real peer memory when the box has two GPUs."""
import ctypes as C
import os

import numpy as np
import pytest

import nsb200
from conftest import EDGE_QUERIES, assert_same_as_oracle, make_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def four_seg_case(workdir):
    return make_case(workdir, "xchg4", nsb200.CorpusSpec(vocab=4000), 8000, 4)


def _run_world(case, devices, queries, k, steps=3, slots=2):
    world = len(devices)
    engines = [nsb200.Engine(case.path, device=devices[r], rank=r, world=world) for r in range(world)]
    for e in engines:
        assert e.reload(), e.last_error
    xs = [nsb200.Exchange(devices[r], world, r, max_queries=max(1, len(queries)), slots=slots) for r in range(world)]
    for r in range(world):       # all-gather shape: every rank publishes to every rank (itself included)
        for p in range(world):
            xs[r].attach(xs[p])
    out = None
    for step in range(steps):
        batches = []
        for r in range(world):
            q_off, terms, has = engines[r].resolve_batch(queries)
            batches.append(engines[r].index.prepare(q_off, terms, k))
        streams = [b.stream for b in batches]
        for r in range(world):
            xs[r].launch(batches[r], step, streams[r])
        # All score kernels are enqueued before any polling kernel: with every "rank" on ONE device a poller
        # enqueued ahead of a publisher's kernel could sit in front of it in a shared hardware queue.
        for r in range(world):
            xs[r].merge(step, len(queries), k, spin=True, stream=streams[r])
        res = []
        for r in range(world):
            hits, nhits, found = xs[r].fetch(step, len(queries), k)
            res.append(nsb200.BatchResult(hits, nhits, found, has, nsb200.clamp_k(k)))
        for b in batches:
            b.close()
        for r in range(1, world):  # every receiver holds the same merged answer
            assert np.array_equal(res[r].nhits, res[0].nhits) and np.array_equal(res[r].found, res[0].found)
            assert res[r].hits.tobytes() == res[0].hits.tobytes() or all(
                np.array_equal(res[r].hits[q, :int(res[0].nhits[q])], res[0].hits[q, :int(res[0].nhits[q])])
                for q in range(len(queries)))
        out = res[0]
    for x in xs:
        x.close()
    for e in engines:
        e.close()
    return out


@pytest.mark.parametrize("k", [1, 10, 100])
def test_exchange_world2_one_device(four_seg_case, k):
    queries = nsb200.make_queries(four_seg_case.spec, 300, 1, 5, seed=31) + EDGE_QUERIES
    res = _run_world(four_seg_case, [0, 0], queries, k)
    assert_same_as_oracle(res, four_seg_case.oracle, queries, k)


def test_exchange_world4_one_device_many_steps(four_seg_case):
    queries = nsb200.make_queries(four_seg_case.spec, 200, 1, 4, seed=32)
    res = _run_world(four_seg_case, [0, 0, 0, 0], queries, 10, steps=7, slots=2)
    assert_same_as_oracle(res, four_seg_case.oracle, queries, 10)


def test_two_devices(four_seg_case):
    if nsb200._lib.load().ns_device_count() < 2:
        pytest.skip("needs two GPUs")
    queries = nsb200.make_queries(four_seg_case.spec, 500, 1, 5, seed=33) + EDGE_QUERIES
    res = _run_world(four_seg_case, [0, 1], queries, 10, steps=4)
    assert_same_as_oracle(res, four_seg_case.oracle, queries, 10)


def test_missing_peer_times_out_instead_of_hanging(four_seg_case):
    """Rank 1 never launches: rank 0's bounded wait gives up and fetch reports NS_ERR_STATE."""
    os.environ["NSB200_EXCHANGE_TIMEOUT_MS"] = "50"
    try:
        e0 = nsb200.Engine(four_seg_case.path, device=0, rank=0, world=2)
        assert e0.reload()
        x0 = nsb200.Exchange(0, 2, 0, max_queries=16, slots=2)
        x1 = nsb200.Exchange(0, 2, 1, max_queries=16, slots=2)
    finally:
        del os.environ["NSB200_EXCHANGE_TIMEOUT_MS"]
    x0.attach(x0)
    x1.attach(x0)
    queries = ["t1 t2", "t3"]
    q_off, terms, _ = e0.resolve_batch(queries)
    b = e0.index.prepare(q_off, terms, 10)
    x0.launch(b, 0, b.stream)
    x0.merge(0, len(queries), 10, spin=True, stream=b.stream)
    with pytest.raises(nsb200._lib.NsError) as ei:
        x0.fetch(0, len(queries), 10)
    assert ei.value.status == 6 and "timed out" in str(ei.value)
    b.close()
    x1.close()
    x0.close()
    e0.close()
