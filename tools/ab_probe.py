"""Torch-free A/B timing of the score kernel on the bench workload (BASELINE configs[1]: 1M docs, 4096 queries,
k = 10): kernel time from the library's own CUDA events, parity of 128 queries against the oracle.
usage: NSB200_LIB=<library> python tools/ab_probe.py [rounds] [nseg]   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nsb200  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
nseg = int(sys.argv[2]) if len(sys.argv) > 2 else 1
path = f"/dev/shm/nsb200_bench/docs1000000_seg{nseg}"
if not os.path.exists(path + "/.complete"):
    t0 = time.time()
    nsb200.build_index(path, nsb200.SPEC_1M, 1_000_000, nseg)
    open(path + "/.complete", "w").write("ok\n")
    print(f"[ab] built index in {time.time() - t0:.1f}s", file=sys.stderr)
e = nsb200.Engine(path, device=0)
assert e.reload(), e.last_error
batches = [nsb200.make_queries(nsb200.SPEC_1M, 4096, 1, 5, seed=nsb200.QUERY_SEED + i) for i in range(8)]
prep = []
for qs in batches:
    q_off, terms, _ = e.resolve_batch(qs)
    prep.append(e.index.prepare(q_off, terms, 10))
for _ in range(2):           # one batch at a time: every batch has its own stream, overlapping launches would share the SMs
    for b in prep:
        b.launch()
        b.sync()
ms = []
for _ in range(rounds):
    for b in prep:
        b.launch()
        b.sync()
        ms.append(b.kernel_ms(0))
ms = np.array(ms)
out = {"lib": os.path.basename(nsb200._lib.LIB_PATH), "nseg": nseg, "kernel_ms_mean": float(ms.mean()), "kernel_ms_min": float(ms.min()),
       "kernel_ms_p50": float(np.median(ms)), "launches": int(ms.size), "postings": int(prep[0].posting_count)}
if os.environ.get("AB_PARITY", "1") != "0":
    from oracle import oracle as orc
    oi = orc.OracleIndex(path)
    res = e.search_batch(batches[0][:128], 10)
    _, s, g, d, nh, fo, hf = oi.search_many(batches[0][:128], 10, nthreads=os.cpu_count() or 4)
    ok = np.array_equal(res.nhits, nh) and np.array_equal(res.found, fo)
    for q in range(128):
        n = int(nh[q])
        ok = ok and np.array_equal(res.hits["score"][q, :n].view(np.uint32), s[q, :n].view(np.uint32)) and np.array_equal(res.hits["doc"][q, :n], d[q, :n])
    out["parity_128"] = bool(ok)
print(json.dumps(out))
