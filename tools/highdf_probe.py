"""BASELINE configs[3] at full size: 1M docs, 4096 queries that each contain a very frequent term
(first term drawn from the 150 most frequent ranks, df > 10 % of the corpus), top-100.
Prints ms per batch, algorithmic GB/s and a 64-query parity sample against the oracle."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import nsb200, bench
from oracle import oracle as orc
path = bench.ensure_index(1)
e = nsb200.Engine(path, device=0); assert e.reload()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
HEAD = int(sys.argv[2]) if len(sys.argv) > 2 else 150
batches = [nsb200.make_queries(nsb200.SPEC_1M, 4096, 1, 5, seed=101 + i, head_ranks=HEAD) for i in range(4)]
prep = []
for qs in batches:
    q_off, terms, has = e.resolve_batch(qs)
    prep.append(e.index.prepare(q_off, terms, K))
st = torch.cuda.Stream()
for b in prep: b.launch(st.cuda_stream)
st.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(st)
for r in range(5):
    for b in prep: b.launch(st.cuda_stream)
ev1.record(st); st.synchronize()
ms = ev0.elapsed_time(ev1) / 20
gb = 8.0 * prep[0].posting_count / 1e9
print(f"head_ranks {HEAD} top-{K}: {ms:.3f} ms/batch = {4096/ms*1e3:.0f} q/s, {gb:.2f} GB algorithmic/batch = {gb/ms*1e3:.0f} GB/s")
res = e.search_batch(batches[0][:64], K)
oi = orc.OracleIndex(path)
_, s, g, d, nh, fo, hf = oi.search_many(batches[0][:64], K, nthreads=16)
ok = np.array_equal(res.nhits, nh) and np.array_equal(res.found, fo)
for q in range(64):
    n = int(nh[q])
    ok = ok and np.array_equal(res.hits["score"][q, :n].view(np.uint32), s[q, :n].view(np.uint32)) and np.array_equal(res.hits["doc"][q, :n], d[q, :n])
print(f"parity (64 queries, k={K}) bit-exact:", ok)
