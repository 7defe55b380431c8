import sys, os
sys.path.insert(0, '.')
import numpy as np
import bench, nsb200
from oracle import oracle as orc
path = bench.ensure_index(8)
qs = bench.make_batches(1)[0][:256]
oi = orc.OracleIndex(path)
_, s, g, d, nh, fo, hf = oi.search_many(qs, 10, nthreads=16)
def cmp(name, res):
    badf = [q for q in range(len(qs)) if res.found[q] != fo[q]]
    bads = [q for q in range(len(qs)) if not np.array_equal(res.hits["doc"][q,:nh[q]], d[q,:nh[q]]) or not np.array_equal(res.hits["score"][q,:nh[q]].view(np.uint32), s[q,:nh[q]].view(np.uint32))]
    print(name, "found mismatches", len(badf), "hit mismatches", len(bads), flush=True)
    for q in badf[:3]: print("   ", qs[q], int(res.found[q]), int(fo[q]))
e = nsb200.Engine(path, device=0); assert e.reload()
cmp("world=1 8seg", e.search_batch(qs, 10))
# per-segment found via single-segment shards
for w in (2, 8):
    tot = np.zeros(len(qs), np.uint64)
    for r in range(w):
        er = nsb200.Engine(path, device=0, rank=r, world=w); assert er.reload()
        q_off, terms, has = er.resolve_batch(qs)
        h, n, f = er.index.search_batch(q_off, terms, 10)
        tot += f
        er.close()
    print("world", w, "sum-of-shard found mismatches", int((tot != fo).sum()), flush=True)
