"""What one long query costs the batch it lands in: the bench batch (4096 queries, 1-5 terms) as it is (NG = 1 kernel),
with ONE 40-term query (NG = 2), one 100-term query (NG = 4) and one 250-term query (NG = 8) in place of query 0.
Kernel times from the library's own CUDA events; the long query's result is checked against the oracle.
usage: python tools/long_query_probe.py   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nsb200  # noqa: E402
from oracle import oracle as orc  # noqa: E402

path = "/dev/shm/nsb200_bench/docs1000000_seg1"
if not os.path.exists(path + "/.complete"):
    nsb200.build_index(path, nsb200.SPEC_1M, 1_000_000, 1)
    open(path + "/.complete", "w").write("ok\n")
e = nsb200.Engine(path, device=0)
assert e.reload(), e.last_error
oi = orc.OracleIndex(path)
base = nsb200.make_queries(nsb200.SPEC_1M, 4096, 1, 5, seed=nsb200.QUERY_SEED)
rng = np.random.default_rng(11)
out = {}
for name, nterms in (("bench_batch", 0), ("one_40_term_query", 40), ("one_100_term_query", 100), ("one_250_term_query", 250)):
    qs = list(base)
    if nterms:
        qs[0] = " ".join(f"t{int(r)}" for r in rng.integers(200, 200_000, nterms))   # mid- and low-frequency terms
    q_off, terms, _ = e.resolve_batch(qs)
    b = e.index.prepare(q_off, terms, 10)
    ms = []
    for i in range(8):
        b.launch()
        b.sync()
        if i >= 2:
            ms.append(b.kernel_ms(0))
    h, n, f = b.fetch()
    want = oi.search(qs[0], 10)
    ok = int(f[0]) == (want["found"] or 0) and h["score"][0, :int(n[0])].view(np.uint32).tolist() == [r["score_bits"] for r in want["results"]]
    out[name] = {"kernel_ms": float(np.mean(ms)), "query0_terms": int(q_off[1] - q_off[0]), "query0_equals_oracle": bool(ok)}
    b.close()
print(json.dumps(out))
