"""Per-phase host timings of Engine.search_batch (NSB200_TRACE=1): python tools/e2e_trace.py [ndev] [calls]"""
import os, sys, time
os.environ["NSB200_TRACE"] = "1"
sys.path.insert(0, '.')
import bench, nsb200
ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 1
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 6
path = bench.ensure_index(1 if ndev == 1 else 8)
e = nsb200.Engine(path, devices=list(range(ndev)))
assert e.reload(), e.last_error
qs = bench.make_batches(2)
for i in range(calls):
    t1 = time.perf_counter()
    e.search_batch(qs[i % 2], 10)
    print(f"search_batch {1e3 * (time.perf_counter() - t1):.2f} ms", file=sys.stderr, flush=True)
