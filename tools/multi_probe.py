"""Single-process engine over N GPUs under 1 / 4 / 16 concurrent callers, with per-phase traces.
python tools/multi_probe.py NGPUS"""
import os, sys, time, threading
os.environ["NSB200_TRACE"] = "1"
sys.path.insert(0, '.')
import bench, nsb200
n = int(sys.argv[1])
path = bench.ensure_index(8)
e = nsb200.Engine(path, devices=list(range(n)))
assert e.reload(), e.last_error
qs = bench.make_batches(4)
z = [nsb200.Engine.pack_queries(q) for q in qs]
def call(i): return e.search_batch_packed(z[i % 4], 4096, 10)
for i in range(4): call(i)
for callers in (1, 4, 16):
    team = bench.Callers(callers)
    team.run(call, 8 * callers)
    print(f"=== callers {callers}", file=sys.stderr, flush=True)
    secs, _ = team.run(call, 12 * callers)
    print(f"=== callers {callers}: {12 * callers * 4096 / secs / 1e6:.2f} M q/s", file=sys.stderr, flush=True)
    team.close()
