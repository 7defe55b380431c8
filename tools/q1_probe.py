import sys, time
sys.path.insert(0, '.')
import bench, nsb200
path = bench.ensure_index(1)
e = nsb200.Engine(path, device=0); assert e.reload()
qs = bench.make_batches(1)[0]
for q in qs[:12]:
    t1=time.perf_counter(); r = e.search_batch([q], 10); t2=time.perf_counter()
    print(f"{q!r}: {1e3*(t2-t1):.3f} ms found {int(r.found[0])}", flush=True)
