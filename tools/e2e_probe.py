import sys, time
sys.path.insert(0, '.')
import bench, nsb200
path = bench.ensure_index(1)
e = nsb200.Engine(path, device=0); assert e.reload()
qs = bench.make_batches(2)
for i in range(5):
    t1=time.perf_counter()
    r = e.search_batch(qs[i%2], 10); t2=time.perf_counter()
    print(f"search_batch {1e3*(t2-t1):.2f} ms", flush=True)
