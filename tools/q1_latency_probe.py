"""p50 of single-query (Q=1) end-to-end latency on the 1M-doc index; NSB200_WINDOW_TILES passes through."""
import os, sys, time
sys.path.insert(0, '.')
import bench, nsb200
path = bench.ensure_index(1)
e = nsb200.Engine(path, device=0); assert e.reload()
qs = bench.make_batches(1)[0][:300]
for q in qs[:20]: e.search_batch([q], 10)
lat = []
for q in qs:
    t = time.perf_counter(); e.search_batch([q], 10); lat.append(time.perf_counter() - t)
lat.sort()
print(f"window {os.environ.get('NSB200_WINDOW_TILES', 'auto')}: p50 {1e3*lat[len(lat)//2]:.3f} ms  p90 {1e3*lat[int(len(lat)*0.9)]:.3f} ms  p99 {1e3*lat[int(len(lat)*0.99)]:.3f} ms")
