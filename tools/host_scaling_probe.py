"""Host-side scaling of the query front end (no GPU work): T threads each run ns_engine_resolve_batch_packed on a
host-only engine with NSB200_HOST_THREADS=1 (every call is serial), so the batches/s over T shows how many
cores the box really gives this process.  python tools/host_scaling_probe.py"""
import os, sys, time, threading, ctypes as C
os.environ["NSB200_HOST_THREADS"] = "1"
sys.path.insert(0, '.')
import numpy as np
import bench, nsb200
p = bench.ensure_index(8)
e = nsb200.Engine(p, device=None); assert e.reload()
bs = bench.make_batches(2)
lib = nsb200._lib.load()
z = [nsb200.Engine.pack_queries(b) for b in bs]
Q = 4096
P = lambda a: a.ctypes.data_as(C.c_void_p)
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print("cgroup cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as ex:
    print("cgroup cpu.max unavailable", ex)
def worker(n, out, i):
    q_off = np.zeros(Q + 1, np.uint64); terms = np.empty(400000, dtype=nsb200.QTERM_DTYPE); cnt = C.c_uint64(); has = np.zeros(Q, np.uint8)
    t = time.perf_counter()
    for j in range(n):
        lib.ns_engine_resolve_batch_packed(e._h, Q, z[j % 2], len(z[j % 2]), P(q_off), P(terms), 400000, C.byref(cnt), P(has))
    out[i] = time.perf_counter() - t
for T in (1, 2, 4, 8, 16, 32):
    out = [0] * T
    n = 20
    ths = [threading.Thread(target=worker, args=(n, out, i)) for i in range(T)]
    t0 = time.perf_counter()
    for t in ths: t.start()
    for t in ths: t.join()
    wall = time.perf_counter() - t0
    print(f"T={T:2d}: {T * n / wall:7.1f} batches/s  ({1e3 * wall / n:.2f} ms per batch per thread)")
