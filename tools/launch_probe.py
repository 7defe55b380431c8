"""Host cost of ns_batch_launch vs GPU time, own stream vs torch stream."""
import sys, time
sys.path.insert(0, '.')
import bench, nsb200, torch
path = bench.ensure_index(1)
e = nsb200.Engine(path, device=0); assert e.reload()
qs = bench.make_batches(2)
bs = []
for q in qs:
    q_off, terms, _ = e.resolve_batch(q)
    bs.append(e.index.prepare(q_off, terms, 10))
for mode in ("own", "torch"):
    st = torch.cuda.Stream()
    arg = None if mode == "own" else st.cuda_stream
    for i in range(4): bs[i % 2].launch(arg)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); host = []
    for i in range(20):
        t = time.perf_counter(); bs[i % 2].launch(arg); host.append(time.perf_counter() - t)
    t1 = time.perf_counter()
    for b in bs: b.sync()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{mode}: host per launch {1e3*sum(host)/20:.3f} ms (max {1e3*max(host):.3f}), loop {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0)/20:.3f} ms/step, kernel_ms {bs[0].kernel_ms(0):.3f} merge_ms {bs[0].kernel_ms(1):.3f}", flush=True)
