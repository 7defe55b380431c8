"""Per-GPU work of the 8-way sharded config on ONE GPU: one 125k-doc segment, 4096 queries, k=10.
usage: python tools/small_probe.py [ndocs]   (NSB200_WINDOW_TILES etc. pass through)"""
import os, sys, time
sys.path.insert(0, '.')
import nsb200, bench
ndocs = int(sys.argv[1]) if len(sys.argv) > 1 else 125_000
path = f"/dev/shm/nsb200_probe_{ndocs}"
if not os.path.exists(path + "/.complete"):
    nsb200.build_index(path, nsb200.SPEC_1M, ndocs, 1)
    open(path + "/.complete", "w").write("ok")
e = nsb200.Engine(path, device=0); assert e.reload()
batches = bench.make_batches(4)
prep = []
for qs in batches:
    q_off, terms, _ = e.resolve_batch(qs)
    prep.append(e.index.prepare(q_off, terms, 10))
import torch
st = torch.cuda.Stream()
for r in range(3):
    for b in prep: b.launch(st.cuda_stream)
st.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(st)
for r in range(10):
    for b in prep: b.launch(st.cuda_stream)
ev1.record(st); st.synchronize()
print(f"ndocs {ndocs} window {os.environ.get('NSB200_WINDOW_TILES','auto')}: {ev0.elapsed_time(ev1)/40:.4f} ms/batch, kernel {[round(b.kernel_ms(0),4) for b in prep]} postings {prep[0].posting_count}")
