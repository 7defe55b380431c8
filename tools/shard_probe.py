"""Score-kernel time of ONE GPU's share of BASELINE configs[2] at N=8 (one 125k-doc segment of the 8-segment index),
alone on a GPU: what each rank's kernel costs without any exchange.  NSB200_WINDOW_TILES etc. are read at index
creation, so every setting runs in its own process:   python tools/shard_probe.py [world] [pub]"""
import os, sys
sys.path.insert(0, '.')
import torch
import bench, nsb200
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
pub = len(sys.argv) > 2 and sys.argv[2] == "pub"
path = bench.ensure_index(8)
e = nsb200.Engine(path, device=0, rank=0, world=world)
assert e.reload(), e.last_error
batches = bench.make_batches(4)
prep = []
for qs in batches:
    q_off, terms, has = e.resolve_batch(qs)
    prep.append(e.index.prepare(q_off, terms, 10))
st = torch.cuda.Stream()
x = None
if pub:
    x = nsb200.Exchange(0, 1, 0, 4096, slots=2)
    x.attach(x)
step = [0]
def launch(b):
    if x is None:
        b.launch(st.cuda_stream)
    else:
        x.launch(b, step[0], st.cuda_stream); x.merge(step[0], 4096, 10, spin=True, stream=st.cuda_stream); step[0] += 1
for b in prep: launch(b)
st.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(st)
for r in range(10):
    for b in prep: launch(b)
ev1.record(st); st.synchronize()
ms = ev0.elapsed_time(ev1) / 40
kms = sum(b.kernel_ms(0) for b in prep) / 4
print(f"world {world} pub {pub} WINDOW_TILES={os.environ.get('NSB200_WINDOW_TILES','auto')}: {ms:.4f} ms/step, kernel(+zeroing) {kms:.4f} ms, postings/batch {prep[0].posting_count}")
