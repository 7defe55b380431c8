"""What the single-term queries of the bench mix cost inside the score kernel: kernel time of the full batch,
of the batch without its 1-term queries, and of the 1-term queries alone."""
import sys
sys.path.insert(0, '.')
import torch
import bench, nsb200
path = bench.ensure_index(1)
e = nsb200.Engine(path, device=0); assert e.reload()
st = torch.cuda.Stream()
def run(batches, label):
    prep = []
    for qs in batches:
        q_off, terms, has = e.resolve_batch(qs)
        prep.append(e.index.prepare(q_off, terms, 10))
    for b in prep: b.launch(st.cuda_stream)
    st.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(st)
    for r in range(5):
        for b in prep: b.launch(st.cuda_stream)
    ev1.record(st); st.synchronize()
    ms = ev0.elapsed_time(ev1) / (5 * len(prep))
    print(f"{label}: {ms:.4f} ms/batch, {sum(len(q) for q in batches)/len(batches):.0f} queries, {sum(b.posting_count for b in prep)/len(prep)/1e6:.1f} M postings")
    for b in prep: b.close()
full = bench.make_batches(4)
run(full, "full mix")
run([[q for q in b if len(q.split()) >= 2] for b in full], "without 1-term queries")
run([[q for q in b if len(q.split()) == 1] for b in full], "1-term queries only")
run([[q for q in b if len(q.split()) >= 3] for b in full], ">= 3 terms")
