"""Per-phase device times of the sharded step (score kernel | all-gather | merge).
torchrun --nproc-per-node N tools/dist_probe.py"""
import os, sys, ctypes as C
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import bench, nsb200
from nextsearch_api_b200.dist import ShardedSearcher
from nextsearch_api_b200 import _lib
from nextsearch_api_b200._lib import check
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
if rank == 0: bench.ensure_index(8)
dist.barrier()
s = ShardedSearcher(bench.index_path(8), lr, rank, world); assert s.reload()
batches = bench.make_batches(4)
prep = [s.prepare(qs, 10) for qs in batches]
st = torch.cuda.Stream(device=lr)
lib = _lib.load()
def step(sb, evs=None):
    stream = st.cuda_stream
    if evs: evs[0].record(st)
    sb.batch.launch(stream)
    if evs: evs[1].record(st)
    dist.all_gather_into_tensor(sb.gathered, sb.local_blob)
    if evs: evs[2].record(st)
    base = sb.out.data_ptr()
    check(lib.ns_merge_blobs_device(lr, sb.Q, sb.k, world, C.c_void_p(sb.gathered.data_ptr()), sb.blob_bytes, sb.off_n, sb.off_f,
                                    C.c_void_p(base), C.c_void_p(base + sb.off_n), C.c_void_p(base + sb.off_f), C.c_void_p(stream)))
    if evs: evs[3].record(st)
with torch.cuda.stream(st):
    for i in range(8): step(prep[i % 4])
    st.synchronize(); dist.barrier()
    allev = []
    for i in range(20):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        step(prep[i % 4], evs); allev.append(evs)
    st.synchronize()
import statistics as S
ph = [[e[j].elapsed_time(e[j+1]) for e in allev] for j in range(3)]
tot = [allev[i][0].elapsed_time(allev[i+1][0]) for i in range(19)]
print(f"rank {rank}/{world}: score {S.median(ph[0]):.3f} ms  all-gather {S.median(ph[1]):.3f} ms  merge {S.median(ph[2]):.3f} ms  step-to-step {S.median(tot):.3f} ms", flush=True)
# the product path: exchange on its own stream, overlapping the next batch's score kernel
with torch.cuda.stream(st):
    for i in range(8): s.launch(prep[i % 4])
    s.drain(); st.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(40): s.launch(prep[i % 4])
    s.drain(); e1.record(st); st.synchronize()
print(f"rank {rank}/{world}: ShardedSearcher.launch pipeline {e0.elapsed_time(e1)/40:.3f} ms/step", flush=True)
res = s.fetch(prep[3])
print(f"rank {rank}: found[0..3] {res.found[:4].tolist()}", flush=True)
dist.barrier(); dist.destroy_process_group()
