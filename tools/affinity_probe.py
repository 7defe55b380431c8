import os, torch, torch.distributed as dist
r=int(os.environ["RANK"]); lr=int(os.environ["LOCAL_RANK"])
print(r, "before", len(os.sched_getaffinity(0)), flush=True)
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
print(r, "after init", len(os.sched_getaffinity(0)), flush=True)
t=torch.ones(1, device=f"cuda:{lr}"); dist.all_reduce(t); torch.cuda.synchronize()
print(r, "after allreduce", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:40], flush=True)
import threading
def f(): print(r, "thread", len(os.sched_getaffinity(0)), flush=True)
th=threading.Thread(target=f); th.start(); th.join()
dist.barrier(); dist.destroy_process_group()
