#!/usr/bin/env python
"""bench.py — BM25 top-10 queries/sec over 1M docs (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm (oracle/_ref)

A "step" is one batch of 4096 lexicon-resolved queries (1-5 Zipf-sampled terms, k=10) scored and
top-k selected against the resident index.
  N=1 : BASELINE configs[1] — 1M docs, one segment.
  N>1 : BASELINE configs[2] — the same 1M docs split into 8 segments, segment i on rank i % N; each rank's score
        kernel publishes its per-query results into every peer's gather buffer (P2P stores over NVLink, CUDA
        IPC) and every rank merges ("strong" scaling: the corpus is fixed at 1M docs).
value    = queries/s with the query descriptors already resident in HBM (kernels [+ exchange + merge] only)
e2e      = queries/s through the C ABI with HOST buffers: query strings in, hits out, per call tokenise, lexicon,
           prepare, H2D, kernels, D2H.  N=1: ns_engine_search_batch_packed; N>1: the same call on ONE engine handle
           spanning the N GPUs (ns_engine_create_multi), driven by rank 0 — the shape a C++ api_server links —
           next to the one-process-per-GPU pipeline (e2e.per_rank_value).
roofline = algorithmic posting bytes (8 B x Σ LexEntry.count) / score-kernel time
Extra keys (same line): per-step time distribution, the other BASELINE workloads that fit this run
(configs[3] high-df top-100, the 8-segment index on one GPU, configs[4] 8M docs at N=8), coalescer numbers.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bm25_top10_queries_per_sec_1M_docs"
UNIT = "queries/s"
NDOCS = int(os.environ.get("NSB200_BENCH_DOCS", "1000000"))   # default = BASELINE configs[1]/[2]
SEGS_SHARDED = int(os.environ.get("NSB200_BENCH_SEGS", "8"))
BATCH_Q = 4096
TOPK = 10
BENCH_DIR = os.environ.get("NSB200_BENCH_DIR", "/dev/shm/nsb200_bench")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def index_path(nseg: int, ndocs: int = None) -> str:
    return os.path.join(BENCH_DIR, f"docs{ndocs or NDOCS}_seg{nseg}")


def ensure_index(nseg: int, ndocs: int = None) -> str:
    import nsb200

    ndocs = ndocs or NDOCS
    path = index_path(nseg, ndocs)
    marker = os.path.join(path, ".complete")
    if not os.path.exists(marker):
        t0 = time.time()
        nsb200.build_index(path, nsb200.SPEC_1M, ndocs, nseg)
        open(marker, "w").write("ok\n")
        log(f"[bench] built {ndocs}-doc index with {nseg} segment(s) in {time.time() - t0:.1f}s at {path}")
    return path


def make_batches(nb: int, head_ranks: int = 0, seed0: int = None):
    import nsb200

    seed0 = nsb200.QUERY_SEED if seed0 is None else seed0
    return [nsb200.make_queries(nsb200.SPEC_1M, BATCH_Q, 1, 5, seed=seed0 + i, head_ranks=head_ranks) for i in range(nb)]


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML during the timed region."""

    def __init__(self, device: int, period: float = 0.005):
        super().__init__(daemon=True)
        self.device, self.period = device, period
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[device]) if visible and visible.split(",")[device].isdigit() else device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as ex:  # noqa: BLE001
            log(f"[bench] NVML unavailable: {ex}")
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# The `ncu --set full` capture of the score kernel on this workload that roofline.traffic quotes.  It is a STATIC
# figure from the named file (captured on the commit recorded in profiles/README.md), not measured in this run.
NCU_SUMMARY = os.path.join(ROOT, "profiles", "r2_final_ncu.txt")
NCU_SUMMARY_FALLBACK = os.path.join(ROOT, "profiles", "r1_v12_final_ncu.txt")


def ncu_summary_path():
    return NCU_SUMMARY if os.path.exists(NCU_SUMMARY) else NCU_SUMMARY_FALLBACK


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE score-kernel launch on this workload, from the
    committed `ncu --set full` capture of the same command (profiles/); None when the file is absent."""
    try:
        tot = 0.0
        for ln in open(ncu_summary_path()):
            f = ln.split()
            if f and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[f[1]]
        return tot or None
    except Exception:  # noqa: BLE001
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# reference arm
# ------------------------------------------------------------------------------------------------

def unique_queries(batches, n):
    seen, out = set(), []
    for b in batches:
        for q in b:
            if q not in seen:
                seen.add(q)
                out.append(q)
                if len(out) == n:
                    return out
    return out


def run_ref_replicas(path, queries_per_replica, k):
    """One oracle/_ref/ref_engine process per replica; returns per-replica latency lists (seconds)."""
    from oracle import oracle as orc

    procs, files = [], []
    td = tempfile.mkdtemp(prefix="nsb200_refarm_")
    for r, qs in enumerate(queries_per_replica):
        qf = os.path.join(td, f"q{r}.txt")
        lf = os.path.join(td, f"lat{r}.txt")
        with open(qf, "w") as f:
            f.write("\n".join(qs) + "\n")
        procs.append(subprocess.Popen([orc.REF_ENGINE, "search", path, qf, str(k), "-", "0", "-1", lf],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True))
        files.append(lf)
    lats = []
    for p, lf in zip(procs, files):
        p.communicate()
        if p.returncode != 0:
            raise RuntimeError("ref_engine failed")
        lats.append([float(x) for x in open(lf).read().split()])
    shutil.rmtree(td, ignore_errors=True)
    return lats


def reference_arm(args, rank, world):
    """The reference's own Engine::search (unmodified sources, oracle/_ref) on the host cores.
    The engine serialises on one mutex, so 'all the host threads it can use' = independent
    replicas, one per core, each serving a disjoint slice of every step's sample."""
    if rank != 0:
        return
    from oracle import oracle as orc

    nseg = 1 if args.gpus == 1 else SEGS_SHARDED
    path = ensure_index(nseg)
    cores = os.cpu_count() or 1
    replicas = max(1, min(cores, args.ref_replicas or cores))
    per = args.ref_queries_per_replica
    W, K = args.warmup, args.steps
    need = replicas * per * (W + K)
    qs = unique_queries(make_batches(max(1, (need + BATCH_Q - 1) // BATCH_Q + 1)), need)
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, nseg),
            # the timed searches run in ref_engine processes (the reference's own code); the INDEX they read was
            # written by this repo's generator/writer, whose files are sha256-identical to the reference
            # SegmentWriter's (tests/test_oracle_golden.py) — that is the only product code on this arm
            "index_built_by": "nsb200.build_index (libnsb200.so corpus writer, byte-identical to the reference SegmentWriter)"}
    if not orc.have_ref():
        # the compiled reference did not travel: time the oracle port on all cores instead
        oi = orc.OracleIndex(path)
        step_q = replicas * per
        times = []
        for s in range(W + K):
            sec, *_ = oi.search_many(qs[s * step_q:(s + 1) * step_q], TOPK, nthreads=cores, want_results=False)
            if s >= W:
                times.append(sec)
        total = sum(times)
        val = K * step_q / total
        kind, used = "port", cores
    else:
        slices = [[q for s in range(W + K) for q in qs[(s * replicas + r) * per:(s * replicas + r + 1) * per]]
                  for r in range(replicas)]
        lats = run_ref_replicas(path, slices, TOPK)
        step_times = []
        for s in range(W, W + K):
            step_times.append(max(sum(l[s * per:(s + 1) * per]) for l in lats))
        total = sum(step_times)
        step_q = replicas * per
        val = K * step_q / total
        kind, used = "reference", replicas
    line.update({
        "value": val, "ms_per_step": 1e3 * total / K,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": used, "kind": kind,
                         "sample": f"{step_q} unique queries/step ({per} per replica x {used}), k={TOPK}, "
                                   f"same 1M-doc index and query distribution"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, nseg, ndocs=None, k=TOPK, queries="1-5 Zipf-sampled terms"):
    ndocs = ndocs or NDOCS
    if ndocs != 1_000_000 or nseg not in (1, 8):
        wl = f"BASELINE configs[4]-style: {ndocs} docs in {nseg} segments, segment-sharded, peer-memory exchange + merge"
    elif nseg == 1:
        wl = "BASELINE configs[1]: 1M docs, 1 segment"
    else:
        wl = "BASELINE configs[2]: 1M docs in 8 segments, segment-sharded, peer-memory exchange + merge"
    return {"workload": wl,
            "docs": ndocs, "segments": nseg, "batch_queries": BATCH_Q, "terms_per_query": "1-5", "k": k,
            "queries": queries,
            "corpus": "shifted Zipf s=1 q=25, V=400000, doc_len U[100,250), seed 20260101; query seeds 7+i",
            "l2_policy": "index (1.1 GB of resident postings) is larger than L2; distinct query batches rotate between steps",
            "parallelism": f"segments%{n_gpus}"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def timed_steps(torch, step, sync_all, stream, W, K):
    """W untimed steps, then exactly K steps between two CUDA events on the launching stream; one event per
    step boundary gives the per-step distribution.  Returns (total ms, [ms per step])."""
    for i in range(W):
        step(i)
    sync_all()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    evs[0].record(stream)
    for i in range(K):
        step(W + i)
        evs[i + 1].record(stream)
    sync_all()
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
    return evs[0].elapsed_time(evs[K]), per


def dist_stats(per):
    s = sorted(per)
    return {"p50": s[len(s) // 2], "min": s[0], "max": s[-1], "p90": s[min(len(s) - 1, int(0.9 * len(s)))]}


def device_workload(torch, eng, batches, k, stream, W, K):
    """Device-resident run of one workload on a single-GPU engine: prepared batches, K timed launches.
    Returns dict(value, ms_per_step, per_step, roofline pieces)."""
    prepared = []
    for qs in batches:
        q_off, terms, _ = eng.resolve_batch(qs)
        prepared.append(eng.index.prepare(q_off, terms, k))
    nb = len(prepared)

    def step(i):
        prepared[i % nb].launch(stream.cuda_stream)

    def sync_all():
        stream.synchronize()

    ms_total, per = timed_steps(torch, step, sync_all, stream, W, K)
    k_ms = [b.kernel_ms(0) for b in prepared]
    used = [i for i in range(nb) if k_ms[i] > 0]
    alg = [8.0 * prepared[i].posting_count for i in used]
    ach = [alg[j] / (k_ms[i] * 1e-3) / 1e9 for j, i in enumerate(used)]
    out = {"value": K * len(batches[0]) / (ms_total / 1e3), "ms_per_step": ms_total / K, "step_ms": dist_stats(per),
           "achieved_gbs": sum(ach) / max(1, len(ach)), "alg_bytes": sum(alg) / max(1, len(alg)),
           "kernel_ms": sum(k_ms[i] for i in used) / max(1, len(used)), "launches_per_step": prepared[0].num_launches,
           "upload_bytes": prepared[0].upload_bytes, "result_bytes": prepared[0].result_bytes}
    for b in prepared:
        b.close()
    return out


class Callers:
    """`callers` persistent host threads issuing calls.  The threads outlive one run: a thread's first CUDA call on
    every device of a multi-GPU engine binds it to that device's context, which belongs to warm-up, not to the
    timed region."""

    def __init__(self, callers: int):
        from concurrent.futures import ThreadPoolExecutor

        self.n = max(1, callers)
        self.pool = ThreadPoolExecutor(max_workers=self.n) if self.n > 1 else None

    def run(self, call, n_calls):
        """n_calls invocations of call(i).  Returns (seconds, result of the last call)."""
        if self.pool is None:
            t0 = time.perf_counter()
            last = None
            for i in range(n_calls):
                last = call(i)
            return time.perf_counter() - t0, last
        nxt = [0]
        lock = threading.Lock()
        last_box = [None]

        def worker():
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= n_calls:
                    return
                r = call(i)
                if i == n_calls - 1:
                    last_box[0] = r

        t0 = time.perf_counter()
        futs = [self.pool.submit(worker) for _ in range(self.n)]
        for f in futs:
            f.result()
        return time.perf_counter() - t0, last_box[0]

    def close(self):
        if self.pool is not None:
            self.pool.shutdown()


def run_callers(call, n_calls, callers):
    c = Callers(callers)
    try:
        return c.run(call, n_calls)
    finally:
        c.close()


def measure_e2e(call, e2e_steps, callers):
    """single-caller and `callers`-thread throughput of call(i); returns (single_s, multi_s, multi_calls, last)."""
    for i in range(2):
        call(i)
    single_s, last = run_callers(call, e2e_steps, 1)
    team = Callers(callers)
    n = e2e_steps * max(1, callers // 2)
    team.run(call, 8 * callers)      # warm-up at full concurrency: the engine's pooled per-call resources (pinned staging,
                                     # device blobs, exchange groups) are allocated on first use and reused afterwards
    multi_s, last2 = team.run(call, n)
    team.close()
    return single_s, multi_s, n, (last2 if last2 is not None else last)


def own_copy(res):
    """A BatchResult that does not alias a caller thread's reusable buffers."""
    import nsb200

    return nsb200.BatchResult(res.hits.copy(), res.nhits.copy(), res.found.copy(), res.has_found.copy(), res.k)


def oracle_parity(path, queries, result, k, nchk=256):
    import numpy as np

    from oracle import oracle as orc

    oi = orc.OracleIndex(path)
    nchk = min(nchk, len(queries))
    qs = queries[:nchk]
    _, s, g, d, nh, fo, hf = oi.search_many(qs, k, nthreads=os.cpu_count() or 1)
    ok = bool(np.array_equal(result.nhits[:nchk], nh) and np.array_equal(result.found[:nchk], fo))
    for q in range(nchk):
        n = int(nh[q])
        same = (np.array_equal(result.hits["score"][q, :n].view(np.uint32), s[q, :n].view(np.uint32))
                and np.array_equal(result.hits["doc"][q, :n], d[q, :n])
                and np.array_equal(result.hits["seg"][q, :n], g[q, :n]))
        if not same and ok:
            log(f"[bench] parity: query {q} {qs[q]!r} differs from the oracle")
        ok = ok and same
    return {"queries": nchk, "k": k, "bit_exact_vs_oracle": bool(ok)}


def ours(args, rank, world, local_rank):
    import numpy as np
    import torch

    import nsb200
    from nextsearch_api_b200.dist import ShardedSearcher

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dist = None
    gloo = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        gloo = dist.new_group(backend="gloo")  # host-side barriers that do not put a spinning kernel on the GPUs
    dev = local_rank
    torch.cuda.set_device(dev)
    nseg = 1 if world == 1 else SEGS_SHARDED

    if rank == 0:
        ensure_index(nseg)
    if world > 1:
        dist.barrier()
    path = index_path(nseg)

    t0 = time.time()
    searcher = ShardedSearcher(path, dev, rank, world, max_queries=BATCH_Q, mode=args.exchange)
    assert searcher.reload(), searcher.engine.last_error
    eng = searcher.engine
    rstats = eng.reload_stats()
    log(f"[bench] rank {rank}: reload+upload {time.time() - t0:.2f}s, device index {eng.index.device_bytes / 1e9:.2f} GB "
        f"({rstats['device_bytes'] / max(1, rstats['posting_bytes']):.2f}x the posting bytes), exchange mode {searcher.mode}")

    W, K = args.warmup, args.steps
    nb = max(1, min(args.distinct_batches, W + K))
    batches = make_batches(nb)
    stream = torch.cuda.Stream(device=dev)
    extra = {}
    with torch.cuda.stream(stream):
        prepared = [searcher.prepare(qs, TOPK) for qs in batches]
        total_postings = [b.batch.posting_count for b in prepared]
        launches_per_step = prepared[0].batch.num_launches + (2 if world > 1 else 0)  # + flag poll + merge

        def step(i):
            searcher.launch(prepared[i % nb], stream.cuda_stream)

        def sync_all():
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        sampler = ClockSampler(dev)
        sampler.start()
        ms_total, per_step = timed_steps(torch, step, sync_all, stream, W, K)
        sampler.stop_flag = True
        sampler.join()
    if world > 1:
        t = torch.tensor([ms_total], device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = K * BATCH_Q / (ms_total / 1e3)

    # per-launch score-kernel time (CUDA events recorded by the library on the launching stream)
    k_ms = [b.batch.kernel_ms(0) for b in prepared]
    used = [i for i in range(nb) if k_ms[i] > 0]
    alg_bytes = [8.0 * total_postings[i] for i in used]
    ach = [alg_bytes[j] / (k_ms[i] * 1e-3) / 1e9 for j, i in enumerate(used)]
    ach_local = sum(ach) / max(1, len(ach))
    peak, peak_src = measured_peak()
    h2d = int(prepared[0].batch.upload_bytes)
    d2h = int(prepared[0].batch.result_bytes)

    if args.profile_mode:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
                              "warmup": W, "ms_per_step": ms_total / K, "profile_mode": True, "step_ms": dist_stats(per_step),
                              "roofline_achieved_gbs": ach_local, "roofline_frac": ach_local / peak,
                              "kernel_ms": [round(x, 4) for x in k_ms]}), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    for b in prepared:
        b.batch.close()

    # ---- e2e through the C ABI with host buffers ----
    e2e_steps = max(3, min(K, args.e2e_steps))
    parity = None
    e2e = {"unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h, "steps": e2e_steps}
    # host input of one e2e call: the batch's query strings as ONE packed byte buffer (NUL-separated), the form
    # ns_engine_search_batch_packed takes — what a request-coalescing front end accumulates.  Packing is done
    # once, outside the timed region, like any other host-side input preparation.
    packed = [nsb200.Engine.pack_queries(qs) for qs in batches]
    tls = threading.local()

    def out_buffers():
        # every caller thread owns its host result buffers and reuses them, as a serving thread would
        if not hasattr(tls, "out"):
            tls.out = nsb200.Engine.result_buffers(BATCH_Q, TOPK)
        return tls.out

    if world == 1:
        def call(i):
            return eng.search_batch_packed(packed[i % nb], BATCH_Q, TOPK, out=out_buffers())
        callers = max(1, args.e2e_callers)
        single_s, e2e_s, n_multi, last = measure_e2e(call, e2e_steps, callers)
        e2e.update({"value": n_multi * BATCH_Q / e2e_s, "callers": callers, "calls": n_multi,
                    "single_caller_value": e2e_steps * BATCH_Q / single_s,
                    "path": "ns_engine_search_batch_packed(query strings): tokenise, lexicon, prepare, H2D, kernels, D2H per "
                            "call; `callers` host threads issue the calls"})
        extra["p50_ms_batch_e2e"] = 1e3 * single_s / e2e_steps
        last_idx = (n_multi - 1) % nb
        multi = eng
    else:
        # (a) one process per GPU: every rank runs the pipeline (host front end of batch i+1 under the GPUs' batch i)
        for r in searcher.search_many([(packed[i % nb], BATCH_Q) for i in range(2)], TOPK):
            pass
        dist.barrier(group=gloo)
        t1 = time.perf_counter()
        res = searcher.search_many([(packed[i % nb], BATCH_Q) for i in range(e2e_steps)], TOPK)
        torch.cuda.synchronize(dev)
        per_rank_s = time.perf_counter() - t1
        t = torch.tensor([per_rank_s], device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per_rank_s = float(t.item())
        per_rank_last = res[-1]
        e2e["per_rank_value"] = e2e_steps * BATCH_Q / per_rank_s
        e2e["per_rank_path"] = ("ShardedSearcher.search_many on every rank (one process per GPU): tokenise, lexicon for the rank's "
                                "segments, prepare, H2D, score + peer publish, merge, D2H; max over ranks")
        # (b) ONE engine handle spanning the N GPUs (ns_engine_create_multi), driven by rank 0; the other ranks wait
        dist.barrier(group=gloo)
        multi = None
        if rank == 0:
            try:
                multi = nsb200.Engine(path, devices=list(range(world)))
                assert multi.reload(), multi.last_error
            except Exception as ex:  # noqa: BLE001  (e.g. this process cannot see the other ranks' GPUs)
                log(f"[bench] single-process engine over {world} GPUs unavailable: {ex!r}")
                multi = None
        if rank == 0 and multi is None:
            e2e.update({"value": e2e["per_rank_value"], "path": e2e["per_rank_path"], "single_process_error": "engine over N GPUs unavailable"})
            last, last_idx, multi = per_rank_last, (e2e_steps - 1) % nb, eng
        elif rank == 0:

            def call(i):
                return multi.search_batch_packed(packed[i % nb], BATCH_Q, TOPK, out=out_buffers())
            # two caller counts (the box's host cores are shared by N device threads, the front-end pool and the callers);
            # the better one is the headline, both are in the line
            cap = max(1, (os.cpu_count() or 8) // 2)
            tried = {}
            best = None
            for callers in sorted({max(1, args.e2e_callers, min(world + 4, cap)), max(1, args.e2e_callers, min(2 * world + 4, cap))}):
                r = measure_e2e(call, e2e_steps, callers)
                tried[str(callers)] = r[2] * BATCH_Q / r[1]
                if best is None or tried[str(callers)] > best[0]:
                    best = (tried[str(callers)], callers, r)
            callers = best[1]
            single_s, e2e_s, n_multi, last = best[2]
            last = own_copy(last)
            e2e["by_callers"] = tried
            e2e.update({"value": n_multi * BATCH_Q / e2e_s, "callers": callers, "calls": n_multi,
                        "single_caller_value": e2e_steps * BATCH_Q / single_s,
                        "path": "ns_engine_search_batch_packed on ONE engine handle spanning the N GPUs (ns_engine_create_multi, "
                                "rank 0's process): tokenise + lexicon ONCE per batch, per-GPU prepare + H2D, score kernels "
                                "publish into GPU 0's gather buffer (peer memory), merge, one D2H; `callers` host threads"})
            extra["p50_ms_batch_e2e"] = 1e3 * single_s / e2e_steps
            last_idx = (n_multi - 1) % nb
            same = (np.array_equal(per_rank_last.nhits, last.nhits) and np.array_equal(per_rank_last.found, last.found))
            extra["per_rank_equals_single_process"] = bool(same)

    if rank == 0:
        # parity of the last e2e batch against the oracle on the same index
        parity = oracle_parity(path, batches[last_idx], last, TOPK)
        extra["parity_sample"] = parity
        # request coalescing: native caller threads issuing ONE query each (the reference's per-request engine.search)
        try:
            lat = []
            for q in batches[0][:args.single_queries]:
                t1 = time.perf_counter()
                multi.search_batch([q], TOPK)
                lat.append(time.perf_counter() - t1)
            lat.sort()
            extra["p50_ms_single_query"] = 1e3 * lat[len(lat) // 2] if lat else None
            multi.coalescer_start(max_batch=BATCH_Q, max_wait_us=args.coalesce_wait_us, dispatchers=3)
            multi.load_test(batches[1][:2048], args.coalesce_threads, 8, TOPK)  # warm
            lt = multi.load_test(batches[1], args.coalesce_threads, args.coalesce_per_thread, TOPK)
            st = multi.coalescer_stats()
            e2e.update({"coalesced_qps": lt["qps"], "coalesced_p50_ms": lt["p50_us"] / 1e3, "coalesced_p99_ms": lt["p99_us"] / 1e3,
                        "coalesced_callers": lt["threads"], "coalesced_mean_batch": st["queries"] / max(1, st["batches"]),
                        "coalesced_max_batch": st["max_batch"], "coalesce_wait_us": args.coalesce_wait_us})
            sweep = {}
            for nthr in (64, 256, 1024, 4096):   # closed loop: at most one request per caller thread in flight
                r = multi.load_test(batches[2 % nb], nthr, max(16, 65536 // nthr), TOPK)
                sweep[str(nthr)] = {"qps": r["qps"], "p50_ms": r["p50_us"] / 1e3, "p99_ms": r["p99_us"] / 1e3}
            e2e["coalesced_by_callers"] = sweep
            multi.coalescer_stop()
            # the same single-request traffic WITHOUT the coalescer (every caller launches its own Q=1 batch)
            multi.load_test(batches[2 % nb], 16, 32, TOPK)
            r = multi.load_test(batches[2 % nb], 16, 512, TOPK)
            e2e["uncoalesced_16_callers"] = {"qps": r["qps"], "p50_ms": r["p50_us"] / 1e3, "p99_ms": r["p99_us"] / 1e3,
                                             "note": "p99 includes each fresh caller thread's first CUDA call on every device"}
        except Exception as ex:  # noqa: BLE001
            e2e["coalesced_error"] = repr(ex)[:200]

    # ---- the other BASELINE workloads that fit this run ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_extras:
        with torch.cuda.stream(stream):
            try:   # configs[3]: every query holds a term with df > 10 % of the corpus, top-100
                hb = make_batches(4, head_ranks=150, seed0=101)
                r3 = device_workload(torch, eng, hb, 100, stream, 3, max(8, min(K, 40)))
                res3 = eng.search_batch(hb[0][:64], 100)
                extra["configs3_highdf_top100"] = {
                    "value": r3["value"], "unit": UNIT, "ms_per_step": r3["ms_per_step"], "step_ms": r3["step_ms"],
                    "config": workload_config(1, 1, k=100, queries="first term from the 150 most frequent (df > 10 %), 1-5 terms"),
                    "roofline": {"bound": "hbm", "achieved": r3["achieved_gbs"], "peak": peak, "unit": "GB/s",
                                 "frac": r3["achieved_gbs"] / peak, "traffic": None, "kernel": "bm25_score_topk_kernel",
                                 "algorithmic_bytes_per_launch": r3["alg_bytes"], "kernel_ms": r3["kernel_ms"]},
                    "parity_sample": oracle_parity(path, hb[0][:64], res3, 100, 64)}
                r3b = device_workload(torch, eng, hb, 10, stream, 3, max(8, min(K, 40)))
                extra["highdf_top10"] = {"value": r3b["value"], "ms_per_step": r3b["ms_per_step"],
                                         "roofline_frac": r3b["achieved_gbs"] / peak}
            except Exception as ex:  # noqa: BLE001
                extra["configs3_highdf_top100"] = {"error": repr(ex)[:200]}
            try:   # the N>1 index (8 segments) on ONE GPU: the like-for-like base of the scaling curve
                p8 = ensure_index(SEGS_SHARDED)
                e8 = nsb200.Engine(p8, device=dev)
                assert e8.reload(), e8.last_error
                r8 = device_workload(torch, e8, batches, TOPK, stream, 3, max(8, min(K, 40)))
                extra["configs2_index_on_one_gpu"] = {"value": r8["value"], "ms_per_step": r8["ms_per_step"], "step_ms": r8["step_ms"],
                                                      "roofline_frac": r8["achieved_gbs"] / peak,
                                                      "note": "1M docs in 8 segments scored by one GPU: the N=1 point of the N>1 curve"}
                e8.close()
            except Exception as ex:  # noqa: BLE001
                extra["configs2_index_on_one_gpu"] = {"error": repr(ex)[:200]}
    if rank == 0 and world == 1:
        cpu_baseline = cpu_baseline_leg(args, path, batches)

    if world > 1 and world == 8 and not args.no_extras and args.configs4:
        extra_c4 = configs4_leg(args, torch, dist, gloo, rank, world, dev, stream, peak)
        if rank == 0:
            extra["configs4_8M_docs_64_segments"] = extra_c4

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world, nseg),
            "step_ms": dist_stats(per_step),
            "exchange": searcher.mode,
            "clocks": sampler.result(),
            "e2e": e2e,
            "gpu_launches": launches_per_step * K,
            "index": {"device_bytes": rstats["device_bytes"], "posting_bytes_on_disk": rstats["posting_bytes"],
                      "device_over_disk": rstats["device_bytes"] / max(1, rstats["posting_bytes"]),
                      "reload_s": rstats["total_s"], "read_plus_upload_s": rstats["read_upload_s"], "dictionary_s": rstats["dict_s"]},
            "roofline": {"bound": "hbm", "achieved": ach_local, "peak": peak, "unit": "GB/s",
                         "frac": ach_local / peak, "traffic": (ncu_traffic() if world == 1 else None),
                         "traffic_source": f"STATIC, not measured in this run: {os.path.relpath(ncu_summary_path(), ROOT)} "
                                           "(ncu --set full, same command, one launch; see profiles/README.md for the commit)",
                         "peak_source": peak_src,
                         "kernel": "bm25_score_topk_kernel",
                         "algorithmic_bytes_per_launch": sum(alg_bytes) / max(1, len(alg_bytes)),
                         "kernel_ms": sum(k_ms[i] for i in used) / max(1, len(used)),
                         "note": "8 B x sum of LexEntry.count over (query term, segment) on rank 0, per launch; the kernel serves most "
                                 "of it from L2 (hot posting lists are shared by the batch), so this is an effective posting bandwidth"},
        }
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        line.update(extra)
        failed = parity is not None and not parity["bit_exact_vs_oracle"]
        if failed:
            # a fast wrong answer is not a result: null the headline numbers and exit non-zero
            line["value"] = None
            line["e2e"]["value"] = None
            line["error"] = "parity sample differs from the oracle"
        print(json.dumps(line), flush=True)
        if multi is not None and multi is not eng:
            multi.close()
    else:
        failed = False
    if world > 1:
        dist.barrier(group=gloo)
        dist.destroy_process_group()
    if failed:
        raise SystemExit(3)


def configs4_leg(args, torch, dist, gloo, rank, world, dev, stream, peak):
    """BASELINE configs[4]: 8M docs in 64 segments, 8 per GPU (weak-scaling view of the same path)."""
    from nextsearch_api_b200.dist import ShardedSearcher

    try:
        ndocs, nseg = 8_000_000, 64
        if rank == 0:
            free = shutil.disk_usage(os.path.dirname(BENCH_DIR) or "/dev/shm").free
            ok = free > 14e9 or os.path.exists(os.path.join(index_path(nseg, ndocs), ".complete"))
            if ok:
                ensure_index(nseg, ndocs)
        else:
            ok = True
        oks = [None] * world
        dist.all_gather_object(oks, ok, group=gloo)
        if not all(oks):
            return {"skipped": "not enough space under " + BENCH_DIR}
        path = index_path(nseg, ndocs)
        s4 = ShardedSearcher(path, dev, rank, world, max_queries=BATCH_Q, mode=args.exchange)
        assert s4.reload(), s4.engine.last_error
        batches = make_batches(4)
        with torch.cuda.stream(stream):
            prepared = [s4.prepare(qs, TOPK) for qs in batches]

            def step(i):
                s4.launch(prepared[i % 4], stream.cuda_stream)

            def sync_all():
                stream.synchronize()
                dist.barrier()
                torch.cuda.synchronize(dev)

            K4 = max(8, min(args.steps, 20))
            ms_total, per = timed_steps(torch, step, sync_all, stream, 3, K4)
        t = torch.tensor([ms_total], device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        k_ms = [b.batch.kernel_ms(0) for b in prepared]
        ach = [8.0 * b.batch.posting_count / (m * 1e-3) / 1e9 for b, m in zip(prepared, k_ms) if m > 0]
        res = s4.search_batch(batches[0][:64], TOPK)
        out = {"value": K4 * BATCH_Q / (ms_total / 1e3), "unit": UNIT, "ms_per_step": ms_total / K4, "step_ms": dist_stats(per),
               "config": workload_config(world, nseg, ndocs),
               "roofline": {"bound": "hbm", "achieved": sum(ach) / max(1, len(ach)), "peak": peak, "unit": "GB/s",
                            "frac": sum(ach) / max(1, len(ach)) / peak, "traffic": None, "kernel": "bm25_score_topk_kernel",
                            "note": "per GPU (rank 0): 1M docs in 8 segments per GPU"}}
        if rank == 0:
            out["parity_sample"] = oracle_parity(path, batches[0][:64], res, TOPK, 64)
        for b in prepared:
            b.batch.close()
        s4.close()
        return out
    except Exception as ex:  # noqa: BLE001
        return {"error": repr(ex)[:300]}


def cpu_baseline_leg(args, path, batches):
    """Times the reference (oracle/_ref, 1 thread: the engine holds one mutex for the whole search) and the oracle
    port (1 thread, all cores) on a bounded sample of the same workload."""
    from oracle import oracle as orc

    sample = unique_queries(batches, args.cpu_sample)
    oi = orc.OracleIndex(path)
    sec1, *_ = oi.search_many(sample, TOPK, nthreads=1, want_results=False)
    cores = os.cpu_count() or 1
    secN, *_ = oi.search_many(sample, TOPK, nthreads=cores, want_results=False)
    port = {"port_qps_1_thread": len(sample) / sec1, "port_qps_all_cores": len(sample) / secN, "host_cores": cores}
    if orc.have_ref():
        summ, _ = orc.ref_search(path, sample, TOPK, want_results=False)
        cb = {"value": summ["qps"], "unit": UNIT, "cores": 1, "kind": "reference",
              "sample": f"{len(sample)} unique queries of the same batches, Engine::search as shipped "
                        f"(1 thread: global mutex), p50 {summ['p50_ms']:.1f} ms"}
    else:
        cb = {"value": port["port_qps_1_thread"], "unit": UNIT, "cores": 1, "kind": "port",
              "sample": f"{len(sample)} unique queries of the same batches, oracle/bm25_oracle.c"}
    cb.update(port)
    return cb


def single_process_multi_gpu(args):
    """python bench.py --gpus N --single-process: the whole run from ONE process through ONE engine handle spanning
    N GPUs (builder-run companion of the torchrun arm; prints the same kind of line with "single_process": true)."""
    import nsb200

    n = args.gpus
    path = ensure_index(SEGS_SHARDED)
    devices = [int(x) for x in args.devices.split(",")] if args.devices else list(range(n))
    n = len(devices)
    eng = nsb200.Engine(path, devices=devices)
    assert eng.reload(), eng.last_error
    nb = args.distinct_batches
    batches = make_batches(nb)

    packed = [nsb200.Engine.pack_queries(qs) for qs in batches]

    def call(i):
        return eng.search_batch_packed(packed[i % nb], BATCH_Q, TOPK)
    steps = max(3, args.e2e_steps)
    callers = max(args.e2e_callers, min(2 * n, (os.cpu_count() or 8) // 2))
    single_s, e2e_s, n_multi, last = measure_e2e(call, steps, callers)
    line = {"metric": METRIC, "single_process": True, "n_gpus": n, "unit": UNIT,
            "e2e": {"value": n_multi * BATCH_Q / e2e_s, "single_caller_value": steps * BATCH_Q / single_s, "callers": callers,
                    "calls": n_multi},
            "parity_sample": oracle_parity(path, batches[(n_multi - 1) % nb], last, TOPK), "config": workload_config(n, SEGS_SHARDED)}
    print(json.dumps(line), flush=True)
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--distinct-batches", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=24)
    ap.add_argument("--e2e-callers", type=int, default=4, help="host threads issuing e2e search_batch calls")
    ap.add_argument("--single-queries", type=int, default=200)
    ap.add_argument("--cpu-sample", type=int, default=256)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: peer = P2P publish fused into the score kernel (CUDA IPC); nccl = all-gather + merge")
    ap.add_argument("--coalesce-threads", type=int, default=128)
    ap.add_argument("--coalesce-per-thread", type=int, default=256)
    ap.add_argument("--coalesce-wait-us", type=int, default=300)
    ap.add_argument("--no-extras", action="store_true", help="skip configs[3], the 8-segment-on-one-GPU point and configs[4]")
    ap.add_argument("--configs4", action="store_true", default=os.environ.get("NSB200_BENCH_CONFIGS4", "1") != "0",
                    help="N=8: also run BASELINE configs[4] (8M docs in 64 segments) as an extra key")
    ap.add_argument("--single-process", action="store_true", help="N GPUs from one process through one engine handle")
    ap.add_argument("--devices", default="", help="--single-process: explicit device slots, e.g. 0,1,0,1 (a GPU may repeat)")
    ap.add_argument("--profile-mode", action="store_true",
                    help="kernels only (for ncu): skip e2e, single-query latency and the CPU baseline leg")
    ap.add_argument("--ref-replicas", type=int, default=0)
    ap.add_argument("--ref-queries-per-replica", type=int, default=8)
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "ours" else max(1, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.single_process:
        single_process_multi_gpu(args)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: re-exec under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd))
    ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
