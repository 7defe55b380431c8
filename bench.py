#!/usr/bin/env python
"""bench.py — BM25 top-10 queries/sec over 1M docs (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm (oracle/_ref)

A "step" is one batch of 4096 lexicon-resolved queries (1-5 Zipf-sampled terms, k=10) scored and
top-k selected against the resident index.
  N=1 : BASELINE configs[1] — 1M docs, one segment.
  N>1 : BASELINE configs[2] — the same 1M docs split into 8 segments, segment i on rank i % N, one
        NCCL all-gather of the per-rank result blobs per batch + device merge ("strong" scaling:
        the corpus is fixed at 1M docs).
value    = queries/s with the query descriptors already resident in HBM (kernels [+ all-gather] only)
e2e      = queries/s through Engine.search_batch(query strings): tokenise, lexicon, H2D, kernels, D2H
roofline = algorithmic posting bytes (8 B x Σ LexEntry.count) / score-kernel time
"""
from __future__ import annotations

import argparse
import json
import os
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
NDOCS = int(os.environ.get("NSB200_BENCH_DOCS", "1000000"))   # default = BASELINE configs[1]/[2]; 8000000 with
SEGS_SHARDED = int(os.environ.get("NSB200_BENCH_SEGS", "8"))  # NSB200_BENCH_SEGS=64 is configs[4] (8M docs, 64 segments)
BATCH_Q = 4096
TOPK = 10
BENCH_DIR = os.environ.get("NSB200_BENCH_DIR", "/dev/shm/nsb200_bench")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def index_path(nseg: int) -> str:
    return os.path.join(BENCH_DIR, f"docs{NDOCS}_seg{nseg}")


def ensure_index(nseg: int) -> str:
    import nsb200

    path = index_path(nseg)
    marker = os.path.join(path, ".complete")
    if not os.path.exists(marker):
        t0 = time.time()
        nsb200.build_index(path, nsb200.SPEC_1M, NDOCS, nseg)
        open(marker, "w").write("ok\n")
        log(f"[bench] built {NDOCS}-doc index with {nseg} segment(s) in {time.time() - t0:.1f}s at {path}")
    return path


def make_batches(nb: int):
    import nsb200

    return [nsb200.make_queries(nsb200.SPEC_1M, BATCH_Q, 1, 5, seed=nsb200.QUERY_SEED + i) for i in range(nb)]


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML during the timed region."""

    def __init__(self, device: int, period: float = 0.01):
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


NCU_SUMMARY = os.path.join(ROOT, "profiles", "r1_v12_final_ncu.txt")


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE score-kernel launch on this workload, from the
    committed `ncu --set full` capture of the same command (profiles/); None when the file is absent."""
    try:
        tot = 0.0
        for ln in open(NCU_SUMMARY):
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
    import shutil

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
            "config": workload_config(args.gpus, nseg)}
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


def workload_config(n_gpus, nseg):
    if NDOCS != 1_000_000 or nseg not in (1, 8):
        wl = f"BASELINE configs[4]-style: {NDOCS} docs in {nseg} segments, segment-sharded + NCCL all-gather merge"
    elif nseg == 1:
        wl = "BASELINE configs[1]: 1M docs, 1 segment"
    else:
        wl = "BASELINE configs[2]: 1M docs in 8 segments, segment-sharded + NCCL all-gather merge"
    return {"workload": wl,
            "docs": NDOCS, "segments": nseg, "batch_queries": BATCH_Q, "terms_per_query": "1-5", "k": TOPK,
            "corpus": "shifted Zipf s=1 q=25, V=400000, doc_len U[100,250), seed 20260101; query seeds 7+i",
            "l2_policy": "index (1.35 GB postings) is larger than L2; distinct query batches rotate between steps",
            "parallelism": f"segments%{n_gpus}"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def ours(args, rank, world, local_rank):
    import numpy as np
    import torch

    import nsb200
    from nextsearch_api_b200.dist import ShardedSearcher

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    torch.cuda.set_device(dev)
    nseg = 1 if world == 1 else SEGS_SHARDED

    if rank == 0:
        path = ensure_index(nseg)
    if world > 1:
        dist.barrier()
    path = index_path(nseg)

    t0 = time.time()
    searcher = ShardedSearcher(path, dev, rank, world)
    assert searcher.reload(), searcher.engine.last_error
    eng = searcher.engine
    log(f"[bench] rank {rank}: reload+upload {time.time() - t0:.2f}s, device index "
        f"{eng.index.device_bytes / 1e9:.2f} GB")

    W, K = args.warmup, args.steps
    nb = max(1, min(args.distinct_batches, W + K))
    batches = make_batches(nb)
    stream = torch.cuda.Stream(device=dev)
    total_postings = []
    with torch.cuda.stream(stream):
        if world == 1:
            prepared = []
            for qs in batches:
                q_off, terms, _ = eng.resolve_batch(qs)
                prepared.append(eng.index.prepare(q_off, terms, TOPK))
            total_postings = [b.posting_count for b in prepared]

            def step(i):
                prepared[i % nb].launch(stream.cuda_stream)

            launches_per_step = prepared[0].num_launches
        else:
            prepared = [searcher.prepare(qs, TOPK) for qs in batches]
            total_postings = [b.batch.posting_count for b in prepared]

            def step(i):
                searcher.launch(prepared[i % nb])

            launches_per_step = prepared[0].batch.num_launches + 1  # + merge of the gathered blobs

        def sync_all():
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)

        for i in range(W):
            step(i)
        sync_all()
        sampler = ClockSampler(dev)
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for i in range(W, W + K):
            step(i)
        if world > 1:
            searcher.drain()  # the last batches' all-gather + merge (exchange stream) are inside the timed region
        ev1.record(stream)
        sync_all()
        sampler.stop_flag = True
        sampler.join()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    value = K * BATCH_Q / (ms_total / 1e3)

    # per-launch score-kernel time (CUDA events recorded by the library on the launching stream)
    raw = [b if world == 1 else b.batch for b in prepared]
    k_ms = [b.kernel_ms(0) for b in raw]
    used = [i for i in range(nb) if k_ms[i] > 0]
    alg_bytes = [8.0 * total_postings[i] for i in used]
    ach = [alg_bytes[j] / (k_ms[i] * 1e-3) / 1e9 for j, i in enumerate(used)]
    ach_local = sum(ach) / max(1, len(ach))
    peak, peak_src = measured_peak()

    if args.profile_mode:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
                              "warmup": W, "ms_per_step": ms_total / K, "profile_mode": True,
                              "roofline_achieved_gbs": ach_local, "kernel_ms": [round(x, 4) for x in k_ms]}), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- e2e through the public API with host buffers ----
    e2e_steps = max(3, min(K, args.e2e_steps))
    if world == 1:
        def e2e_step(i):
            return eng.search_batch(batches[i % nb], TOPK)
    else:
        def e2e_step(i):
            return searcher.search_batch(batches[i % nb], TOPK)
    def run_e2e(callers):
        """e2e_steps search_batch calls issued by `callers` host threads (ns_engine_search_batch is
        thread-safe: while one call's kernels run, another call's tokenise/lexicon/prepare proceeds).
        Every call copies its descriptors H2D and its results D2H.  Returns (seconds, last result)."""
        last_box = [None]
        if world > 1 and callers > 1:
            # sharded: one caller per rank (the collectives must be issued in the same order on every rank);
            # search_many prepares batch i+1 on the host while the GPUs work on batch i
            t0 = time.perf_counter()
            res = searcher.search_many([batches[i % nb] for i in range(e2e_steps)], TOPK)
            torch.cuda.synchronize(dev)
            return time.perf_counter() - t0, res[-1]
        if callers <= 1 or world > 1:
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                last_box[0] = e2e_step(i)
            torch.cuda.synchronize(dev)
            return time.perf_counter() - t0, last_box[0]
        nxt = [0]
        lock = threading.Lock()

        def worker():
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= e2e_steps:
                    return
                r = e2e_step(i)
                if i == e2e_steps - 1:
                    last_box[0] = r

        ths = [threading.Thread(target=worker) for _ in range(callers)]
        t0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        torch.cuda.synchronize(dev)
        return time.perf_counter() - t0, last_box[0]

    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e2e_single_s, last = run_e2e(1)
    callers = 2 if world > 1 else max(1, args.e2e_callers)  # world > 1: "2" = search_many's two batches in flight
    if callers > 1:
        run_e2e(callers)  # warm the extra callers' pooled buffers
        e2e_s, last = run_e2e(callers)
    else:
        e2e_s = e2e_single_s
    if world > 1:
        t = torch.tensor([e2e_s], device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = e2e_steps * BATCH_Q / e2e_s
    q_off0, terms0, _ = eng.resolve_batch(batches[0])
    h2d = (BATCH_Q + 1) * 4 + len(terms0) * 16 + BATCH_Q * 4
    d2h = BATCH_Q * TOPK * 12 + BATCH_Q * 4 + BATCH_Q * 8

    extra = {}
    cpu_baseline = None
    if rank == 0 and world == 1:
        # single-query latency (Q=1 goes through the split path)
        lat = []
        for q in batches[0][:args.single_queries]:
            t1 = time.perf_counter()
            eng.search_batch([q], TOPK)
            lat.append(time.perf_counter() - t1)
        lat.sort()
        extra["p50_ms_single_query"] = 1e3 * lat[len(lat) // 2] if lat else None
        extra["p50_ms_batch_e2e"] = 1e3 * e2e_single_s / e2e_steps
        cpu_baseline, parity = cpu_baseline_leg(args, path, batches, last)
        extra["parity_sample"] = parity
    elif rank == 0:
        # sharded run: the merged answer of the last e2e batch against the oracle on the same 8-segment index
        extra["parity_sample"] = parity_check(path, batches, last, (e2e_steps - 1) % nb)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(world, nseg),
            "clocks": sampler.result(),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "callers": callers,
                    "single_caller_value": e2e_steps * BATCH_Q / e2e_single_s,
                    "path": ("Engine.search_batch(query strings) -> ns_engine_search_batch_packed: tokenise, "
                             "lexicon, prepare, H2D, kernels, D2H per call; `callers` host threads issue the calls"
                             if world == 1 else
                             "ShardedSearcher.search_many(query strings): per batch tokenise, lexicon, prepare, H2D, "
                             "kernels, all-gather, merge, D2H on every rank; the host work of batch i+1 overlaps the "
                             "GPUs on batch i (`callers` = batches in flight per rank)")},
            "gpu_launches": launches_per_step * K,
            "roofline": {"bound": "hbm", "achieved": ach_local, "peak": peak, "unit": "GB/s",
                         "frac": ach_local / peak, "traffic": (ncu_traffic() if world == 1 else None),
                         "traffic_source": "profiles/r1_v12_final_ncu.txt (ncu --set full, same command, one launch)",
                         "peak_source": peak_src,
                         "kernel": "bm25_score_topk_kernel",
                         "algorithmic_bytes_per_launch": sum(alg_bytes) / max(1, len(alg_bytes)),
                         "kernel_ms": sum(k_ms[i] for i in used) / max(1, len(used)),
                         "note": "8 B x sum of LexEntry.count over (query term, segment) on rank 0, per launch"},
        }
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def parity_check(path, batches, last_result, b_idx, nchk=256):
    import numpy as np

    from oracle import oracle as orc

    oi = orc.OracleIndex(path)
    nchk = min(nchk, len(batches[b_idx]))
    qs = batches[b_idx][:nchk]
    _, s, g, d, nh, fo, hf = oi.search_many(qs, TOPK, nthreads=os.cpu_count() or 1)
    ok = bool(np.array_equal(last_result.nhits[:nchk], nh) and np.array_equal(last_result.found[:nchk], fo))
    if not ok:
        bad = [q for q in range(nchk) if last_result.nhits[q] != nh[q] or last_result.found[q] != fo[q]][:3]
        log(f"[bench] parity: nhits/found differ at {bad}: got {[(int(last_result.nhits[q]), int(last_result.found[q])) for q in bad]} "
            f"want {[(int(nh[q]), int(fo[q])) for q in bad]} queries {[qs[q] for q in bad]}")
    for q in range(nchk):
        n = int(nh[q])
        same = (np.array_equal(last_result.hits["score"][q, :n].view(np.uint32), s[q, :n].view(np.uint32))
                and np.array_equal(last_result.hits["doc"][q, :n], d[q, :n])
                and np.array_equal(last_result.hits["seg"][q, :n], g[q, :n]))
        if not same and ok:
            log(f"[bench] parity: query {q} {qs[q]!r}: got {last_result.hits[q, :n].tolist()} want "
                f"{list(zip(s[q, :n].tolist(), g[q, :n].tolist(), d[q, :n].tolist()))}")
        ok = ok and same
    return {"queries": nchk, "bit_exact_vs_oracle": bool(ok)}


def cpu_baseline_leg(args, path, batches, last_result):
    """Times the reference (oracle/_ref, 1 thread: the engine holds one mutex for the whole search)
    on a bounded sample of the same workload, and checks the GPU's answers for that sample against
    the oracle port."""
    import numpy as np

    from oracle import oracle as orc

    sample = unique_queries(batches, args.cpu_sample)
    oi = orc.OracleIndex(path)
    # parity of the last e2e batch's first queries against the oracle
    nchk = min(256, len(batches[0]))
    b_idx = (max(3, min(args.steps, args.e2e_steps)) - 1) % len(batches)
    qs = batches[b_idx][:nchk]
    _, s, g, d, nh, fo, hf = oi.search_many(qs, TOPK, nthreads=os.cpu_count() or 1)
    ok = bool(np.array_equal(last_result.nhits[:nchk], nh) and np.array_equal(last_result.found[:nchk], fo))
    for q in range(nchk):
        n = int(nh[q])
        ok = ok and np.array_equal(last_result.hits["score"][q, :n].view(np.uint32), s[q, :n].view(np.uint32))
        ok = ok and np.array_equal(last_result.hits["doc"][q, :n], d[q, :n])
    parity = {"queries": nchk, "bit_exact_vs_oracle": ok}
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
    return cb, parity


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--distinct-batches", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--e2e-callers", type=int, default=3, help="host threads issuing e2e search_batch calls (N=1)")
    ap.add_argument("--single-queries", type=int, default=200)
    ap.add_argument("--cpu-sample", type=int, default=256)
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
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: re-exec under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd))
    ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
