#!/usr/bin/env python
"""bench.py — the driver's measurement contract for atlas_b200.

A "step" is ONE `search_knn` of the global query batch (256 queries, top-40) over the sharded
768-d fp16 passage bank: BASELINE.json configs[1] at N=1 (4 Mi x 768 bank on one B200), the same
4 Mi-row shard PER GPU at N>1 (weak scaling in bank size; BASELINE.json configs[2] at N=8 = 32 Mi).

  value     queries/s, inputs resident in HBM (device-resident queries, device results), whole job
  e2e       queries/s through the reference-facing call with HOST buffers: pinned fp32 queries are
            copied H2D, searched, and (score, id) results copied D2H inside the timed region
  roofline  HBM: algorithmic bytes of the bank sweep (n_local x 768 x 2) / CUDA-event duration of the
            scan kernel (events recorded inside the library on the launching stream)
  cpu_baseline  the reference's CPU path (torch-CPU matmul + topk restated in oracle/ref_cpu_path.py)
            timed on this box's host cores on a bounded sample

`--impl reference` times that same CPU path as the reference arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LOCAL = 4 * 1024 * 1024   # passages per GPU  (BASELINE.json configs[1])
DIM = 768
NQ = 256                    # global query batch
TOPK = 40
CPU_SAMPLE_ROWS = 1 << 20   # bounded CPU sample: 1 Mi of the 4 Mi rows (scaled linearly, stated)
METRIC = "retrieve queries/sec (exact top-40 MIPS search_knn over the 768-d fp16 passage bank)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=N_LOCAL, help="passages per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md):
    one `nvidia-smi -lms 20` process is started before the region and stopped after it."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.15)  # let the first samples land before the region starts
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[j] for s in self.samples for j in range(4) if s[2 + j].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.samples)}


def make_bank(rows, device, seed):
    """Unit-scale Gaussian passages (norm ~1, Contriever-like), generated on the device in chunks."""
    import torch

    gen = torch.Generator(device=device).manual_seed(seed)
    bank = torch.empty(rows, DIM, dtype=torch.float16, device=device)
    step = 1 << 18
    for s in range(0, rows, step):
        e = min(rows, s + step)
        bank[s:e] = (torch.randn(e - s, DIM, device=device, generator=gen) / (DIM ** 0.5)).half()
    return bank


def make_queries():
    import torch

    return torch.randn(NQ, DIM, generator=torch.Generator().manual_seed(4321))


def cpu_reference_leg(steps, warmup, rows_full):
    """The reference's CPU path on a bounded sample of the workload (all host threads)."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_cpu_path

    rows = min(CPU_SAMPLE_ROWS, rows_full)
    gen = torch.Generator().manual_seed(1234)
    emb = torch.empty(DIM, rows, dtype=torch.float16)
    step = 1 << 16
    for s in range(0, rows, step):
        e = min(rows, s + step)
        emb[:, s:e] = (torch.randn(DIM, e - s, generator=gen) / (DIM ** 0.5)).half()
    q = make_queries()
    doc_map = ref_cpu_path.LazyDocMap(rows)
    for _ in range(warmup):
        ref_cpu_path.reference_search_cpu(emb, doc_map, q, TOPK)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref_cpu_path.reference_search_cpu(emb, doc_map, q, TOPK)
    dt = (time.perf_counter() - t0) / steps
    scale = rows_full / rows  # the scan is linear in the number of passages
    qps = NQ / (dt * scale)
    return {"value": qps, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{NQ} queries x {rows} of {rows_full} passages per step (time scaled x{scale:g}), "
                      f"torch-CPU fp16 matmul + topk + python doc lookup (oracle/ref_cpu_path.py), {steps} steps",
            "ms_per_step_sample": dt * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warmup = max(1, min(args.warmup, 2))
    leg = cpu_reference_leg(steps, warmup, args.rows * args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": leg["value"], "unit": "queries/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": leg["ms_per_step_sample"] * (args.rows * args.gpus / min(CPU_SAMPLE_ROWS, args.rows * args.gpus)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": leg["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"BASELINE configs[1]: {args.rows}x768 fp16 passage bank per GPU, batch={NQ} queries, top-{TOPK} exact MIPS",
            "bank_rows_per_gpu": args.rows, "bank_rows_total": args.rows * args.gpus, "queries_per_step": NQ,
            "topk": TOPK, "parallelism": f"bank sharded over {args.gpus} GPU(s), queries all-gathered",
            "l2": "inputs (6.4 GB bank per GPU) larger than the 126 MB L2; no explicit flush"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from atlas_b200 import ops
    from atlas_b200._lib import lib
    from atlas_b200.index import DistributedIndex
    import ctypes

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_cpu_path

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl")
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    L = lib()
    index = DistributedIndex()
    index._bank = make_bank(args.rows, dev, 1234 + rank)
    index.doc_map = ref_cpu_path.LazyDocMap(args.rows, rank, world)
    index._id_base, index._id_stride = rank, world

    class _SyntheticStore:  # passage text by global id, generated on the fly (no 32M python dicts)
        def lookup(self, owners_locals):
            return [{"id": str(l * world + r), "title": f"t{l * world + r}", "text": f"passage {l * world + r}"}
                    for r, l in owners_locals]

        def close(self):
            pass

    index._store = _SyntheticStore()

    q_host = make_queries().pin_memory()
    per = NQ // world
    q_host_local = q_host[rank * per:(rank + 1) * per].contiguous().pin_memory() if world > 1 else q_host
    q_dev_local = q_host_local.to(dev)

    def barrier_sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- value: device-resident inputs -----------------------------------------
    for _ in range(args.warmup):
        index.search_device(q_dev_local, TOPK)
    launches0 = L.atlas_b200_launch_count()
    L.atlas_b200_profile_enable(1)
    barrier_sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(args.steps):
            s, i = index.search_device(q_dev_local, TOPK)
        e1.record()
        barrier_sync()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = L.atlas_b200_launch_count() - launches0
    kms, kn = ctypes.c_double(0), ctypes.c_int32(0)
    L.atlas_b200_profile_collect(ctypes.byref(kms), ctypes.byref(kn))
    L.atlas_b200_profile_enable(0)
    ms_per_step = total_ms / args.steps
    value = NQ / (ms_per_step * 1e-3)

    # ---------------- e2e: host buffers through the public call ------------------------------
    def e2e_step():
        if world == 1:
            return ops.search_host(index._bank, q_host_local, TOPK, workspace=index._workspace)
        return index.search_knn(q_host_local.to(dev, non_blocking=True), TOPK)

    for _ in range(max(3, args.warmup // 2)):
        e2e_step()
    barrier_sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = e2e_step()
    barrier_sync()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    h2d = q_host_local.numel() * 4 * world
    d2h = NQ * TOPK * (2 + 8) if world > 1 else NQ * TOPK * (4 + 8)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    # the bank sweep is bracketed per launch inside the library (CUDA events on the launching stream);
    # one search sweeps the bank exactly once, split over `sweeps_per_search` launches of the same kernel
    sweeps_per_search = max(1, kn.value // args.steps)
    kernel_ms = kms.value / args.steps                       # all sweep launches of one search
    alg_bytes = args.rows * DIM * 2
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": workload_config(args),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": None, "peak_source": peak_src,
                     "kernel": "mips_scan_ts_kernel (bank sweep; one search = %d launches covering the bank once)"
                               % sweeps_per_search,
                     "kernel_ms_per_search": kernel_ms, "kernel_launches_timed": kn.value,
                     "algorithmic_bytes_per_search": alg_bytes,
                     "kernel_share_of_step": kernel_ms / ms_per_step if ms_per_step else None},
        "e2e": {"value": NQ / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                "call": "atlas_b200_search_host (C ABI, host buffers)" if world == 1 else
                        "DistributedIndex.search_knn (pinned host queries -> passage dicts + scores)"},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "aggregate_bank_GBps": world * alg_bytes / (ms_per_step * 1e-3) / 1e9,
    }
    if not args.no_cpu_baseline and world == 1:
        leg = cpu_reference_leg(3, 1, args.rows)
        line["cpu_baseline"] = {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
