#!/usr/bin/env python
"""bench.py — the driver's measurement contract for atlas_b200.

Metric (BASELINE.json): end-to-end queries/sec of the retrieve-then-read step — Contriever query embedding,
exact top-40 search over the GPU-resident 768-d fp16 passage bank (BASELINE configs[1]: 4 Mi passages per GPU,
configs[2] at 8 GPUs = 32 Mi), FiD-base forward over the 40 retrieved passages (configs[3]: T5-v1.1-base, n_docs 40,
text_maxlength 384, 32 target tokens) — plus the MIPS scan against the HBM roofline.

  value         queries/s of the whole job, step inputs already resident in HBM, CUDA events, max over ranks
  e2e           the same step with HOST inputs: pinned token ids are copied H2D and the retrieved ids / scores and the
                loss are copied D2H inside the timed region
  roofline      the dominant kernel of the step (tcgen05 GEMM): FLOPs / CUDA-event time, against the measured bf16 peak
  mips          the retrieval kernel alone at its BASELINE batch (256 queries): queries/s, C-ABI e2e with host buffers,
                and the bank sweep against the measured HBM peak (`mips.roofline`)
  train         supplementary: FiD-base forward + BACKWARD step (the training path's kernels), reader tokens/s
  refresh       supplementary: index refresh in place (Contriever-base passage embedding into bank rows), passages/s
  cpu_baseline  the reference's CPU path (oracle/: torch-CPU restatements pinned to the reference's goldens) on this
                box's host cores, bounded sample
`--impl reference` times that CPU path as the reference arm.  One process per GPU; weak scaling (per-GPU batch and
per-GPU bank shard fixed).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LOCAL = 4 * 1024 * 1024   # passages per GPU  (BASELINE.json configs[1])
DIM = 768
NQ = 256                    # global query batch
TOPK = 40
CPU_SAMPLE_ROWS = 1 << 20   # bounded CPU sample: 1 Mi of the 4 Mi rows (scaled linearly, stated)
METRIC = "end-to-end queries/sec (retrieve top-40 + FiD-base fwd)"
N_DOCS, TEXT_LEN, TARGET_LEN, QUERY_TOKENS = 40, 384, 32, 20   # BASELINE configs[3] / finetune_qa defaults
FID_FLOPS_PER_QUERY = 3.29e12


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="queries per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=N_LOCAL, help="passages per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-step", action="store_true",
                    help="run warm-up + the timed steps inside an NVTX range 'atlas_b200_timed' and exit (for ncu "
                         "--nvtx --nvtx-include 'atlas_b200_timed/'; no JSON line is printed)")
    return ap.parse_args()


def peaks(kind="hbm"):
    """(peak, source).  hbm: GB/s (burst copy figure: the sweep is timed alone); tensor: dense bf16 TFLOP/s,
    the SUSTAINED figure (the GEMMs are timed inside a long step)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    key = "hbm_gbs" if kind == "hbm" else "bf16_tflops_sustained"
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        if key in d:
            return float(d[key]), f"measured (MEASURED_PEAKS.json {key})"
    return (6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)") if kind == "hbm" else \
        (1500.0, "fallback (B200_PROFILING.md dense bf16)")


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


def cpu_reference_leg(batch, rows_full, fid_queries=1, search_steps=2):
    """The reference's CPU path on a bounded sample of the step (all host threads): Contriever query embedding and
    FiD-base forward (oracle/fid_cpu.py, fp32 torch-CPU) for `fid_queries` queries, and matmul + topk + doc lookup
    (oracle/ref_cpu_path.py) for `batch` queries over a 1 Mi-row sample of the bank, scaled linearly to the bank."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fid_cpu
    import ref_cpu_path

    torch.manual_seed(0)
    rows = min(CPU_SAMPLE_ROWS, rows_full)
    gen = torch.Generator().manual_seed(1234)
    emb = torch.empty(DIM, rows, dtype=torch.float16)
    step = 1 << 16
    for s in range(0, rows, step):
        e = min(rows, s + step)
        emb[:, s:e] = (torch.randn(DIM, e - s, generator=gen) / (DIM ** 0.5)).half()
    doc_map = ref_cpu_path.LazyDocMap(rows)
    bert = fid_cpu.bert_random_state(fid_cpu.BERT_BASE)
    t5 = fid_cpu.t5_random_state(fid_cpu.T5_BASE)
    q_ids, q_mask, dec, labels = make_step_inputs(batch, 0)
    with torch.no_grad():
        t0 = time.perf_counter()
        q = fid_cpu.contriever_forward(bert, fid_cpu.BERT_BASE, q_ids, q_mask)
        t_embed = (time.perf_counter() - t0) / batch
        ref_cpu_path.reference_search_cpu(emb, doc_map, q, TOPK)
        t0 = time.perf_counter()
        for _ in range(search_steps):
            docs, _ = ref_cpu_path.reference_search_cpu(emb, doc_map, q, TOPK)
        t_search = (time.perf_counter() - t0) / search_steps * (rows_full / rows) / batch
        ids = torch.randint(2, 32000, (fid_queries, N_DOCS * TEXT_LEN))
        mask = torch.ones(fid_queries, N_DOCS * TEXT_LEN, dtype=torch.bool)
        t0 = time.perf_counter()
        fid_cpu.fid_forward(t5, fid_cpu.T5_BASE, ids, mask, dec[:fid_queries], labels[:fid_queries], n_context=N_DOCS)
        t_read = (time.perf_counter() - t0) / fid_queries
    per_query = t_embed + t_search + t_read
    return {"value": 1.0 / per_query, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"per query: Contriever embed {t_embed:.3f} s ({batch} queries timed) + search {t_search:.3f} s "
                      f"({batch} queries x {rows} of {rows_full} passages, time scaled x{rows_full / rows:g}) + FiD-base "
                      f"fp32 forward {t_read:.3f} s ({fid_queries} query timed); torch-CPU, oracle/fid_cpu.py + "
                      f"oracle/ref_cpu_path.py",
            "s_per_query": per_query}


def make_step_inputs(batch, seed):
    """Synthetic NQ-shaped batch: queries of ~20 tokens padded to text_maxlength like the reference does
    (src/atlas.py:187-199), 32 target tokens."""
    import torch

    g = torch.Generator().manual_seed(777 + seed)
    q_ids = torch.zeros(batch, TEXT_LEN, dtype=torch.long)
    q_mask = torch.zeros(batch, TEXT_LEN, dtype=torch.long)
    q_ids[:, :QUERY_TOKENS] = torch.randint(1000, 30000, (batch, QUERY_TOKENS), generator=g)
    q_mask[:, :QUERY_TOKENS] = 1
    labels = torch.randint(2, 32000, (batch, TARGET_LEN), generator=g)
    dec = torch.cat([torch.zeros(batch, 1, dtype=torch.long), labels[:, :-1]], dim=1)
    return q_ids, q_mask, dec, labels


def workload_config(args):
    return {"workload": f"BASELINE configs[1]+[3]: per GPU a {args.rows}x768 fp16 passage bank (exact top-{TOPK}) and "
                        f"{args.batch} queries/step through Contriever-base query embedding -> search_knn -> FiD-base "
                        f"(T5-v1.1-base, n_docs {N_DOCS}, text_maxlength {TEXT_LEN}, {TARGET_LEN} target tokens) forward + loss",
            "bank_rows_per_gpu": args.rows, "bank_rows_total": args.rows * args.gpus,
            "queries_per_step": args.batch * args.gpus, "per_gpu_batch": args.batch, "topk": TOPK,
            "parallelism": f"dp{args.gpus}: bank sharded over {args.gpus} GPU(s) (queries all-gathered, one all-gather "
                           f"of per-shard top-k), reader data-parallel",
            "reader_tokens": "synthetic: token ids derived on the device from the retrieved passage ids (stand-in for a "
                             "device-resident token cache; no tokenizer vocabulary offline)",
            "weights": "random init (Contriever-base / T5-v1.1-base shapes), bf16 reader + retriever, fp16 bank",
            "l2": "bank (6.4 GB) and per-step activations (> 1 GB) exceed the 126 MB L2; no explicit flush"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    leg = cpu_reference_leg(args.batch, args.rows * args.gpus, fid_queries=1, search_steps=max(1, min(args.steps, 2)))
    line = {
        "impl": "reference", "metric": METRIC, "value": leg["value"], "unit": "queries/s", "n_gpus": args.gpus,
        "steps": 1, "warmup": 0, "ms_per_step": leg["s_per_query"] * 1e3 * args.batch * args.gpus,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": leg["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args):
    import ctypes

    import torch
    import torch.distributed as dist

    from atlas_b200 import ops
    from atlas_b200._lib import lib
    from atlas_b200.fid import FiD, T5ConfigLite
    from atlas_b200.index import DistributedIndex
    from atlas_b200.retrievers import BertConfigLite, Contriever

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
    torch.manual_seed(0)          # identical weights on every rank
    retriever = Contriever(BertConfigLite()).to(torch.bfloat16).to(dev).eval()
    reader = FiD(T5ConfigLite()).to(torch.bfloat16).to(dev).eval()
    reader.encoder.config.n_context, reader.encoder.config.bsz = N_DOCS, args.batch
    B = args.batch

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

    host = [t.pin_memory() for t in make_step_inputs(B, rank)]
    resident = [t.to(dev) for t in host]
    pos = torch.arange(TEXT_LEN, device=dev, dtype=torch.long)
    reader_mask = torch.ones(B, N_DOCS * TEXT_LEN, dtype=torch.bool, device=dev)

    def step(from_host):
        """One retrieve-then-read step of this rank's B queries (collective inside search_device)."""
        q_ids, q_mask, dec, labels = [t.to(dev, non_blocking=True) for t in host] if from_host else resident
        with torch.no_grad():
            q_emb = retriever(input_ids=q_ids, attention_mask=q_mask)                  # [B, 768]
            scores, gids = index.search_device(q_emb, TOPK)                            # [B, 40] fp16 / int64 global ids
            # reader tokens of (query, passage) pairs: synthetic ids keyed by the retrieved passage id
            reader_ids = ((gids[:, :, None] * 1315423911 + pos * 2654435761) % 32000 + 2).view(B, N_DOCS * TEXT_LEN)
            out = reader(input_ids=reader_ids, attention_mask=reader_mask, decoder_input_ids=dec, labels=labels)
        if from_host:
            return out[0].float().cpu(), gids.cpu(), scores.cpu()                     # D2H: loss + retrieved ids/scores
        return out[0], gids, scores

    # ---------------- value: device-resident inputs -----------------------------------------
    for _ in range(max(args.warmup, 3)):
        step(False)
    launches0 = L.atlas_b200_launch_count()
    barrier_sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        torch.cuda.nvtx.range_push("atlas_b200_timed")
        e0.record()
        for _ in range(args.steps):
            loss, gids, _ = step(False)
        e1.record()
        barrier_sync()
        torch.cuda.nvtx.range_pop()
    if args.profile_step:
        if world > 1:
            dist.destroy_process_group()
        return
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    launches_eager = None
    ms_per_step = total_ms / args.steps
    value = B * world / (ms_per_step * 1e-3)
    assert bool(torch.isfinite(loss.float())), "non-finite loss in the benchmark step"

    # ---------------- e2e: host inputs / host results ------------------------------------------
    for _ in range(2):
        step(True)
    barrier_sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(True)
    barrier_sync()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    h2d = sum(t.numel() * t.element_size() for t in host) * world
    d2h = (4 + B * TOPK * (8 + 2)) * world

    # ---------------- per-kernel time of the step (eager launches bracketed with CUDA events in the library) -------
    reader.cuda_graphs = False
    prof = {}
    for kind, name in ((2, "gemm"), (3, "attention")):
        step(False)
        torch.cuda.synchronize()
        l0 = L.atlas_b200_launch_count()
        L.atlas_b200_profile_enable(kind)
        step(False)
        torch.cuda.synchronize()
        work = L.atlas_b200_profile_work()
        kms, kn = ctypes.c_double(0), ctypes.c_int32(0)
        L.atlas_b200_profile_collect(ctypes.byref(kms), ctypes.byref(kn))
        L.atlas_b200_profile_enable(0)
        prof[name] = (kms.value, kn.value, work)
        launches_eager = L.atlas_b200_launch_count() - l0
    reader.cuda_graphs = True

    # ---------------- the retrieval kernel alone at its BASELINE batch (256 queries) -------------
    mips = mips_leg(args, index, dev, world, rank, L, barrier_sync, max_over_ranks)

    # ---------------- the reader's TRAINING step (forward + backward kernels), BASELINE configs[3] shapes -------------
    try:
        train = train_leg(args, reader, dev, world, L, barrier_sync, max_over_ranks)
    except Exception as e:   # the headline line must survive a failure of this supplementary leg
        train = {"error": repr(e)[:300]}

    # ---------------- index refresh in place (BASELINE configs[2]: re-embed the local shard), one embedder batch -------
    try:
        refresh = refresh_leg(args, retriever, index, dev, world, L, barrier_sync, max_over_ranks)
    except Exception as e:
        refresh = {"error": repr(e)[:300]}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = peaks("tensor")
    gemm_traffic, gemm_traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if os.path.exists(tpath):     # dram bytes of ONE representative launch from the committed `ncu --set full` capture
        with open(tpath) as f:
            tj = json.load(f)
        gemm_traffic = tj.get("dram_bytes_per_launch")
        gemm_traffic_note = f"{tj.get('launch')}: algorithmic {tj.get('algorithmic_bytes_per_launch')} B; {tj.get('source')}"
    g_ms, g_n, g_flops = prof["gemm"]
    a_ms, a_n, a_flops = prof["attention"]
    achieved = g_flops / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak if peak else None, "traffic": gemm_traffic, "traffic_note": gemm_traffic_note,
                     "peak_source": peak_src,
                     "kernel": "gemm_kernel (tcgen05 linear layers of FiD-base / Contriever-base, fused epilogues)",
                     "kernel_ms_per_step": g_ms, "kernel_launches_per_step": g_n, "algorithmic_flops_per_step": g_flops,
                     "kernel_share_of_step": g_ms / ms_per_step if ms_per_step else None,
                     "attention_kernel": {"ms_per_step": a_ms, "launches_per_step": a_n,
                                          "achieved_tflops": a_flops / (a_ms * 1e-3) / 1e12 if a_ms > 0 else None,
                                          "share_of_step": a_ms / ms_per_step if ms_per_step else None},
                     "model_flops_utilisation": FID_FLOPS_PER_QUERY * B / (ms_per_step * 1e-3) / 1e12 / peak},
        "e2e": {"value": B * world / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                "call": "Contriever.forward -> DistributedIndex.search_device -> FiD.forward (pinned host token ids in, "
                        "loss + retrieved ids / scores out)"},
        "gpu_launches": int(launches_eager) * args.steps if launches_eager else 0,
        "gpu_launches_note": "kernels per step counted on an eager step; the timed steps replay the reader's launches "
                             "from a CUDA graph",
        "clocks": clocks.summary(),
        "mips": mips,
        "train": train,
        "refresh": refresh,
    }
    if not args.no_cpu_baseline and world == 1:
        leg = cpu_reference_leg(B, args.rows)
        line["cpu_baseline"] = {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def refresh_leg(args, retriever, index, dev, world, L, barrier_sync, max_over_ranks, steps=5, warmup=2):
    """Index refresh (`Atlas.build_index`, src/atlas.py:61-88): one embedder batch of 512 synthetic passages (lengths
    U[64, 192] tokens, padded to the longest like the reference's tokenizer call) through Contriever-base with fp16
    weight copies, pooled rows written straight into bank rows (`Contriever.embed_into`).  Passages/s over all ranks (no
    communication: every rank rewrites its own shard) and the tensor-roofline fraction with SURVEY.md §8(d)'s FLOPs per
    token (169.9 MFLOP + 36 864 L).  Runs last: it overwrites the first 512 rows of the synthetic bank."""
    import torch

    nb, lmax = 512, 192
    g = torch.Generator().manual_seed(4242)
    lens = torch.randint(64, lmax + 1, (nb,), generator=g)
    lens[0] = lmax
    ids = torch.randint(1000, 30000, (nb, lmax), generator=g)
    mask = (torch.arange(lmax)[None, :] < lens[:, None]).to(torch.int64)
    ids = (ids * mask).to(dev)
    mask = mask.to(dev)
    rows = index._bank[:nb]

    def step():
        retriever.embed_into(ids, mask, rows, dtype=torch.float16)

    for _ in range(warmup):
        step()
    barrier_sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier_sync()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    assert bool(torch.isfinite(rows.float()).all()), "non-finite embeddings in the refresh leg"
    flops = nb * lmax * (169.9e6 + 36864.0 * lmax)
    peak, peak_src = peaks("tensor")
    achieved = flops / (ms * 1e-3) / 1e12
    return {"metric": "index refresh passages/sec (Contriever-base fp16 embed of 512-passage batches into bank rows)",
            "value": nb * world / (ms * 1e-3), "unit": "passages/s", "ms_per_batch": ms, "steps": steps,
            "passages_per_batch": nb, "padded_tokens": lmax, "tokens_per_s": nb * lmax * world / (ms * 1e-3),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None, "peak_source": peak_src,
                         "algorithmic_flops_per_batch": flops},
            "shard_refresh_estimate_s": args.rows / (nb / (ms * 1e-3))}


def train_leg(args, reader, dev, world, L, barrier_sync, max_over_ranks, steps=3, warmup=2):
    """FiD-base forward + backward (`loss.backward()` through grad_ops.py's kernels; at N > 1 followed by one NCCL
    all-reduce of the flattened gradients, what DDP does in train.py) on `train_batch` queries x 40 passages x 384 tokens
    per GPU: reader tokens/s (B * n_docs * text_maxlength per step, SURVEY.md §8d C5's unit) and the share of the GEMM /
    attention-backward kernels.  Supplementary to the headline metric (which is the forward step)."""
    import ctypes

    import torch
    import torch.distributed as dist

    tb = min(args.batch, 2)
    reader.train()
    cfg = reader.encoder.config
    old = (cfg.n_context, cfg.bsz)
    cfg.n_context, cfg.bsz = N_DOCS, tb
    g = torch.Generator().manual_seed(99)
    ids = torch.randint(2, 32000, (tb, N_DOCS * TEXT_LEN), generator=g).to(dev)
    mask = torch.ones(tb, N_DOCS * TEXT_LEN, dtype=torch.bool, device=dev)
    labels = torch.randint(2, 32000, (tb, TARGET_LEN), generator=g).to(dev)

    def step():
        reader.zero_grad(set_to_none=True)
        out = reader(input_ids=ids, attention_mask=mask, labels=labels)
        out[0].backward()
        if world > 1:
            flat = torch.cat([p.grad.reshape(-1) for p in reader.parameters() if p.grad is not None])
            dist.all_reduce(flat)
        return out[0]

    try:
        for _ in range(warmup):
            loss = step()
        barrier_sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        barrier_sync()
        ms = max_over_ranks(e0.elapsed_time(e1)) / steps
        assert bool(torch.isfinite(loss.float())), "non-finite training loss"
        shares = {}
        for kind, name in ((2, "gemm"), (4, "attention_bwd")):
            L.atlas_b200_profile_enable(kind)
            step()
            torch.cuda.synchronize()
            work = L.atlas_b200_profile_work()
            kms, kn = ctypes.c_double(0), ctypes.c_int32(0)
            L.atlas_b200_profile_collect(ctypes.byref(kms), ctypes.byref(kn))
            L.atlas_b200_profile_enable(0)
            shares[name] = {"ms_per_step": kms.value, "launches_per_step": kn.value,
                            "achieved_tflops": work / (kms.value * 1e-3) / 1e12 if kms.value > 0 else None,
                            "share_of_step": kms.value / ms if ms else None}
    finally:
        reader.zero_grad(set_to_none=True)
        reader.eval()
        cfg.n_context, cfg.bsz = old
    tokens = tb * N_DOCS * TEXT_LEN * world
    return {"metric": "reader training tokens/sec (FiD-base forward + backward, bf16, dropout 0)",
            "value": tokens / (ms * 1e-3), "unit": "tokens/s", "ms_per_step": ms, "steps": steps,
            "queries_per_step": tb * world, "reader_tokens_per_step": tokens,
            "gradient_allreduce": "one NCCL all-reduce of the flattened bf16 gradients" if world > 1 else "none (1 GPU)",
            "kernels": shares}


def mips_leg(args, index, dev, world, rank, L, barrier_sync, max_over_ranks):
    """search_knn alone: 256 queries / top-40 over the sharded bank (BASELINE configs[1] at N=1, configs[2] at N=8)."""
    import ctypes

    import torch

    from atlas_b200 import ops

    steps, warmup = max(20, args.steps), 5
    q_host = make_queries().pin_memory()
    per = NQ // world
    q_host_local = q_host[rank * per:(rank + 1) * per].contiguous().pin_memory() if world > 1 else q_host
    q_dev_local = q_host_local.to(dev)
    for _ in range(warmup):
        index.search_device(q_dev_local, TOPK)
    L.atlas_b200_profile_enable(1)
    barrier_sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        index.search_device(q_dev_local, TOPK)
    e1.record()
    barrier_sync()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    kms, kn = ctypes.c_double(0), ctypes.c_int32(0)
    L.atlas_b200_profile_collect(ctypes.byref(kms), ctypes.byref(kn))
    L.atlas_b200_profile_enable(0)
    ms_per_step = total_ms / steps

    def e2e_step():
        if world == 1:
            return ops.search_host(index._bank, q_host_local, TOPK, workspace=index._workspace)
        return index.search_knn(q_host_local.to(dev, non_blocking=True), TOPK)

    for _ in range(3):
        e2e_step()
    barrier_sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    barrier_sync()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / steps
    peak, peak_src = peaks("hbm")
    kernel_ms = kms.value / steps
    alg_bytes = args.rows * DIM * 2
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_mips_scan_traffic.json")
    if os.path.exists(tpath):     # dram bytes of the sweep from the committed `ncu --set full` capture (see profiles/)
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_search")
    return {
        "metric": "retrieve queries/sec (exact top-40 search_knn, 256-query batches)", "value": NQ / (ms_per_step * 1e-3),
        "unit": "queries/s", "ms_per_step": ms_per_step, "steps": steps, "queries_per_step": NQ,
        "e2e": {"value": NQ / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": q_host_local.numel() * 4 * world,
                "d2h_bytes_per_step": NQ * TOPK * (2 + 8) if world > 1 else NQ * TOPK * (4 + 8),
                "call": "atlas_b200_search_host (C ABI, host buffers)" if world == 1 else
                        "DistributedIndex.search_knn (pinned host queries -> passage dicts + scores)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                     "kernel": "mips_scan_ts_kernel (bank sweep; one search = %d launches covering the bank once)"
                               % max(1, kn.value // steps),
                     "kernel_ms_per_search": kernel_ms, "kernel_launches_timed": kn.value,
                     "algorithmic_bytes_per_search": alg_bytes,
                     "kernel_share_of_step": kernel_ms / ms_per_step if ms_per_step else None},
        "aggregate_bank_GBps": world * alg_bytes / (ms_per_step * 1e-3) / 1e9,
    }


_JSON_FD = None


def emit(line):
    """The ONE JSON line of the contract goes to the process's original stdout; everything else that writes to fd 1
    during the run (NCCL prints its version there when NCCL_DEBUG=VERSION) has been pointed at stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
