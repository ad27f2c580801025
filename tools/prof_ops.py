"""Single-op drivers for ncu captures: python tools/prof_ops.py attention|gemm|contriever|fid [reps]."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atlas_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
what = sys.argv[1] if len(sys.argv) > 1 else "attention"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3

def timed(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

if what == "attention":
    S, H, L = 160, 12, 384      # FiD-base encoder, 4 queries x 40 passages
    qkv = torch.randn(S * L, 3 * H * 64, device=dev).bfloat16() * 0.2
    bias = torch.randn(H, 2 * L - 1, device=dev)
    am = torch.zeros(S, L, device=dev)
    fn = lambda: ops.attention(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, S, H, L, L, add_mask=am, bias_delta=bias)
    ms = timed(fn, reps)
    print(f"attention S={S} H={H} L={L}: {ms:.3f} ms = {4 * S * H * L * L * 64 / ms / 1e9:.0f} TFLOP/s")
elif what == "gemm":
    M, N, K = 61440, 4096, 768  # FiD-base wi_0|wi_1 gated projection, 4 queries
    x = torch.randn(M, K, device=dev).bfloat16() * 0.1
    w = torch.randn(N, K, device=dev).bfloat16() * 0.03
    fn = lambda: ops.linear(x, w, epilogue=ops.EPI_GATED)
    ms = timed(fn, reps)
    print(f"gemm M={M} N={N} K={K} gated: {ms:.3f} ms = {2 * M * N * K / ms / 1e9:.0f} TFLOP/s")
elif what == "gemm4":
    # the four linear-layer shapes of a FiD-base encoder block at 4 queries x 40 passages x 384 tokens, next to cuBLAS
    # (torch.matmul of the same shape, plain epilogue): ATLAS_B200_GEMM_QUAD=0 selects the 2-CTA pair kernel for A/B runs
    M = 61440
    for name, N, K, epi in (("qkv", 2304, 768, None), ("o-proj+res", 768, 768, "res"), ("wi gated", 4096, 768, "gated"),
                            ("wo+res", 768, 2048, "res")):
        x = torch.randn(M, K, device=dev).bfloat16() * 0.1
        w = torch.randn(N, K, device=dev).bfloat16() * 0.03
        r = torch.randn(M, N, device=dev).bfloat16() if epi == "res" else None
        if epi == "gated":
            fn = lambda: ops.linear(x, w, epilogue=ops.EPI_GATED)
        elif epi == "res":
            fn = lambda: ops.linear(x, w, None, residual=r, epilogue=ops.EPI_RESIDUAL)
        else:
            fn = lambda: ops.linear(x, w)
        ms = timed(fn, reps)
        wt = w.t().contiguous()
        ms_ref = timed(lambda: torch.matmul(x, wt), reps)
        fl = 2 * M * N * K
        print(f"gemm {name:12s} M={M} N={N} K={K}: {ms:.3f} ms = {fl / ms / 1e9:.0f} TFLOP/s   cuBLAS plain: {ms_ref:.3f} ms = "
              f"{fl / ms_ref / 1e9:.0f} TFLOP/s")
elif what == "mips":
    # one 256-query top-40 search over a 4 Mi x 768 fp16 bank (BASELINE configs[1]): for `ncu --metrics gpu__time_duration.sum`
    n, nq, k = 4 * 1024 * 1024, 256, 40
    bank = torch.empty((n, 768), dtype=torch.float16, device=dev)
    for a in range(0, n, 1 << 19):
        bank[a:a + (1 << 19)] = (torch.randn((1 << 19, 768), device=dev) * 0.05).half()
    q = (torch.randn(nq, 768, device=dev) * 0.05).half()
    fn = lambda: ops.mips_topk(bank, q, k)
    ms = timed(fn, reps)
    print(f"mips n={n} nq={nq} k={k}: {ms:.3f} ms = {n * 768 * 2 / ms / 1e6:.0f} GB/s over the bank")
elif what == "packed":
    # the padding-compacted FiD-base encoder ops at the bench's length distribution (8 queries x 40 passages, U[148, 276] real
    # tokens padded to 384): packed three-lane attention vs the padded kernel with block skipping vs no skipping, and the gated
    # projection over the packed rows (device-side row count) vs all rows
    S, H, L = 320, 12, 384
    lens = torch.randint(148, 277, (S,), device=dev)
    mask = (torch.arange(L, device=dev)[None, :] >= lens[:, None]).float() * -10000.0
    live = ops.key_block_live(mask)
    keep, off, src, count = ops.segment_tile_scan(live)
    qkv = torch.randn(S * L, 3 * H * 64, device=dev).bfloat16() * 0.2
    bias = torch.randn(H, 2 * L - 1, device=dev)
    dense_fl = 4 * S * H * L * L * 64
    frac = float(keep.float().mean())
    ms_p = timed(lambda: ops.attention_packed(qkv, keep, off, S, H, L, mask, bias), reps)
    ms_s = timed(lambda: ops.attention(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, S, H, L, L, add_mask=mask, bias_delta=bias,
                                       block_live=live), reps)
    ms_d = timed(lambda: ops.attention(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, S, H, L, L, add_mask=mask, bias_delta=bias), reps)
    print(f"attention S={S} H={H} L={L}, kept tiles {frac:.3f}: packed {ms_p:.3f} ms | padded, dead key blocks skipped {ms_s:.3f} ms | "
          f"padded, every block {ms_d:.3f} ms = {dense_fl / ms_d / 1e9:.0f} TFLOP/s; packed executes "
          f"{dense_fl * float((keep.float().sum(1) ** 2).mean()) / 36 / ms_p / 1e9:.0f} TFLOP/s on kept x kept tiles")
    M, N, K = S * L, 4096, 768
    x = torch.randn(M, K, device=dev).bfloat16() * 0.1
    w = torch.randn(N, K, device=dev).bfloat16() * 0.03
    ms_r = timed(lambda: ops.linear(x, w, epilogue=ops.EPI_GATED, rows=count), reps)
    ms_a = timed(lambda: ops.linear(x, w, epilogue=ops.EPI_GATED), reps)
    rows = int(count)
    print(f"gemm gated M={M} N={N} K={K}: first {rows} rows (device-side count) {ms_r:.3f} ms = {2 * rows * N * K / ms_r / 1e9:.0f} "
          f"TFLOP/s | all rows {ms_a:.3f} ms = {2 * M * N * K / ms_a / 1e9:.0f} TFLOP/s")
elif what == "refresh":
    # one index-refresh embedder batch (bench.py refresh leg): 512 passages of U[64, 192] tokens padded to 192 through
    # Contriever-base with fp16 weight copies, pooled rows written into bank rows (Contriever.embed_into)
    from atlas_b200.retrievers import BertConfigLite, Contriever
    model = Contriever(BertConfigLite()).to(torch.bfloat16).to(dev).eval()
    nb, lmax = 512, 192
    g = torch.Generator().manual_seed(4242)
    lens = torch.randint(64, lmax + 1, (nb,), generator=g)
    lens[0] = lmax
    ids = torch.randint(1000, 30000, (nb, lmax), generator=g)
    mask = (torch.arange(lmax)[None, :] < lens[:, None]).to(torch.int64)
    ids, mask = (ids * mask).to(dev), mask.to(dev)
    rows = torch.zeros((nb, 768), dtype=torch.float16, device=dev)
    fn = lambda: model.embed_into(ids, mask, rows, dtype=torch.float16)
    ms = timed(fn, reps)
    print(f"refresh batch {nb} x {lmax} (kept tiles {float(((lens + 63) // 64).sum()) / (nb * 3):.3f}): {ms:.3f} ms = "
          f"{nb / ms * 1e3:.0f} passages/s")
elif what == "xattn":
    # FiD-base decoder cross-attention of the teacher-forced forward: 8 queries x 32 target tokens against 40 x 384 encoder keys,
    # passages of U[148, 276] real tokens padded to 384 (the bench's length distribution)
    B, H, T, n, Lp = 8, 12, 32, 40, 384
    Lk = n * Lp
    q = torch.randn(B * T, H * 64, device=dev).bfloat16() * 0.3
    kv = torch.randn(B * Lk, 2 * H * 64, device=dev).bfloat16() * 0.3
    lens = torch.randint(148, 277, (B, n), device=dev)
    valid = torch.arange(Lp, device=dev)[None, None, :] < lens[..., None]
    mask = ((~valid).reshape(B, Lk).float() * -1e9)
    live = ops.key_block_live(mask)
    fn = lambda: ops.cross_attention_split(q, 0, kv, 0, H * 64, B, H, T, Lk, add_mask=mask, scale=1.0, split=384, tile_live=live)
    ms = timed(fn, reps)
    frac = float(live.float().mean()) if live is not None else 1.0
    byt = B * Lk * 2 * H * 64 * 2
    print(f"cross-attention B={B} H={H} T={T} Lk={Lk}: {ms:.3f} ms (stream + combine); K/V {byt / 1e6:.0f} MB, live tiles {frac:.2f}: "
          f"{byt * frac / ms / 1e6:.0f} GB/s on the live bytes, {byt / ms / 1e6:.0f} GB/s dense-equivalent")
elif what == "attn_bwd":
    S, H, L = 80, 12, 384       # FiD-base encoder, 2 queries x 40 passages (the bench's training leg)
    qkv = torch.randn(S * L, 3 * H * 64, device=dev).bfloat16() * 0.2
    bias = torch.randn(H, 2 * L - 1, device=dev)
    am = torch.zeros(S, L, device=dev)
    out, lse = ops.attention(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, S, H, L, L, add_mask=am, bias_delta=bias, return_lse=True)
    do = torch.randn_like(out)
    dqkv = torch.empty_like(qkv)
    fn = lambda: ops.attention_bwd(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, out, do, dqkv, 0, dqkv, H * 64, dqkv, 2 * H * 64,
                                   S, H, L, L, add_mask=am, bias_delta=bias, need_dbias=True, lse=lse)
    ms = timed(fn, reps)
    print(f"attention bwd S={S} H={H} L={L} (lse from the forward, dbias): {ms:.3f} ms = "
          f"{14 * S * H * L * L * 64 / ms / 1e9:.0f} TFLOP/s executed (7 tile GEMMs), {10 * S * H * L * L * 64 / ms / 1e9:.0f} useful")
elif what == "wgrad":
    M, N, K = 30720, 4096, 768  # dW of the interleaved wi_0|wi_1 projection, 2 queries
    x = torch.randn(M, K, device=dev).bfloat16() * 0.1
    dy = torch.randn(M, N, device=dev).bfloat16() * 0.1
    fn = lambda: ops.linear_wgrad(dy, x)
    ms = timed(fn, reps)
    print(f"wgrad tokens={M} N={N} K={K} (MN-major tcgen05): {ms:.3f} ms = {2 * M * N * K / ms / 1e9:.0f} TFLOP/s")
