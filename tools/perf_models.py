"""Throughput of the model-level paths at BASELINE shapes (random-init weights, synthetic tokens)."""
import os, sys, time, ctypes
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atlas_b200.fid import FiD, T5ConfigLite
from atlas_b200.retrievers import Contriever, BertConfigLite
from atlas_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

# ---- raw GEMM shapes of FiD-base encoder (tokens = 15360 per query) ----
for (M, N, K, epi) in [(15360 * 2, 768, 768, 0), (15360 * 2, 2304, 768, 0), (15360 * 2, 4096, 768, 4), (15360 * 2, 768, 2048, 3), (512 * 256, 3072, 768, 2)]:
    x = torch.randn(M, K, device=dev).bfloat16() * 0.1
    w = torch.randn(N, K, device=dev).bfloat16() * 0.03
    r = torch.randn(M, N // 2 if epi == 4 else N, device=dev).bfloat16()
    b = torch.zeros(N, device=dev).bfloat16()
    ms = timeit(lambda: ops.linear(x, w, b if epi in (2,) else None, r if epi == 3 else None, epilogue=epi), n=10)
    ms_t = timeit(lambda: torch.nn.functional.linear(x, w), n=10)
    print(f"GEMM M={M} N={N} K={K} epi={epi}: {ms:.3f} ms = {2*M*N*K/ms/1e9:.0f} TFLOP/s   (torch/cuBLAS plain: {ms_t:.3f} ms = {2*M*N*K/ms_t/1e9:.0f})", flush=True)

# ---- attention at FiD encoder shape ----
S, H, L = 80, 12, 384
qkv = torch.randn(S * L, 3 * H * 64, device=dev).bfloat16() * 0.2
bias = torch.randn(H, 2 * L - 1, device=dev)
am = torch.zeros(S, L, device=dev)
ms = timeit(lambda: ops.attention(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, S, H, L, L, add_mask=am, bias_delta=bias), n=10)
fl = 4 * S * H * L * L * 64
print(f"attention S={S} H={H} L={L}: {ms:.3f} ms = {fl/ms/1e9:.0f} TFLOP/s", flush=True)

# ---- FiD-base forward ----
model = FiD(T5ConfigLite()).to(torch.bfloat16).to(dev).eval()
for B in (1, 2, 4):
    n_ctx, L, T = 40, 384, 32
    model.encoder.config.n_context, model.encoder.config.bsz = n_ctx, B
    ids = torch.randint(2, 32000, (B, n_ctx * L), device=dev)
    mask = torch.ones(B, n_ctx * L, dtype=torch.bool, device=dev)
    dec = torch.randint(2, 32000, (B, T), device=dev)
    with torch.no_grad():
        t_enc = timeit(lambda: model.encode(ids, mask), n=3, warm=1)
        enc = model.encode(ids, mask)
        t_kv = timeit(lambda: model.cross_kv(enc), n=3, warm=1)
        kv = model.cross_kv(enc)
        t_dec = timeit(lambda: model.decode(dec, enc, mask, cross_kv=kv), n=3, warm=1)
        t_all = timeit(lambda: model(input_ids=ids, attention_mask=mask, decoder_input_ids=dec), n=3, warm=1)
    print(f"FiD-base B={B}: encode {t_enc:.2f} ms, cross_kv {t_kv:.2f} ms, decode {t_dec:.2f} ms, forward {t_all:.2f} ms "
          f"-> {B/t_all*1e3:.1f} queries/s, {3.29*B/t_all:.3f} PFLOP/s", flush=True)

# ---- Contriever-base passage embedding ----
cm = Contriever(BertConfigLite()).half().to(dev).eval()
for (B, L) in [(512, 128), (512, 256)]:
    ids = torch.randint(1, 30000, (B, L), device=dev)
    mask = torch.ones(B, L, dtype=torch.long, device=dev)
    with torch.no_grad():
        ms = timeit(lambda: cm(input_ids=ids, attention_mask=mask), n=3, warm=1)
    fl = B * L * (169.9e6 + 36864 * L)
    print(f"Contriever-base B={B} L={L}: {ms:.2f} ms -> {B/ms*1e3:.0f} passages/s, {fl/ms/1e9:.0f} TFLOP/s", flush=True)
