"""One-shot check of the experimental tcgen05 dQ kernel (ATLAS_B200_ATTN_BWD_TC=1) against the validated warp-MMA path.
    python tools/try_tc_bwd.py ref   -> runs the default path, saves results to /tmp/tc_bwd_ref.pt
    ATLAS_B200_ATTN_BWD_TC=1 python tools/try_tc_bwd.py tc -> runs the tcgen05 dQ kernel, compares, times all shapes
    ATLAS_B200_ATTN_BWD_TC=2 python tools/try_tc_bwd.py tc -> also the tcgen05 dK / dV kernel.
Run under `timeout 60`: a protocol bug in an mbarrier pipeline traps through the watchdog of common.cuh."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atlas_b200 import ops  # noqa: E402

mode = sys.argv[1]
dev = torch.device("cuda:0")
CASES = [(6, 4, 384, True, True, torch.bfloat16), (3, 2, 200, True, True, torch.float16), (2, 3, 64, False, True, torch.bfloat16),
         (80, 12, 384, True, True, torch.bfloat16)]
out = {}
for ci, (B, H, L, use_bias, use_mask, dt) in enumerate(CASES):
    g = torch.Generator().manual_seed(100 + ci)
    qkv = (torch.randn(B * L, 3 * H * 64, generator=g) * 0.35).to(dt).to(dev)
    bias = (0.5 * torch.randn(H, 2 * L - 1, generator=g)).to(dev) if use_bias else None
    lens = torch.randint(max(1, L // 3), L + 1, (B,), generator=g)
    lens[0] = L
    mask = ((torch.arange(L)[None, :] >= lens[:, None]).float() * -10000.0).to(dev) if use_mask else None
    o, lse = ops.attention(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, B, H, L, L, add_mask=mask, bias_delta=bias, return_lse=True)
    do = torch.randn(B * L, H * 64, generator=g).to(dt).to(dev)

    def run():
        dqkv = torch.zeros_like(qkv)
        db = ops.attention_bwd(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, o, do, dqkv, 0, dqkv, H * 64, dqkv, 2 * H * 64, B, H, L, L,
                               add_mask=mask, bias_delta=bias, need_dbias=use_bias, lse=lse)
        return dqkv, db

    dqkv, db = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[ci] = (dqkv.float().cpu(), db.float().cpu() if db is not None else None)
    print(f"case {ci} B={B} H={H} L={L} {dt}: {ms:.3f} ms, finite={bool(torch.isfinite(dqkv.float()).all())}", flush=True)
if mode == "ref":
    torch.save(out, "/tmp/tc_bwd_ref.pt")
else:
    ref = torch.load("/tmp/tc_bwd_ref.pt")
    for ci in out:
        H = CASES[ci][1]
        a, b = out[ci][0], ref[ci][0]
        for name, sl in (("dQ", slice(0, H * 64)), ("dK", slice(H * 64, 2 * H * 64)), ("dV", slice(2 * H * 64, 3 * H * 64))):
            d = float((a[:, sl] - b[:, sl]).abs().max())
            print(f"case {ci} {name}: max |tc - ref| {d:.3e} (ref max {float(b[:, sl].abs().max()):.3e})")
        if out[ci][1] is not None:
            print(f"case {ci} dbias: max |tc - ref| {float((out[ci][1] - ref[ci][1]).abs().max()):.3e} "
                  f"(ref max {float(ref[ci][1].abs().max()):.3e})")
