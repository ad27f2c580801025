"""Where does the TS scan kernel spend its time?  Per-CTA cycle counters from the MMA / producer / epilogue."""
import os, sys, ctypes
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atlas_b200 import ops
from atlas_b200._lib import lib
dev = torch.device("cuda:0")
L = lib()
n = 4 * 1024 * 1024
g = torch.Generator(device=dev).manual_seed(0)
bank = torch.empty(n, 768, device=dev, dtype=torch.float16)
for s0 in range(0, n, 1 << 18):
    bank[s0:s0 + (1 << 18)] = (torch.randn(1 << 18, 768, device=dev, generator=g) / 27.7).half()
for nq in (256, 128):
    q = torch.randn(nq, 768, device=dev, generator=g)
    for _ in range(3):
        ops.mips_topk(bank, q, 40)
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    L.atlas_b200_mips_set_debug_counters(ctypes.c_void_p(dbg.data_ptr()))
    L.atlas_b200_profile_enable(1)
    ops.mips_topk(bank, q, 40)
    torch.cuda.synchronize()
    ms, cnt = ctypes.c_double(0), ctypes.c_int32(0)
    L.atlas_b200_profile_collect(ctypes.byref(ms), ctypes.byref(cnt))
    L.atlas_b200_profile_enable(0)
    L.atlas_b200_mips_set_debug_counters(None)
    d = dbg.cpu().numpy().reshape(148, 16)
    cnts = ops._default_ws.buf[:1024].view(torch.int32).cpu().numpy()[:nq]
    print("   candidates per query after final sweep: mean", cnts.mean(), "max", cnts.max())
    lead = d[d[:, 5] > 0]
    print(f"nq={nq}: bank sweeps ({cnt.value} launches) {ms.value:.3f} ms; leader CTAs {len(lead)}; tiles/CTA {lead[:,5].mean():.0f}")
    names = ["producer wait_empty", "mma wait_tmem_empty", "mma wait_full", "mma wait_a_ready", "mma total", "tiles", "epi wait_tmem_full", "epi ld+wait", "epi syncwarp+arrive", "epi max tree", "epi hit path"]
    for j, nm in enumerate(names):
        col = (lead if j in (1, 2, 3, 4, 5) else d[d[:, j] > 0])[:, j]
        if len(col):
            print(f"   {nm:22s} mean {col.mean():12.0f}  min {col.min():12.0f}  max {col.max():12.0f}")
    tot = lead[:, 4].mean()
    print(f"   per tile: total {tot/lead[:,5].mean():.0f} cyc, wait_full {lead[:,2].mean()/lead[:,5].mean():.0f}, wait_tmem_empty {lead[:,1].mean()/lead[:,5].mean():.0f}, issue+other {(tot-lead[:,2].mean()-lead[:,1].mean()-lead[:,3].mean())/lead[:,5].mean():.0f}")
