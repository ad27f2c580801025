import os, sys, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from atlas_b200 import ops
dev = torch.device("cuda:0")
for dtype in (torch.float16, torch.bfloat16):
    for (M, N, K) in [(333, 768, 768), (333, 768, 3072), (4096, 768, 3072)]:
        g = torch.Generator(device="cpu").manual_seed(1)
        x = (torch.randn(M, K, generator=g) * 0.5).to(dtype).to(dev)
        w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(dtype).to(dev)
        ref64 = x.double() @ w.double().T
        ref32 = x.float() @ w.float().T
        y = ops.linear(x, w)
        yt = x @ w.T            # cuBLAS same dtype
        exact = ref64.to(dtype)  # correctly rounded
        def stats(a):
            diff = (a.double() - ref64).abs()
            ulp = (ref64.abs().clamp_min(1e-30)).log2().floor().exp2() * torch.finfo(dtype).eps
            return float((diff / ulp).max()), float((a != exact).float().mean()), float(diff.max())
        print(dtype, (M, N, K), "ours: max err %.3f ulp, frac!=exact %.4f, maxabs %.5f" % stats(y),
              "| cublas: %.3f ulp, %.4f, %.5f" % stats(yt), "| fp32 ref vs fp64 maxabs %.2e" % float((ref32.double()-ref64).abs().max()))
        bad = (y != exact)
        if bad.any():
            idx = bad.nonzero()[0]
            i, j = int(idx[0]), int(idx[1])
            print("   example", i, j, "ours", float(y[i, j]), "exact", float(exact[i, j]), "ref64", float(ref64[i, j]), "cublas", float(yt[i, j]))
