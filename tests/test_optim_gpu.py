"""GPU: the multi-tensor optimiser step and gradient statistics (csrc/optim.cu, atlas_b200/optim.py) against the reference's
own `AdamWFP32Copy.step` arithmetic (src/AdamWFP32Copy.py:79-169 = torch.optim AdamW on fp32 copies + copy back) and
`compute_grad_stats` (src/util.py:200-222) restated with torch ops on the same tensors.  Tolerance: a few fp32 ulps per step
(FMA contraction / reciprocal ordering differ from ATen's foreach kernels); the bf16 parameters must round identically
except where the fp32 value sits within those ulps of a rounding boundary."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from atlas_b200._lib import lib

    lib()
    return torch.device("cuda:0")


def _reference_step(master, m, v, grads, step, lr, betas, eps, wd, scale):
    """torch.optim._functional.adamw single-tensor arithmetic on the fp32 copies (what the reference's step executes)."""
    b1, b2 = betas
    for p, ea, es, g in zip(master, m, v, grads):
        g = g.float() / scale
        p.mul_(1 - lr * wd)
        ea.lerp_(g, 1 - b1)
        es.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        denom = (es.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(ea, denom, value=-(lr / bc1))


@pytest.mark.parametrize("pdtype", [torch.bfloat16, torch.float32])
def test_adamw_fp32copy_matches_reference_arithmetic(dev, pdtype):
    from atlas_b200.optim import AdamWFP32Copy

    g = torch.Generator(device="cpu").manual_seed(4)
    shapes = [(768, 768), (5,), (1000, 33), (1,), (2048, 130), (70001,)]
    params = [torch.nn.Parameter((torch.randn(*s, generator=g) * 0.05).to(pdtype).to(dev)) for s in shapes]
    lr, betas, eps, wd = 3e-4, (0.9, 0.999), 1e-8, 0.01
    opt = AdamWFP32Copy(params, lr=lr, betas=betas, eps=eps, weight_decay=wd)
    master = [p.detach().float().clone() for p in params]
    m = [torch.zeros_like(x) for x in master]
    v = [torch.zeros_like(x) for x in master]
    for step in range(1, 6):
        scale = 1.0 if step % 2 else 4.0
        grads = [(torch.randn(*s, generator=g) * 0.1 * scale).to(pdtype).to(dev) for s in shapes]
        if step == 3:
            grads[1] = None                      # a parameter without gradient is skipped (its step count stays behind)
        for p, gr in zip(params, grads):
            p.grad = gr
        opt.step(scale=scale)
        for i, gr in enumerate(grads):
            if gr is None:
                continue
            st = int(opt.state[params[i]]["step"])
            _reference_step([master[i]], [m[i]], [v[i]], [gr], st, lr, betas, eps, wd, scale)
        for i, p in enumerate(params):
            stt = opt.state[p]
            assert set(stt.keys()) >= {"step", "float32copy", "exp_avg", "exp_avg_sq"}
            for got, want, what in ((stt["float32copy"], master[i], "master"), (stt["exp_avg"], m[i], "exp_avg"),
                                    (stt["exp_avg_sq"], v[i], "exp_avg_sq")):
                err = float((got - want).abs().max())
                tol = 4e-6 * float(want.abs().max()) + 1e-12
                assert err <= tol, (step, i, what, err, tol)
            # the parameter is the rounded master copy
            assert torch.equal(p.detach(), stt["float32copy"].to(pdtype))
    assert int(opt.state[params[1]]["step"]) == 4 and int(opt.state[params[0]]["step"]) == 5
    # state dict round trip keeps the reference's keys
    sd = opt.state_dict()
    opt2 = AdamWFP32Copy(params, lr=lr, betas=betas, eps=eps, weight_decay=wd)
    opt2.load_state_dict(sd)
    # torch's Optimizer.load_state_dict casts floating state to the parameter dtype (bf16 here), for the reference too
    assert torch.equal(opt2.state[params[0]]["float32copy"].float(),
                       opt.state[params[0]]["float32copy"].to(pdtype).float())
    for p, gr in zip(params, grads):
        p.grad = gr
    opt2.step()                      # the kernel re-widens a downcast state to fp32 instead of failing
    assert opt2.state[params[0]]["float32copy"].dtype == torch.float32


def test_grad_stats_match_torch(dev):
    from atlas_b200.optim import compute_grad_stats, grad_stats_tensor

    g = torch.Generator(device="cpu").manual_seed(9)
    shapes = [(768, 768), (3,), (4096, 65), (1,), (200000,)]
    dts = [torch.bfloat16 if i % 2 == 0 else torch.float32 for i in range(len(shapes))]     # grad dtype = parameter dtype
    params = [torch.nn.Parameter(torch.zeros(*s, dtype=dt, device=dev)) for s, dt in zip(shapes, dts)]
    for i, p in enumerate(params):
        p.grad = (torch.randn(*shapes[i], generator=g) * (0.1 + i)).to(dts[i]).to(dev)
    params[3].grad = None
    stats = grad_stats_tensor(params).cpu()
    for i, p in enumerate(params):
        if p.grad is None:
            assert stats[i].tolist() == [0.0, 0.0, 0.0, 0.0]
            continue
        a = p.grad.float().abs()
        want = torch.stack([a.min(), a.max(), a.mean(), torch.linalg.norm(p.grad.float())]).cpu()
        assert stats[i, 0] == want[0] and stats[i, 1] == want[1]
        assert abs(float(stats[i, 2] - want[2])) <= 1e-5 * float(want[2])
        assert abs(float(stats[i, 3] - want[3])) <= 1e-5 * float(want[3])

    class _M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.reader = torch.nn.Module()
            for i, p in enumerate(params):
                self.reader.register_parameter(f"p{i}", p)

    res = compute_grad_stats(_M())
    live = [p for p in params if p.grad is not None]
    assert res["skip_example"] is False
    assert res["min"] == 0.0                                   # the parameter without gradient contributes (0, 0, 0, 0)
    assert res["max"] == max(float(p.grad.float().abs().max()) for p in live)
    want_mean = sum(float(p.grad.float().abs().mean()) for p in live) / len(params)
    assert abs(res["mean"] - want_mean) <= 1e-5 * want_mean
    params[2].grad[5, 5] = float("nan")
    assert compute_grad_stats(_M())["skip_example"] is True
    params[2].grad[5, 5] = float("inf")
    assert compute_grad_stats(_M())["skip_example"] is True
