"""GPU: dropout of the training path (csrc/dropout.cu, csrc/dropout.cuh, the dropout branches of csrc/attention.cu and
csrc/attention_bwd.cu) - the reference's nn.Dropout sites `src/modeling_t5.py:266,286,310,515-516,561,597,964,1059`,
`src/modeling_bert.py:246,356,384,463`.

torch's own Philox stream cannot be reproduced element for element by a fused kernel, so the gates are:
  * distribution: realised keep rate within 5 sigma of 1 - p, survivors scaled by exactly 1 / (1 - p_eff);
  * exactness given the mask: the kernels export the keep mask of a (seed, offset) key; forward AND backward must agree with
    torch autograd of an fp32 restatement of the reference op fed that mask (same tolerances as tests/test_backward_gpu.py);
  * determinism: same torch seed -> same result, gradient checkpointing (recompute in the backward) -> identical gradients.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float16]
REL = {torch.bfloat16: 4e-2, torch.float16: 6e-3}


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from atlas_b200._lib import lib

    lib()
    return torch.device("cuda:0")


def close(got, ref, rel, what=""):
    got, ref = got.float(), ref.float()
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max())
    l2 = float((got - ref).norm() / (ref.norm() + 1e-30))
    assert err <= rel * scale + 1e-6, f"{what}: max err {err:.3e} vs {rel:.1e} x {scale:.3e}"
    assert l2 <= rel, f"{what}: relative L2 error {l2:.3e} > {rel:.1e}"


def peek_key(dev):
    """The (seed, offset) the NEXT dropout site will draw (grad_ops.next_dropout_key without advancing)."""
    gen = torch.cuda.default_generators[dev.index]
    return gen.initial_seed() & 0xFFFFFFFFFFFFFFFF, gen.get_offset()


def p_eff(p):
    return int(p * 65536 + 0.5) / 65536.0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("p", [0.1, 0.5])
def test_elementwise_dropout(dev, dtype, p):
    from atlas_b200 import grad_ops, ops

    torch.manual_seed(1234)
    M, N = 777, 768
    x = torch.randn(M, N, device=dev).to(dtype).requires_grad_()
    res = torch.randn(M, N, device=dev).to(dtype).requires_grad_()
    seed, off = peek_key(dev)
    y = grad_ops.dropout(x, p, residual=res)
    mask = ops.dropout_mask(M, N, p, seed, off, dev).bool()
    # distribution
    keep = float(mask.float().mean())
    sigma = math.sqrt(p_eff(p) * (1 - p_eff(p)) / (M * N))
    assert abs(keep - (1 - p_eff(p))) < 5 * sigma, (keep, 1 - p_eff(p), sigma)
    # per-row and per-column rates are not degenerate (a striped generator would show here)
    assert float(mask.float().mean(0).std()) < 4 * math.sqrt(p_eff(p) * (1 - p_eff(p)) / M)
    assert float(mask.float().mean(1).std()) < 4 * math.sqrt(p_eff(p) * (1 - p_eff(p)) / N)
    # exact value given the mask: round(x / (1 - p)) where kept, then the residual add (two roundings like two torch ops)
    inv = 1.0 / (1.0 - p_eff(p))
    want = (torch.where(mask, x.detach().float() * inv, torch.zeros((), device=dev)).to(dtype).float()
            + res.detach().float()).to(dtype)
    assert torch.equal(y.detach(), want)
    # backward: the same mask on dy; the residual passes dy through
    dy = torch.randn(M, N, device=dev).to(dtype)
    y.backward(dy)
    assert torch.equal(x.grad, torch.where(mask, dy.float() * inv, torch.zeros((), device=dev)).to(dtype))
    assert torch.equal(res.grad, dy)
    # a different site (offset) gives a different mask; the same key reproduces it
    seed2, off2 = peek_key(dev)
    assert off2 != off
    assert not torch.equal(ops.dropout_mask(M, N, p, seed2, off2, dev).bool(), mask)
    assert torch.equal(ops.dropout_mask(M, N, p, seed, off, dev).bool(), mask)
    # p = 0 is the identity
    assert torch.equal(grad_ops.dropout(x, 0.0), x)


def _ref_attention_dropout(q, k, v, add_mask, bias_delta, scale, causal_value, keep, inv):
    """fp32 restatement of T5Attention / BertSelfAttention with dropout on the probabilities, fed the keep mask."""
    B, Lq, H, _ = q.shape
    Lk = k.shape[1]
    s = torch.einsum("bihd,bjhd->bhij", q, k) * scale
    i = torch.arange(Lq, device=q.device)[:, None]
    j = torch.arange(Lk, device=q.device)[None, :]
    if bias_delta is not None:
        s = s + bias_delta[:, (j - i + Lq - 1)][None]
    if add_mask is not None:
        s = s + add_mask[:, None, None, :]
    if causal_value != 0.0:
        s = s + (j > i).float()[None, None] * causal_value
    p = torch.softmax(s, dim=-1) * keep.float() * inv
    return torch.einsum("bhij,bjhd->bihd", p, v).reshape(B, Lq, H * 64)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,H,L,use_bias,use_mask,scale,causal", [
    (3, 2, 100, True, True, 1.0, 0.0),        # T5 encoder style, ragged tile
    (2, 12, 384, True, True, 1.0, 0.0),       # FiD-base passage segment
    (4, 3, 64, False, True, 0.125, 0.0),      # BERT / Contriever
    (2, 2, 7, True, False, 1.0, -10000.0),    # T5 decoder self-attention (causal)
    (1, 1, 512, False, False, 0.125, 0.0),    # Contriever maximum length
])
def test_self_attention_dropout_fwd_bwd(dev, dtype, B, H, L, use_bias, use_mask, scale, causal):
    from atlas_b200 import grad_ops, ops

    p = 0.1
    torch.manual_seed(99 + L)
    g = torch.Generator(device="cpu").manual_seed(17 + L)
    qkv = (torch.randn(B * L, 3 * H * 64, generator=g) * (0.35 if scale == 1.0 else 1.0)).to(dtype).to(dev).requires_grad_()
    bias = (0.5 * torch.randn(H, 2 * L - 1, generator=g)).to(dev).requires_grad_() if use_bias else None
    mask = None
    if use_mask:
        lens = torch.randint(max(1, L // 3), L + 1, (B,), generator=g)
        lens[0] = L
        mask = ((torch.arange(L)[None, :] >= lens[:, None]).float() * -10000.0).to(dev)
    dout = torch.randn(B * L, H * 64, generator=g).to(dtype).to(dev)
    seed, off = peek_key(dev)
    out = grad_ops.self_attention(qkv, B, H, L, add_mask=mask, bias_delta=bias, scale=scale, causal_value=causal,
                                  dropout_p=p)
    out.backward(dout)
    keep = ops.attention_dropout_mask(B, H, L, L, p, seed, off, dev).bool()
    rate = float(keep.float().mean())
    n = keep.numel()
    assert abs(rate - (1 - p_eff(p))) < 5 * math.sqrt(p_eff(p) * (1 - p_eff(p)) / n) + 1e-9

    ref_in = qkv.detach().float().requires_grad_()
    q, k, v = (t.reshape(B, L, H, 64) for t in ref_in.split(H * 64, dim=1))
    br = bias.detach().clone().requires_grad_() if use_bias else None
    ref = _ref_attention_dropout(q, k, v, mask, br, scale, causal, keep, 1.0 / (1.0 - p_eff(p))).reshape(B * L, H * 64)
    ref.backward(dout.float())
    rel = REL[dtype]
    close(out, ref, rel / 2, "attention fwd (dropout)")
    gq, gk, gv = qkv.grad.float().split(H * 64, dim=1)
    rq, rk, rv = ref_in.grad.split(H * 64, dim=1)
    close(gv, rv, rel, "dV")
    close(gq, rq, rel, "dQ")
    close(gk, rk, rel, "dK")
    if use_bias:
        close(bias.grad, br.grad, rel, "dbias")
    # the dropped forward differs from the plain one (the mask is really applied)
    plain = grad_ops.self_attention(qkv.detach(), B, H, L, add_mask=mask, bias_delta=bias.detach() if use_bias else None,
                                    scale=scale, causal_value=causal)
    assert float((plain.float() - out.detach().float()).abs().max()) > 1e-3


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,H,T,Lk,split", [(2, 2, 5, 640, 320), (1, 12, 32, 1152, 384), (3, 1, 1, 192, 192)])
def test_cross_attention_dropout_fwd_bwd(dev, dtype, B, H, T, Lk, split):
    from atlas_b200 import grad_ops, ops

    p = 0.1
    torch.manual_seed(7 + T)
    g = torch.Generator(device="cpu").manual_seed(23 + T)
    q = (torch.randn(B * T, H * 64, generator=g) * 0.4).to(dtype).to(dev).requires_grad_()
    kv = (torch.randn(B * Lk, 2 * H * 64, generator=g) * 0.4).to(dtype).to(dev).requires_grad_()
    valid = torch.rand(B, Lk, generator=g) > 0.2
    valid[:, 0] = True
    neg = -1e4 if dtype == torch.float16 else -1e9
    mask = ((~valid).float() * neg).to(dev)
    dout = torch.randn(B * T, H * 64, generator=g).to(dtype).to(dev)
    seed, off = peek_key(dev)
    out = grad_ops.cross_attention(q, kv, B, H, T, Lk, add_mask=mask, scale=1.0, split=split, dropout_p=p)
    out.backward(dout)
    keep = ops.attention_dropout_mask(B, H, T, Lk, p, seed, off, dev).bool()

    qr, kvr = q.detach().float().requires_grad_(), kv.detach().float().requires_grad_()
    k, v = (t.reshape(B, Lk, H, 64) for t in kvr.split(H * 64, dim=1))
    ref = _ref_attention_dropout(qr.reshape(B, T, H, 64), k, v, mask, None, 1.0, 0.0, keep,
                                 1.0 / (1.0 - p_eff(p))).reshape(B * T, H * 64)
    ref.backward(dout.float())
    rel = REL[dtype]
    close(out, ref, rel / 2, "cross fwd (dropout)")
    close(q.grad, qr.grad, rel, "dq")
    gk, gv = kv.grad.float().split(H * 64, dim=1)
    rk, rv = kvr.grad.split(H * 64, dim=1)
    close(gv, rv, rel, "dV")
    close(gk, rk, rel, "dK")


def _tiny_fid(dev, dropout):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import model_synth
    from atlas_b200.fid import FiD, T5ConfigLite

    cfg = {k: v for k, v in model_synth.T5_CFG.items() if k not in ("dropout_rate", "is_encoder_decoder", "use_cache")}
    reader = FiD(T5ConfigLite(dropout_rate=dropout, **cfg))
    sd, _ = model_synth.fill_state_dict(reader.state_dict(), 202)
    reader.load_state_dict(sd)
    reader = reader.to(torch.bfloat16).to(dev)
    reader.encoder.config.n_context, reader.encoder.config.bsz = 3, 2
    ids, mask, labels = model_synth.fid_inputs()
    return reader, ids.to(dev), mask.to(dev), labels.to(dev)


def _step(reader, ids, mask, labels, seed):
    torch.manual_seed(seed)
    reader.zero_grad(set_to_none=True)
    out = reader(input_ids=ids, attention_mask=mask, labels=labels)
    out[0].backward()
    grads = {n: p.grad.detach().float().clone() for n, p in reader.named_parameters() if p.grad is not None}
    return float(out[0]), grads


def test_fid_train_step_with_dropout(dev):
    """FiD in .train() with the reference's default --dropout 0.1: the loss moves with the torch seed, is reproducible for a
    fixed seed, is finite with finite gradients on every parameter, gradient checkpointing (the block forward is recomputed
    in the backward and must re-derive the SAME masks) changes nothing, and .eval() ignores dropout."""
    reader, ids, mask, labels = _tiny_fid(dev, 0.1)
    reader.train()
    l1, g1 = _step(reader, ids, mask, labels, 5)
    l1b, g1b = _step(reader, ids, mask, labels, 5)
    l2, _ = _step(reader, ids, mask, labels, 6)
    assert math.isfinite(l1) and l1 == l1b and l1 != l2
    assert all(torch.isfinite(v).all() for v in g1.values()) and len(g1) > 40
    def same(a, b):     # norm-weight / embedding gradients accumulate with fp32 atomics: equal up to summation order;
        return float((a - b).abs().max()) <= 1e-3 * float(a.abs().max()) + 1e-7     # a different MASK would be an O(1) change

    for n in g1:
        assert same(g1[n], g1b[n]), n
    reader.gradient_checkpointing_enable()
    l1c, g1c = _step(reader, ids, mask, labels, 5)
    assert l1c == l1
    for n in g1:
        assert same(g1[n], g1c[n]), f"checkpointed recompute drew different masks: {n}"
    # dropout really acts: the undropped loss differs, and is what --dropout 0 gives
    ref0, ids0, mask0, labels0 = _tiny_fid(dev, 0.0)
    ref0.train()
    l0, _ = _step(ref0, ids0, mask0, labels0, 5)
    assert abs(l0 - l1) > 1e-4
    reader.eval()
    with torch.no_grad():
        le = float(reader(input_ids=ids, attention_mask=mask, labels=labels)[0])
    assert abs(le - l0) < 5e-2     # eval = no dropout (graph / fused path vs the autograd path: bf16 noise only)


def test_contriever_train_step_with_dropout(dev):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import model_synth
    from atlas_b200.retrievers import BertConfigLite, Contriever

    cfg = dict(model_synth.CONTRIEVER_CFG)
    cfg["hidden_dropout_prob"], cfg["attention_probs_dropout_prob"] = 0.1, 0.1
    model = Contriever(BertConfigLite(**cfg))
    sd, _ = model_synth.fill_state_dict(model.state_dict(), 101)
    model.load_state_dict(sd)
    model = model.to(torch.bfloat16).to(dev).train()
    g = torch.Generator(device="cpu").manual_seed(3)
    ids = torch.randint(1, cfg["vocab_size"], (4, 48), generator=g).to(dev)
    mask = torch.ones_like(ids)
    mask[1, 30:] = 0

    def step(seed):
        torch.manual_seed(seed)
        model.zero_grad(set_to_none=True)
        emb = model(input_ids=ids, attention_mask=mask)
        (emb.float() ** 2).sum().backward()
        return emb.detach().float().clone(), {n: p.grad.float().clone() for n, p in model.named_parameters() if p.grad is not None}

    e1, g1 = step(11)
    e1b, g1b = step(11)
    e2, _ = step(12)
    assert torch.equal(e1, e1b) and not torch.equal(e1, e2)
    assert all(torch.isfinite(v).all() for v in g1.values())
    for n in g1:
        assert float((g1[n] - g1b[n]).abs().max()) <= 1e-3 * float(g1[n].abs().max()) + 1e-7, n
    model.eval()
    with torch.no_grad():
        ee = model(input_ids=ids, attention_mask=mask).float()
    assert float((ee - e1).abs().max()) > 1e-3      # training-mode output is the dropped one
