"""GPU numerics of the normalisation / embedding / pooling / attention kernels against plain PyTorch
fp32 references of the same ops (rounding points restated from the reference model code)."""
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


def _bert_ln_ref(x16, w, b, eps):
    """BertLayerNorm on `.float()` input then `.type_as` (src/modeling_bert.py:104-114,386)."""
    x = x16.float()
    mean = x.mean(-1, keepdim=True)
    var = x.pow(2).mean(-1, keepdim=True)
    h = ((x - mean) * torch.rsqrt(var + eps)).to(w.dtype)
    return w * h + b


def _rms_ref(x16, w, eps):
    var = x16.float().pow(2).mean(-1, keepdim=True)
    h = (x16 * torch.rsqrt(var + eps)).to(w.dtype)
    return w * h


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("H", [768, 2048, 64])
def test_layernorm_kinds(dev, dtype, H):
    from atlas_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(H)
    x = (torch.randn(1000, H, generator=g) * 2 + 0.3).to(dtype).to(dev)
    w = (1 + 0.1 * torch.randn(H, generator=g)).to(dtype).to(dev)
    b = (0.1 * torch.randn(H, generator=g)).to(dtype).to(dev)
    y0 = ops.layernorm(x, w, b, eps=1e-12, kind=0)
    r0 = _bert_ln_ref(x, w, b, 1e-12)
    y1 = ops.layernorm(x, w, None, eps=1e-6, kind=1)
    r1 = _rms_ref(x, w, 1e-6)
    for y, r, bb in ((y0, r0, b), (y1, r1, torch.zeros_like(b))):
        # identical rounding points; the fp32 statistics may differ in reduction order, which can move the
        # 16-bit intermediate w*h by one ulp -> tolerance in ulps of the INTERMEDIATE magnitude (|y| + |b|),
        # and almost all outputs must be bit-identical
        ulp = torch.finfo(dtype).eps * (r.float().abs() + bb.float().abs()) + 1e-6
        assert torch.all((y.float() - r.float()).abs() <= 2.0 * ulp), float((y.float() - r.float()).abs().max())
        assert float((y == r).float().mean()) > 0.97


def test_bert_embed_ln(dev):
    from atlas_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(3)
    V, L, H, B = 1000, 37, 768, 5
    we = (torch.randn(V, H, generator=g) * 0.02).half().to(dev)
    te = (torch.randn(2, H, generator=g) * 0.02).half().to(dev)
    pe = (torch.randn(512, H, generator=g) * 0.02).half().to(dev)
    w = (1 + 0.1 * torch.randn(H, generator=g)).half().to(dev)
    b = (0.1 * torch.randn(H, generator=g)).half().to(dev)
    ids = torch.randint(0, V, (B, L), generator=g).to(dev)
    y = ops.bert_embed_ln(ids, None, we, te, pe, w, b, 1e-12)
    emb = we[ids] + te[torch.zeros_like(ids)]
    emb += pe[torch.arange(L, device=dev)][None]
    ref = _bert_ln_ref(emb, w, b, 1e-12)
    tol = 2 * torch.finfo(torch.float16).eps * (ref.float().abs() + b.float().abs()) + 1e-6
    assert torch.all((y.float() - ref.float()).abs() <= tol)
    assert float((y == ref).float().mean()) > 0.97


def test_masked_mean_pool_into_bank_rows(dev):
    from atlas_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(4)
    B, L, H = 9, 50, 768
    x = torch.randn(B, L, H, generator=g).half().to(dev)
    lens = torch.randint(1, L + 1, (B,), generator=g)
    mask = (torch.arange(L)[None] < lens[:, None]).long().to(dev)
    bank = torch.zeros(20, H, dtype=torch.float16, device=dev)
    ops.masked_mean_pool(x, mask, out=bank[4:4 + B])
    last = x.masked_fill(~mask[..., None].bool(), 0.0)
    ref = last.sum(dim=1) / mask.sum(dim=1)[..., None]       # src/retrievers.py:50-53 on a half tensor
    assert torch.all((bank[4:4 + B].float() - ref.float()).abs() <= torch.finfo(torch.float16).eps * ref.float().abs() + 1e-7)
    assert float(bank[:4].abs().max()) == 0 and float(bank[4 + B:].abs().max()) == 0


def _attn_ref(qkv, B, H, L, add_mask, bias_delta, scale, causal_value):
    """fp32 reference: scores = scale*QK^T + bias + mask, softmax, PV (inputs are the 16-bit values)."""
    d = 64
    q = qkv[:, : H * d].float().view(B, L, H, d).permute(0, 2, 1, 3)
    k = qkv[:, H * d: 2 * H * d].float().view(B, L, H, d).permute(0, 2, 1, 3)
    v = qkv[:, 2 * H * d:].float().view(B, L, H, d).permute(0, 2, 1, 3)
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    if bias_delta is not None:
        i = torch.arange(L, device=qkv.device)
        idx = i[None, :] - i[:, None] + (L - 1)              # [i, j] -> j - i + L - 1
        s = s + bias_delta[:, idx][None]
    if add_mask is not None:
        s = s + add_mask[:, None, None, :]
    if causal_value != 0.0:
        i = torch.arange(L, device=qkv.device)
        s = s + (i[None, :] > i[:, None]).float() * causal_value
    p = torch.softmax(s, dim=-1)
    return torch.matmul(p, v).permute(0, 2, 1, 3).reshape(B * L, H * d)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,H,L,mode", [(3, 12, 384, "t5enc"), (2, 12, 128, "bert"), (5, 12, 173, "bert"),
                                        (2, 12, 512, "bert"), (4, 12, 32, "t5dec"), (1, 32, 384, "t5enc")])
def test_attention_matches_fp32_reference(dev, dtype, B, H, L, mode):
    from atlas_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(B * 100 + L)
    qkv = (torch.randn(B * L, 3 * H * 64, generator=g) * (0.35 if mode == "bert" else 0.12)).to(dtype).to(dev)
    lens = torch.randint(max(1, L // 3), L + 1, (B,), generator=g)
    keep = (torch.arange(L)[None] < lens[:, None]).float().to(dev)
    add_mask = (1.0 - keep) * -10000.0
    bias = None
    scale, causal = 1.0, 0.0
    if mode == "bert":
        scale = 1.0 / 8.0
    else:
        bias = (torch.randn(H, 2 * L - 1, generator=g) * 0.5).to(dev)
    if mode == "t5dec":
        causal = -10000.0
    out = ops.attention(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, B, H, L, L, add_mask=add_mask, bias_delta=bias, scale=scale,
                        causal_value=causal)
    ref = _attn_ref(qkv, B, H, L, add_mask, bias, scale, causal)
    err = (out.float() - ref).abs()
    tol = 2e-3 if dtype == torch.float16 else 1.2e-2   # P and O are rounded to the 16-bit type (8 / 11 mantissa bits)
    assert float(err.max()) <= tol * max(1.0, float(ref.abs().max())), float(err.max())
    # valid rows only matter, but every row must be finite
    assert torch.isfinite(out.float()).all()


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,H,L,mode", [(3, 12, 384, "nomask"), (2, 4, 200, "causal"), (2, 12, 384, "ramp"),
                                        (3, 12, 256, "bert_nomask"), (37, 12, 384, "t5enc"), (2, 12, 480, "t5enc")])
def test_attention_lanes_kernel_cases(dev, dtype, B, H, L, mode):
    """The three-lane kernel (csrc/attention_lanes.cu, >= 2 query tiles per segment) on the cases the generic test does not
    reach: no key mask at all (the bench's shape), causal + bias with several query tiles, a bias RAMP that makes every
    later key block exceed the lazy reference maximum by more than 2^8 (the in-place rescale of O runs on every block),
    more (segment, head) items than SMs x 3 with the bias table changing inside a CTA's item range, and 5 key blocks."""
    from atlas_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(B * 1000 + L + len(mode))
    bert = mode.startswith("bert")
    qkv = (torch.randn(B * L, 3 * H * 64, generator=g) * (0.35 if bert else 0.12)).to(dtype).to(dev)
    add_mask, bias, scale, causal = None, None, 1.0, 0.0
    if mode in ("t5enc", "ramp", "causal"):
        lens = torch.randint(max(1, L // 3), L + 1, (B,), generator=g)
        add_mask = ((1.0 - (torch.arange(L)[None] < lens[:, None]).float()) * -10000.0).to(dev)
    if bert:
        scale = 1.0 / 8.0
    else:
        bias = (torch.randn(H, 2 * L - 1, generator=g) * 0.5)
        if mode == "ramp":
            bias = bias + 0.15 * torch.arange(2 * L - 1)[None, :].float()      # +14 (natural log units) per 96-key block
        bias = bias.to(dev)
    if mode == "causal":
        causal = -10000.0
    out, lse = ops.attention(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, B, H, L, L, add_mask=add_mask, bias_delta=bias,
                             scale=scale, causal_value=causal, return_lse=True)
    ref = _attn_ref(qkv, B, H, L, add_mask, bias, scale, causal)
    err = (out.float() - ref).abs()
    tol = 2e-3 if dtype == torch.float16 else 1.2e-2
    assert torch.isfinite(out.float()).all()
    assert float(err.max()) <= tol * max(1.0, float(ref.abs().max())), float(err.max())
    # the saved log-sum-exp (backward pass input) against the fp32 scores
    d = 64
    q = qkv[:, : H * d].float().view(B, L, H, d).permute(0, 2, 1, 3)
    k = qkv[:, H * d: 2 * H * d].float().view(B, L, H, d).permute(0, 2, 1, 3)
    s = torch.matmul(q, k.transpose(-1, -2)) * scale
    if bias is not None:
        idx = torch.arange(L, device=dev)[None, :] - torch.arange(L, device=dev)[:, None] + (L - 1)
        s = s + bias[:, idx][None]
    if add_mask is not None:
        s = s + add_mask[:, None, None, :]
    if causal != 0.0:
        i = torch.arange(L, device=dev)
        s = s + (i[None, :] > i[:, None]).float() * causal
    assert float((lse - torch.logsumexp(s, dim=-1)).abs().max()) <= 2e-3


def test_fp16_inf_clamp(dev):
    """ops.clamp_inf_ = the reference's `if dtype == fp16 and isinf(h).any(): h = clamp(h, +-(65504 - 1000))`
    (src/modeling_t5.py:657-708) decided on the device: untouched without an inf, torch.clamp's values with one."""
    from atlas_b200 import ops

    g = torch.Generator().manual_seed(3)
    x = (torch.randn(777, 768, generator=g) * 30000).half()            # plenty of values between 64512 and 65504, some inf
    finite = x.clone()
    finite[torch.isinf(finite)] = 65504.0
    assert not torch.isinf(finite).any() and (finite.abs() > 64512).any()
    y = ops.clamp_inf_(finite.clone().to(dev))
    assert torch.equal(y.cpu(), finite)                                 # no inf anywhere: nothing is clamped
    withinf = finite.clone()
    withinf[5, 7] = float("inf")
    withinf[700, 0] = float("-inf")
    withinf[3, 3] = float("nan")
    clamp_value = torch.finfo(torch.float16).max - 1000
    want = torch.clamp(withinf.float(), min=-clamp_value, max=clamp_value).half()
    ss = torch.zeros(777, dtype=torch.float32, device=dev)
    got = ops.clamp_inf_(withinf.clone().to(dev), row_ss=ss)
    assert torch.equal(torch.nan_to_num(got.cpu().float(), nan=-1.0), torch.nan_to_num(want.float(), nan=-1.0))
    ref_ss = want.float().pow(2).sum(-1)
    ok = ~torch.isnan(ref_ss)
    assert torch.allclose(ss.cpu()[ok], ref_ss[ok], rtol=1e-5)
    # other dtypes pass through (the reference tests `dtype == torch.float16`)
    b = torch.full((8, 768), float("inf"), dtype=torch.bfloat16, device=dev)
    assert torch.isinf(ops.clamp_inf_(b)).all()
    # strided view (a column slice of a wider buffer)
    wide = torch.zeros(64, 1536, dtype=torch.float16, device=dev)
    wide[:, 768:] = 65000.0
    wide[1, 800] = float("inf")
    ops.clamp_inf_(wide[:, 768:])
    assert float(wide[:, 768:].max()) == 64512.0 and float(wide[:, :768].abs().max()) == 0.0


def test_fid_fp16_overflow_is_clamped_like_the_reference(dev):
    """A FiD block whose feed-forward output overflows fp16: with the device-side clamp the forward stays finite and equals an
    fp16 torch restatement of the reference's block arithmetic (clamp included) on the same kernels' inputs."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import model_synth
    from atlas_b200.fid import FiD, T5ConfigLite

    cfg = {k: v for k, v in model_synth.T5_CFG.items() if k not in ("dropout_rate", "is_encoder_decoder", "use_cache")}
    reader = FiD(T5ConfigLite(**cfg))
    sd, _ = model_synth.fill_state_dict(reader.state_dict(), 202)
    sd["encoder.block.0.layer.1.DenseReluDense.wo.weight"] = sd["encoder.block.0.layer.1.DenseReluDense.wo.weight"] * 3000.0
    reader.load_state_dict(sd)
    reader = reader.half().to(dev).eval()
    reader.encoder.config.n_context, reader.encoder.config.bsz = 3, 2
    ids, mask, labels = model_synth.fid_inputs()
    for fuse in (True, False):
        reader.fuse_norm = fuse
        reader.cuda_graphs = False
        with torch.no_grad():
            enc = reader.encode(ids.to(dev), mask.to(dev))
        assert torch.isfinite(enc).all(), f"fuse_norm={fuse}: the overflow reached the encoder output"
        assert float(enc.abs().max()) > 0


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,H,T,Lk", [(2, 3, 32, 4000), (1, 12, 7, 15360), (3, 2, 64, 1088), (2, 1, 1, 2048)])
def test_cross_attention_stream_kernel(dev, dtype, B, H, T, Lk):
    """csrc/attention_stream.cu (few target tokens against the concatenated encoder keys, src/fid.py:298-349): against an fp32
    restatement, ragged key counts, masked keys, and against the tcgen05 split-KV path where that one applies."""
    from atlas_b200 import ops

    g = torch.Generator().manual_seed(B * 1000 + T)
    q = (torch.randn(B * T, H * 64, generator=g) * 0.5).to(dtype).to(dev)
    kv = (torch.randn(B * Lk, 2 * H * 64, generator=g) * 0.5).to(dtype).to(dev)
    valid = torch.rand(B, Lk, generator=g) > 0.25
    valid[:, 0] = True
    valid[0, Lk // 2:] = False                       # a long fully-masked tail (whole chunks without a live key)
    neg = -1e4 if dtype == torch.float16 else -1e9
    mask = ((~valid).float() * neg).to(dev)
    assert ops._XATTN_STREAM
    out, lse = ops.cross_attention_split(q, 0, kv, 0, H * 64, B, H, T, Lk, add_mask=mask, scale=1.0, split=Lk, return_lse=True)
    qf = q.float().reshape(B, T, H, 64)
    kf, vf = (t.reshape(B, Lk, H, 64) for t in kv.float().split(H * 64, dim=1))
    s = torch.einsum("bihd,bjhd->bhij", qf, kf) + mask[:, None, None, :]
    ref = torch.einsum("bhij,bjhd->bihd", torch.softmax(s, -1), vf).reshape(B * T, H * 64)
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    err = float((out.float() - ref).abs().max())
    assert err <= tol * max(1.0, float(ref.abs().max())), err
    assert torch.allclose(lse, torch.logsumexp(s, -1), rtol=1e-3, atol=2e-3)
    # skipping the fully masked 64-key tiles (the default above) does not change a bit
    ops._SKIP_MASKED = False
    try:
        full = ops.cross_attention_split(q, 0, kv, 0, H * 64, B, H, T, Lk, add_mask=mask, scale=1.0, split=Lk)
    finally:
        ops._SKIP_MASKED = True
    assert torch.equal(full, out)
    split = next((x for x in range(384, 63, -1) if Lk % x == 0), None)
    if split is not None:
        ops._XATTN_STREAM = False
        try:
            old = ops.cross_attention_split(q, 0, kv, 0, H * 64, B, H, T, Lk, add_mask=mask, scale=1.0, split=split)
        finally:
            ops._XATTN_STREAM = True
        assert float((old.float() - out.float()).abs().max()) <= tol * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("S,H,L,use_bias", [(7, 3, 384, True), (5, 2, 200, True), (4, 12, 512, False), (3, 1, 512, True)])
def test_masked_key_blocks_are_skipped_without_changing_the_result(dev, dtype, S, H, L, use_bias):
    """ops.key_block_live: 64-key blocks made of masked keys only (padding to text_maxlength) are neither loaded nor computed
    by the three-lane encoder kernel.  Their softmax weights are exactly 0 in fp32, so the output must be BIT-identical to
    the run that computes every block - and equal to the fp32 restatement."""
    from atlas_b200 import ops

    g = torch.Generator().manual_seed(S * 100 + L)
    qkv = (torch.randn(S * L, 3 * H * 64, generator=g) * 0.3).to(dtype).to(dev)
    bias = (0.5 * torch.randn(H, 2 * L - 1, generator=g)).to(dev) if use_bias else None
    lens = torch.randint(1, L + 1, (S,), generator=g)
    lens[0], lens[1] = L, 3                                     # nothing to skip / almost everything
    mask = ((torch.arange(L)[None, :] >= lens[:, None]).float() * -10000.0).to(dev)
    if S > 3:
        mask[3, 70:140] = -10000.0                              # a hole of masked keys in the middle (one whole block)
    live = ops.key_block_live(mask)
    assert live is not None and live.shape == (S, (L + 63) // 64) and int(live[1].sum()) == 1 and bool(live[0].all())
    a = ops.attention(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, S, H, L, L, add_mask=mask, bias_delta=bias, block_live=live)
    b = ops.attention(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, S, H, L, L, add_mask=mask, bias_delta=bias)
    assert torch.equal(a, b), float((a.float() - b.float()).abs().max())
    q, k, v = (t.float().reshape(S, L, H, 64) for t in qkv.split(H * 64, dim=1))
    s = torch.einsum("bihd,bjhd->bhij", q, k) + mask[:, None, None, :]
    if use_bias:
        i = torch.arange(L, device=dev)[:, None]
        j = torch.arange(L, device=dev)[None, :]
        s = s + bias[:, (j - i + L - 1)][None]
    ref = torch.einsum("bhij,bjhd->bihd", torch.softmax(s, -1), v).reshape(S * L, H * 64)
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    assert float((a.float() - ref).abs().max()) <= tol * max(1.0, float(ref.abs().max()))
    # a fully masked segment keeps every block (uniform-over-masked-keys softmax like the reference)
    allm = torch.full((2, L), -10000.0, device=dev)
    assert bool(ops.key_block_live(allm).all())



@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_compacted_cross_kv_path_is_identical(dev, dtype):
    """Cross K | V projected only for the encoder positions inside live 64-key tiles (ops.compact_live_tiles + linear_dynm with
    a device-side row count + the stream kernel reading the compacted rows) against the dense projection + the same kernel:
    bit-identical attention output; the compaction tables against a torch restatement."""
    from atlas_b200 import ops

    g = torch.Generator().manual_seed(5)
    B, H, T, n, Lp, d = 3, 12, 32, 5, 384, 768
    Lk = n * Lp
    enc = (torch.randn(B * Lk, d, generator=g) * 0.5).to(dtype).to(dev)
    wkv = (torch.randn(2 * H * 64, d, generator=g) / 27.7).to(dtype).to(dev)
    q = (torch.randn(B * T, H * 64, generator=g) * 0.5).to(dtype).to(dev)
    lens = torch.randint(20, Lp + 1, (B, n), generator=g)
    lens[0, 0], lens[1, 2] = Lp, 1
    valid = torch.arange(Lp)[None, None, :] < lens[..., None]
    mask = ((~valid).reshape(B, Lk).float() * -1e9).to(dev)
    live = ops.key_block_live(mask)
    enc_live, tile_off, count = ops.compact_live_tiles(enc, live)
    flags = live.reshape(-1).bool().cpu()
    want_off = torch.where(flags, torch.cumsum(flags.int(), 0) - 1, torch.full_like(flags.int(), -1))
    assert torch.equal(tile_off.cpu(), want_off.int()) and int(count) == 64 * int(flags.sum())
    rows = enc.view(-1, 64, d)[flags.to(dev)].reshape(-1, d)
    assert torch.equal(enc_live[: rows.shape[0]], rows)
    kv_live = ops.linear_dynm(enc_live, wkv, count)
    kv_dense = ops.linear(enc, wkv)
    assert torch.equal(kv_live[: rows.shape[0]].view(-1, 64, 2 * H * 64), kv_dense.view(-1, 64, 2 * H * 64)[flags.to(dev)])
    a = ops.cross_attention_stream_compact(q, kv_live, live, tile_off, B, H, T, Lk, mask, scale=1.0)
    b = ops.cross_attention_split(q, 0, kv_dense, 0, H * 64, B, H, T, Lk, add_mask=mask, scale=1.0, split=384, tile_live=live)
    assert torch.equal(a, b)
