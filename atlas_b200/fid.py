"""Drop-in for `src.fid.FiD` (reference src/fid.py:28-357 over the vendored T5 v1.1, src/modeling_t5.py):
Fusion-in-Decoder on B200 kernels: the no-grad forward (evaluation / scoring / greedy generation, replayed from a
CUDA graph), the with-grad forward + hand-written backward kernels of the training step (`_forward_train`, grad_ops.py),
and the cross-attention score capture used for retriever distillation (src/fid.py:126-235,333-343).

Surface kept: `FiD(config)`, `forward(input_ids [B, n*L], attention_mask, decoder_input_ids,
labels, encoder_outputs=None, use_cache=False)` -> output indexable as `out[0]` = loss, `out[1]` = logits with
`.logits`, `.loss`, `.encoder_last_hidden_state`; `encoder.config.{n_context,bsz}` are read like the reference's
`FiDStack` (src/fid.py:47-49,66-76); `_shift_right`; `generate` (greedy); parameter names identical to HF T5
(`shared`, `encoder.block.N.layer.0.SelfAttention.{q,k,v,o}`, `relative_attention_bias` on block 0,
`layer.1.EncDecAttention`, `DenseReluDense.{wi_0,wi_1,wo}`, `lm_head`) so reference checkpoints load.

What runs where (all in csrc/): RMSNorm (T5LayerNorm, modeling_t5.py:244-253); q/k/v/o and FF projections on the
tcgen05 GEMM with fused residual and gated-GELU epilogues (modeling_t5.py:281-289); encoder self-attention as
B*n independent L-token segments with the relative-position bias added on the fly from a [H, 2L-1] table
(modeling_t5.py:352-416,478-524) - the [B*n, H, L, L] bias / probability tensors are never materialised;
decoder cross-attention over the n*L concatenated keys as split-KV + combine (fid.py:298-349).
Dropout: counter-based masks inside the kernels (csrc/dropout.cuh).  The fp16 `isinf` clamps of every sub-layer
(modeling_t5.py:657-708, three host syncs per block in the reference) are decided on the device (ops.clamp_inf_).
"""
import copy
import math
import os
from types import SimpleNamespace

import torch
from torch import nn

from . import grad_ops, ops
from ._lib import AtlasB200Error
from .retrievers import HalfCache


class T5ConfigLite(SimpleNamespace):
    """The T5Config fields this path reads (defaults = google/t5-base-lm-adapt, i.e. T5 v1.1 base)."""

    def __init__(self, **kw):
        d = dict(vocab_size=32128, d_model=768, d_kv=64, d_ff=2048, num_layers=12, num_decoder_layers=12, num_heads=12,
                 relative_attention_num_buckets=32, layer_norm_epsilon=1e-6, feed_forward_proj="gated-gelu",
                 tie_word_embeddings=False, decoder_start_token_id=0, pad_token_id=0, eos_token_id=1)
        d.update(kw)
        super().__init__(**d)


class _Norm(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))


class _Attn(nn.Module):
    def __init__(self, c, has_bias):
        super().__init__()
        inner = c.num_heads * c.d_kv
        self.q = nn.Linear(c.d_model, inner, bias=False)
        self.k = nn.Linear(c.d_model, inner, bias=False)
        self.v = nn.Linear(c.d_model, inner, bias=False)
        self.o = nn.Linear(inner, c.d_model, bias=False)
        if has_bias:
            self.relative_attention_bias = nn.Embedding(c.relative_attention_num_buckets, c.num_heads)


class _SelfAttnLayer(nn.Module):
    def __init__(self, c, has_bias):
        super().__init__()
        self.SelfAttention = _Attn(c, has_bias)
        self.layer_norm = _Norm(c.d_model)


class _CrossAttnLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.EncDecAttention = _Attn(c, False)
        self.layer_norm = _Norm(c.d_model)


class _FF(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.wi_0 = nn.Linear(c.d_model, c.d_ff, bias=False)
        self.wi_1 = nn.Linear(c.d_model, c.d_ff, bias=False)
        self.wo = nn.Linear(c.d_ff, c.d_model, bias=False)


class _FFLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.DenseReluDense = _FF(c)
        self.layer_norm = _Norm(c.d_model)


class _Block(nn.Module):
    def __init__(self, c, is_decoder, has_bias):
        super().__init__()
        layers = [_SelfAttnLayer(c, has_bias)]
        if is_decoder:
            layers.append(_CrossAttnLayer(c))
        layers.append(_FFLayer(c))
        self.layer = nn.ModuleList(layers)


class FiDStack(nn.Module):
    def __init__(self, config, embed_tokens, is_decoder):
        super().__init__()
        self.config = copy.copy(config)
        self.config.is_decoder = is_decoder
        self.is_decoder = is_decoder
        self.embed_tokens = embed_tokens
        n = config.num_decoder_layers if is_decoder else config.num_layers
        self.block = nn.ModuleList([_Block(config, is_decoder, i == 0) for i in range(n)])
        self.final_layer_norm = _Norm(config.d_model)


def relative_position_bucket(relative_position, bidirectional=True, num_buckets=32, max_distance=128):
    """T5's bucket function restated with the same torch ops and rounding as src/modeling_t5.py:352-397."""
    relative_buckets = 0
    if bidirectional:
        num_buckets //= 2
        relative_buckets += (relative_position > 0).to(torch.long) * num_buckets
        relative_position = torch.abs(relative_position)
    else:
        relative_position = -torch.min(relative_position, torch.zeros_like(relative_position))
    max_exact = num_buckets // 2
    is_small = relative_position < max_exact
    if_large = max_exact + (
        torch.log(relative_position.float() / max_exact) / math.log(max_distance / max_exact) * (num_buckets - max_exact)
    ).to(torch.long)
    if_large = torch.min(if_large, torch.full_like(if_large, num_buckets - 1))
    relative_buckets += torch.where(is_small, relative_position, if_large)
    return relative_buckets


def bias_by_delta(weight, lq, lk, bidirectional, num_buckets):
    """[H, lq+lk-1] fp32 table: entry (j - i) + (lq - 1) holds the head's bias for key j / query i
    (`compute_bias`, src/modeling_t5.py:399-416, indexed by offset instead of a [lq, lk] matrix)."""
    delta = torch.arange(-(lq - 1), lk, device=weight.device, dtype=torch.long)
    buckets = relative_position_bucket(delta, bidirectional=bidirectional, num_buckets=num_buckets)
    return weight[buckets].t().float().contiguous()


class _GraphRunner:
    """One captured CUDA graph of a static-shape function of device tensors.  Every kernel of this package is
    launched through the C ABI on `torch.cuda.current_stream()`, so capturing the whole forward removes the ~150
    host launches of the decoder stack (launch-bound at T = 32 target tokens) from the critical path."""

    def __init__(self, fn, example_inputs):
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):           # eager warm-up: lazy kernel attributes, allocator, weight caches
            fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out


class FiDOutput(tuple):
    """(loss, logits, encoder_last_hidden_state) with the attribute names Atlas reads (src/atlas.py:292-300,583-590)."""

    def __new__(cls, loss, logits, enc):
        o = super().__new__(cls, (loss, logits, enc))
        o.loss, o.logits, o.encoder_last_hidden_state = loss, logits, enc
        return o


class FiD(nn.Module):
    def __init__(self, config=None):
        super().__init__()
        config = config or T5ConfigLite()
        if config.d_kv != 64:
            raise AtlasB200Error("atlas_b200 attention kernels need d_kv = 64 (all T5 v1.1 sizes)")
        if getattr(config, "feed_forward_proj", "gated-gelu") != "gated-gelu":
            raise AtlasB200Error("only T5 v1.1 (gated-gelu) is implemented")
        self.config = config
        self.model_dim = config.d_model
        self.shared = nn.Embedding(config.vocab_size, config.d_model)
        self.encoder = FiDStack(config, self.shared, is_decoder=False)
        self.decoder = FiDStack(config, self.shared, is_decoder=True)
        self.lm_head = nn.Linear(config.d_model, config.vocab_size, bias=False)
        self.encoder.config.n_context = 1
        self.encoder.config.bsz = 1
        self._half = HalfCache()
        # CUDA-graph replay of the static-shape forward (eval / scoring); ATLAS_B200_CUDA_GRAPH=0 or
        # `model.cuda_graphs = False` runs every kernel launch eagerly (needed under profilers' event bracketing)
        self.cuda_graphs = os.environ.get("ATLAS_B200_CUDA_GRAPH", "1") != "0"
        self.fuse_norm = os.environ.get("ATLAS_B200_FUSE_NORM", "1") != "0"   # RMSNorm folded into the encoder GEMMs
        self._graphs = {}

    @classmethod
    def from_pretrained(cls, path, **kw):
        """Local HF-layout directory (config.json + weights) of a T5 v1.1 checkpoint, src/model_io.py:77."""
        from ._pretrained import load_pretrained

        return load_pretrained(cls, T5ConfigLite, path, **kw)

    # ---- reference surface that is configuration only -------------------------------------
    def set_checkpoint(self, use_checkpoint):
        pass

    def gradient_checkpointing_enable(self):
        """Recompute each block in the backward instead of keeping its activations (torch.utils.checkpoint around the
        autograd Functions of grad_ops.py), `--use_gradient_checkpoint_reader` (src/model_io.py:86-87)."""
        self._grad_ckpt = True

    def gradient_checkpointing_disable(self):
        self._grad_ckpt = False

    # ---- cross-attention score capture (src/fid.py:126-235,333-343) -------------------------------------
    def reset_score_storage(self):
        self._xattn = []

    def overwrite_forward_crossattention(self):
        """From now on every forward also records, per decoder layer, the head-means of the cross-attention logits,
        probabilities and ||V||-weighted probabilities (src/model_io.py:79-81 calls this for the eval* / std* gold score
        modes and `compute_crossattention_stats`).  The recording forward runs eagerly (no CUDA graph)."""
        self._capture = True

    def create_crossattention_storage(self):
        self._xattn = []

    def _record_xattn(self, q, kv, B, H, T, Lk, lse, mask, layer=None):
        """Keep ONE record per decoder layer, overwritten by every forward (the reference keeps one buffer per attention
        module, src/fid.py:333-343): constant memory however many forwards run between two `reset_score_storage()` calls."""
        if getattr(self, "_capture", False):
            n = self.config.num_decoder_layers
            if not isinstance(getattr(self, "_xattn", None), list):
                self._xattn = []
            rec = ops.cross_attention_stats(q.detach(), 0, kv.detach(), 0, H * 64, B, H, T, Lk, lse, add_mask=mask, scale=1.0)
            if layer is None:                       # sequential recording: a new forward starts after n records
                if len(self._xattn) >= n:
                    self._xattn = []
                self._xattn.append(rec)
            else:
                if len(self._xattn) != n:
                    self._xattn = [None] * n
                self._xattn[layer] = rec

    @torch.no_grad()
    def get_crossattention_scores(self, n_passages, mask, labels, ids, mode="all", mask_query=None):
        """One scalar per (query, passage) from the recorded cross-attention maps: `FiD.get_crossattention_scores`
        (src/fid.py:136-164) -> dict of [B, n_passages] tensors named {scores,probs,norms}{top5,top10,top20,nosep,first,
        sum,avg,woquery}."""
        rec = getattr(self, "_xattn", None)
        if not rec or len(rec) < self.config.num_decoder_layers or any(r is None for r in rec):
            raise AtlasB200Error("get_crossattention_scores: no recorded forward (call overwrite_forward_crossattention() "
                                 "and run a forward first)")
        rec = rec[-self.config.num_decoder_layers:]
        output = {}
        for idx, prefix in ((0, "scores"), (2, "norms"), (1, "probs")):
            if prefix in mode or "all" in mode:
                self.aggregate_value(torch.stack([r[idx] for r in rec]), mask, labels, n_passages, ids, mask_query, output,
                                     prefix=prefix)
        return output

    def aggregate_value(self, scores, mask, labels, n_passages, ids, mask_query=None, output=None, prefix=""):
        """src/fid.py:166-203, same reductions and the same (hard-coded 256) normalisers."""
        output = {} if output is None else output
        n_layers, bsz, n_tokens, _ = scores.size()
        ids = ids.view(bsz, n_passages, -1)
        scores = scores.view(n_layers, bsz, n_tokens, n_passages, -1)
        mask = mask.view(bsz, n_passages, -1).bool()
        scores = scores.masked_fill(~mask[None, :, None], 0.0)
        valid = ~(labels == -100)
        ntokens_sum = 256 * n_layers * valid.sum(dim=[1])[:, None]
        ntokens_wquery = mask.sum(dim=[2]) * n_layers * valid.sum(dim=[1])[:, None]
        ntokens_first = mask.sum(dim=[2]) * n_layers
        scores = scores.sum(dim=[0])
        for k in (5, 10, 20):
            output[f"{prefix}top{k}"] = self.get_topk_score(k, scores, mask, labels, n_layers)
        scores = scores.masked_fill((labels == -100)[:, :, None, None], 0.0)
        scores_wquery = scores.sum(dim=[1, 3])
        output[f"{prefix}nosep"] = scores.masked_fill(~(ids == 1)[:, None], 0).sum(dim=[1, 3]) / ntokens_sum
        output[f"{prefix}first"] = scores[:, 0].sum(dim=[2]) / ntokens_first
        output[f"{prefix}sum"] = scores_wquery / ntokens_sum
        output[f"{prefix}avg"] = scores_wquery / ntokens_wquery
        if mask_query is not None:
            output[f"{prefix}woquery"] = self.get_woquery_score(scores, mask_query, mask, labels, n_layers)
        return output

    def get_topk_score(self, topk, scores, mask, labels, n_layers):   # src/fid.py:205-210
        top = torch.topk(scores, k=topk, dim=-1)[0].sum(dim=[3])
        top = top.masked_fill((labels == -100)[:, :, None], 0.0)
        ntokens_top = n_layers * (~(labels == -100)).sum(dim=[1])[:, None]
        return top.sum(dim=1) / (topk * ntokens_top)

    def get_woquery_score(self, scores, mask_query, mask, labels, n_layers):   # src/fid.py:212-224
        if scores.size(-1) > mask_query.size(-1):
            pad = torch.zeros([mask_query.size(0), scores.size(-1) - mask_query.size(-1)], device=mask_query.device,
                              dtype=torch.bool)
            mask_query = torch.cat([mask_query, pad], dim=-1)
        mq = mask * (~mask_query[:, None])
        woq = scores.masked_fill(~mq[:, None], 0.0)
        ntokens_woquery = 256 * n_layers * (~(labels == -100)).sum(dim=[1])[:, None]
        return woq.sum(dim=[1, 3]) / ntokens_woquery

    def _shift_right(self, input_ids):  # src/modeling_t5.py:789-813
        start, pad = self.config.decoder_start_token_id, self.config.pad_token_id
        shifted = input_ids.new_zeros(input_ids.shape)
        shifted[..., 1:] = input_ids[..., :-1].clone()
        shifted[..., 0] = start
        shifted.masked_fill_(shifted == -100, pad)
        return shifted

    # ---- weights ---------------------------------------------------------------------------
    def _dtype(self):
        d = self.shared.weight.dtype
        return d if d in (torch.float16, torch.bfloat16) else torch.bfloat16

    def _weights(self):
        dt = self._dtype()
        W = self._half.get(self, dt)
        return W, self._half.derived(dt, self._fuse), dt

    @staticmethod
    def _fuse(W):
        """Weights derived once per parameter version: interleaved wi_0/wi_1, fused q|k|v and cross k|v."""
        g = {}
        for name in list(W.keys()):
            if name.endswith("DenseReluDense.wi_0.weight"):
                w0, w1 = W[name], W[name.replace("wi_0", "wi_1")]
                # rows interleaved (wi_0[j], wi_1[j]) for the gated-GELU GEMM epilogue
                g[name.replace("wi_0.weight", "wi_01")] = torch.stack([w0, w1], dim=1).reshape(-1, w0.shape[1]).contiguous()
            elif name.endswith("SelfAttention.q.weight"):
                # one [3*H*64, d] projection: a single GEMM writes the [tokens, q|k|v] buffer the attention reads
                g[name.replace("q.weight", "qkv")] = torch.cat(
                    [W[name], W[name.replace("q.weight", "k.weight")], W[name.replace("q.weight", "v.weight")]], 0)
            elif name.endswith("EncDecAttention.k.weight"):
                g[name.replace("k.weight", "kv")] = torch.cat([W[name], W[name.replace("k.weight", "v.weight")]], 0)
        # projections with the preceding RMSNorm weight folded in (W'[n, k] = W[n, k] * ln[k]) for the fused norm
        for name in list(g):
            if name.endswith("SelfAttention.qkv"):
                ln = W[name.replace("SelfAttention.qkv", "layer_norm.weight")]
            elif name.endswith("DenseReluDense.wi_01"):
                ln = W[name.replace("DenseReluDense.wi_01", "layer_norm.weight")]
            else:
                continue
            g[name + "_n"] = (g[name].float() * ln.float()[None, :]).to(g[name].dtype).contiguous()
        for name in list(W):        # decoder cross-attention query projection behind layer.1's norm
            if name.endswith("EncDecAttention.q.weight"):
                ln = W[name.replace("EncDecAttention.q.weight", "layer_norm.weight")]
                g[name + "_n"] = (W[name].float() * ln.float()[None, :]).to(W[name].dtype).contiguous()
        return g

    # ---- encoder ---------------------------------------------------------------------------
    def _ff(self, W, G, prefix, h, eps, rows=None):
        n = ops.layernorm(h, W[prefix + "layer_norm.weight"], None, eps, kind=1)
        g = ops.linear(n, G[prefix + "DenseReluDense.wi_01"], epilogue=ops.EPI_GATED, rows=rows)
        # + the fp16 overflow clamp the reference applies after every sub-layer (src/modeling_t5.py:657-708; no-op in bf16)
        return ops.clamp_inf_(ops.linear(g, W[prefix + "DenseReluDense.wo.weight"], None, residual=h,
                                         epilogue=ops.EPI_RESIDUAL, rows=rows))

    @torch.no_grad()
    def encode(self, input_ids, attention_mask):
        """FiDStack encoder (src/fid.py:32-78): [B, n*L] -> each passage encoded independently -> [B, n*L, d]."""
        r = self._encode_rows(input_ids, attention_mask, packed=ops._ENC_PACKED)
        if isinstance(r, tuple):                    # packed rows -> the reference's padded layout (zeros at dropped tiles)
            r = ops.expand_packed_tiles(r[0], r[2])
        return r.view(self.encoder.config.bsz, -1, self.config.d_model)

    @torch.no_grad()
    def _encode_rows(self, input_ids, attention_mask, packed):
        """The encoder stack on a [rows, d] matrix.  packed=False (or a shape the packed path does not take): all S * L padded
        positions, returns the final hidden states [S * L, d].  packed=True: the PADDING-COMPACTED encoder - every passage keeps
        its 64-row tiles up to its last live key, packed back to back (ops.segment_tile_scan); the embedding, all projections
        (device-side row count, static launch shapes), the attention (ops.attention_packed) and the norms only see those rows.
        A padded position never influences a live one (its softmax weight is exactly 0; everything else is row-wise) and the
        cross-attention masks it, so logits, loss and generated tokens are those of the padded computation; the dropped rows
        of the returned encoder states read 0 instead of the reference's never-read values.  Returns (rows [S * L, d] of which
        the first *count are valid, keep uint8 [S, L / 64], tile_off int32 [S * L / 64], count int32 [1])."""
        c = self.config
        W, G, dt = self._weights()
        n_ctx = self.encoder.config.n_context
        ids = input_ids.reshape(input_ids.size(0) * n_ctx, -1)
        mask = attention_mask.reshape(attention_mask.size(0) * n_ctx, -1)
        S, L = ids.shape
        d, H = c.d_model, c.num_heads
        add_mask = (1.0 - mask.to(torch.float32)) * -10000.0                      # 4.18 get_extended_attention_mask
        # 64-key blocks made of padding only (every passage is padded to text_maxlength) weigh exactly 0 in the softmax:
        # the attention kernel skips them (ops.key_block_live), once per forward for all layers
        live = ops.key_block_live(add_mask)
        bias = bias_by_delta(W["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"], L, L, True,
                             c.relative_attention_num_buckets)
        packed = packed and live is not None and L % 64 == 0 and 64 <= L <= 384
        rows = None
        if packed:
            keep, tile_off, tile_src, rows = ops.segment_tile_scan(live)
            h = ops.embed_packed_tiles(ids, W["shared.weight"], tile_src)         # embedding gather of the kept tiles

            def attn(qkv):
                return ops.attention_packed(qkv, keep, tile_off, S, H, L, add_mask, bias, scale=1.0)
        else:
            h = W["shared.weight"][ids.reshape(-1)]                              # embedding gather, [S*L, d]

            def attn(qkv):
                return ops.attention(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, S, H, L, L, add_mask=add_mask, bias_delta=bias,
                                     scale=1.0, block_live=live)
        qkv = torch.empty((S * L, 3 * H * 64), dtype=dt, device=h.device)
        eps = c.layer_norm_epsilon
        if self.fuse_norm:
            # RMSNorm fused around the GEMMs (atlas_b200_linear_ex): the residual GEMMs emit each row's sum of squares,
            # the consuming projection reads the UN-normalised hidden state with the norm weight folded into its matrix
            # and scales its accumulator rows by rsqrt(ss / d + eps).  Only the very first and the final norm run as
            # kernels; 2 x num_layers normalised [tokens, d] tensors are never written or read.
            ss = torch.zeros((2 * c.num_layers, S * L), dtype=torch.float32, device=h.device)
            for i in range(c.num_layers):
                p = f"encoder.block.{i}.layer.0."
                if i == 0:
                    n = ops.layernorm(h, W[p + "layer_norm.weight"], None, eps, kind=1)
                    ops.linear(n, G[p + "SelfAttention.qkv"], out=qkv, rows=rows)
                else:
                    ops.linear(h, G[p + "SelfAttention.qkv_n"], out=qkv, row_ss=ss[2 * i - 1], rs_eps=eps, rows=rows)
                ctx = attn(qkv)
                h = ops.linear(ctx, W[p + "SelfAttention.o.weight"], None, residual=h, epilogue=ops.EPI_RESIDUAL,
                               out_ss=ss[2 * i], rows=rows)
                ops.clamp_inf_(h, row_ss=ss[2 * i])
                p = f"encoder.block.{i}.layer.1."
                g = ops.linear(h, G[p + "DenseReluDense.wi_01_n"], epilogue=ops.EPI_GATED, row_ss=ss[2 * i], rs_eps=eps,
                               rows=rows)
                h = ops.linear(g, W[p + "DenseReluDense.wo.weight"], None, residual=h, epilogue=ops.EPI_RESIDUAL,
                               out_ss=ss[2 * i + 1], rows=rows)
                ops.clamp_inf_(h, row_ss=ss[2 * i + 1])
        else:
            for i in range(c.num_layers):
                p = f"encoder.block.{i}.layer.0."
                n = ops.layernorm(h, W[p + "layer_norm.weight"], None, eps, kind=1)
                ops.linear(n, G[p + "SelfAttention.qkv"], out=qkv, rows=rows)
                ctx = attn(qkv)
                h = ops.clamp_inf_(ops.linear(ctx, W[p + "SelfAttention.o.weight"], None, residual=h,
                                              epilogue=ops.EPI_RESIDUAL, rows=rows))
                h = self._ff(W, G, f"encoder.block.{i}.layer.1.", h, eps, rows=rows)
        h = ops.layernorm(h, W["encoder.final_layer_norm.weight"], None, c.layer_norm_epsilon, kind=1)
        return (h, keep, tile_off, rows) if packed else h

    # ---- decoder ---------------------------------------------------------------------------
    @torch.no_grad()
    def cross_kv(self, enc, out=None):
        """K/V projections of the encoder output for every decoder layer ([B*n*L, 2*H*64] each), computed once
        per forward / per generation (the reference recomputes them every decoding step without use_cache).
        `out`: list of pre-allocated buffers (the static cross K|V cache of the decode graph)."""
        c = self.config
        W, G, dt = self._weights()
        flat = enc.reshape(-1, c.d_model)
        if flat.dtype != dt:
            flat = flat.to(dt)
        return [ops.linear(flat, G[f"decoder.block.{i}.layer.1.EncDecAttention.kv"], out=None if out is None else out[i])
                for i in range(c.num_decoder_layers)]

    @torch.no_grad()
    def decode(self, decoder_input_ids, enc, enc_mask, cross_kv=None, enc_packed=None):
        """Decoder stack + LM head: [B, T] -> logits [B, T, vocab] (src/modeling_t5.py:875-1083,1619-1647).
        enc_packed: the tuple of `_encode_rows(packed=True)` that `enc` was expanded from - the cross K | V projections then
        read the packed rows directly."""
        c = self.config
        W, G, dt = self._weights()
        B, T = decoder_input_ids.shape
        d, H = c.d_model, c.num_heads
        Lk = enc.shape[1]
        # key segments of the split-KV cross-attention: the largest divisor of Lk up to 384 keys (the two-half kernel,
        # attention_split_kernel), else up to 512 (single-accumulator kernel).  256-key segments (K / V double-buffered by
        # that kernel) were measured SLOWER for the FiD decoder: 2.32 vs 1.92 ms per 8-query step - the per-segment
        # softmax / MMA chain, not the load latency, sets the segment time (profiles/r02_launches_step_visit_i.csv)
        split = next((s for s in range(min(384, Lk), 63, -1) if Lk % s == 0), None) or \
            next(s for s in range(min(512, Lk), 0, -1) if Lk % s == 0)
        if split < 64 and Lk > 512:
            raise AtlasB200Error(f"n_context*text_maxlength = {Lk} has no divisor in [64, 512] for the split-KV kernel")
        # invert_attention_mask (4.18): -1e4 for fp16, -1e9 otherwise (src/modeling_t5.py:950)
        neg = -1e4 if dt == torch.float16 else -1e9
        cross_mask = (1.0 - enc_mask.reshape(B, Lk).to(torch.float32)) * neg
        cross_live = ops.key_block_live(cross_mask)           # padded 64-key tiles of the encoder output: skipped, once
        capture = getattr(self, "_capture", False)
        # Only the encoder positions inside a live 64-key tile are ever read by the cross-attention: project K | V for those
        # rows only (compacted, device-side row count: the shapes stay static for the CUDA graph).  At BASELINE configs[3]
        # the 12 K | V projections are 14 % of the step's FLOPs and ~36 % of their rows are padding-only tiles.
        compact = (cross_kv is None and cross_live is not None and not capture and ops._XKV_COMPACT and ops._XATTN_STREAM
                   and T <= 64 and Lk % 64 == 0 and Lk >= 1024)
        xkv = None
        if compact and enc_packed is not None:
            # the encoder already ran on its kept tiles (a superset of the live ones, same tile order): project those rows
            enc_live, keep, tile_off, n_live_rows = enc_packed
            cross_live = keep.view(B, Lk // 64)
            xkv = [ops.linear_dynm(enc_live, G[f"decoder.block.{i}.layer.1.EncDecAttention.kv"], n_live_rows)
                   for i in range(c.num_decoder_layers)]
        elif compact:
            flat = enc.reshape(-1, d)
            if flat.dtype != dt:
                flat = flat.to(dt)
            enc_live, tile_off, n_live_rows = ops.compact_live_tiles(flat, cross_live)
            xkv = [ops.linear_dynm(enc_live, G[f"decoder.block.{i}.layer.1.EncDecAttention.kv"], n_live_rows)
                   for i in range(c.num_decoder_layers)]
        elif cross_kv is None:
            cross_kv = self.cross_kv(enc)
        h = W["shared.weight"][decoder_input_ids.reshape(-1)]
        bias = bias_by_delta(W["decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"], T, T, False,
                             c.relative_attention_num_buckets)
        qkv = torch.empty((B * T, 3 * H * 64), dtype=dt, device=h.device)

        def cross(q, i):
            if xkv is not None:
                return ops.cross_attention_stream_compact(q, xkv[i], cross_live, tile_off, B, H, T, Lk, cross_mask, scale=1.0)
            return ops.cross_attention_split(q, 0, cross_kv[i], 0, H * 64, B, H, T, Lk, add_mask=cross_mask, scale=1.0,
                                             split=split, tile_live=cross_live)

        if self.fuse_norm and not capture:
            # RMSNorm fused around the decoder GEMMs exactly like the encoder's (see encode): the three residual GEMMs of a
            # block emit each row's sum of squares, the consuming projections read the un-normalised rows with the norm weight
            # folded into their matrices.  3 x num_decoder_layers - 1 launches of a ~150-launch dependent chain disappear.
            eps = c.layer_norm_epsilon
            nl = c.num_decoder_layers
            ss = torch.zeros((3 * nl, B * T), dtype=torch.float32, device=h.device)
            for i in range(nl):
                p = f"decoder.block.{i}.layer.0."
                if i == 0:
                    n = ops.layernorm(h, W[p + "layer_norm.weight"], None, eps, kind=1)
                    ops.linear(n, G[p + "SelfAttention.qkv"], out=qkv)
                else:
                    ops.linear(h, G[p + "SelfAttention.qkv_n"], out=qkv, row_ss=ss[3 * i - 1], rs_eps=eps)
                ctx = ops.attention(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, B, H, T, T, bias_delta=bias, scale=1.0,
                                    causal_value=-10000.0)
                h = ops.linear(ctx, W[p + "SelfAttention.o.weight"], None, residual=h, epilogue=ops.EPI_RESIDUAL,
                               out_ss=ss[3 * i])
                ops.clamp_inf_(h, row_ss=ss[3 * i])
                p = f"decoder.block.{i}.layer.1."
                q = ops.linear(h, G[p + "EncDecAttention.q.weight_n"], row_ss=ss[3 * i], rs_eps=eps)
                ctx = cross(q, i)
                h = ops.linear(ctx, W[p + "EncDecAttention.o.weight"], None, residual=h, epilogue=ops.EPI_RESIDUAL,
                               out_ss=ss[3 * i + 1])
                ops.clamp_inf_(h, row_ss=ss[3 * i + 1])
                p = f"decoder.block.{i}.layer.2."
                g = ops.linear(h, G[p + "DenseReluDense.wi_01_n"], epilogue=ops.EPI_GATED, row_ss=ss[3 * i + 1], rs_eps=eps)
                h = ops.linear(g, W[p + "DenseReluDense.wo.weight"], None, residual=h, epilogue=ops.EPI_RESIDUAL,
                               out_ss=ss[3 * i + 2])
                ops.clamp_inf_(h, row_ss=ss[3 * i + 2])
            h = ops.layernorm(h, W["decoder.final_layer_norm.weight"], None, eps, kind=1)
            if getattr(c, "tie_word_embeddings", False):
                h = (h.float() * (d ** -0.5)).to(dt)                              # src/modeling_t5.py:1642-1645
            return ops.linear(h, W["lm_head.weight"]).view(B, T, -1)
        for i in range(c.num_decoder_layers):
            p = f"decoder.block.{i}.layer.0."
            n = ops.layernorm(h, W[p + "layer_norm.weight"], None, c.layer_norm_epsilon, kind=1)
            ops.linear(n, G[p + "SelfAttention.qkv"], out=qkv)
            ctx = ops.attention(qkv, 0, qkv, H * 64, qkv, 2 * H * 64, B, H, T, T, bias_delta=bias, scale=1.0,
                                causal_value=-10000.0)
            h = ops.clamp_inf_(ops.linear(ctx, W[p + "SelfAttention.o.weight"], None, residual=h, epilogue=ops.EPI_RESIDUAL))
            p = f"decoder.block.{i}.layer.1."
            n = ops.layernorm(h, W[p + "layer_norm.weight"], None, c.layer_norm_epsilon, kind=1)
            q = ops.linear(n, W[p + "EncDecAttention.q.weight"])
            if capture:
                ctx, lse = ops.cross_attention_split(q, 0, cross_kv[i], 0, H * 64, B, H, T, Lk, add_mask=cross_mask,
                                                     scale=1.0, split=split, return_lse=True, tile_live=cross_live)
                self._record_xattn(q, cross_kv[i], B, H, T, Lk, lse, cross_mask, layer=i)
            else:
                ctx = cross(q, i)
            h = ops.clamp_inf_(ops.linear(ctx, W[p + "EncDecAttention.o.weight"], None, residual=h,
                                          epilogue=ops.EPI_RESIDUAL))
            h = self._ff(W, G, f"decoder.block.{i}.layer.2.", h, c.layer_norm_epsilon)
        h = ops.layernorm(h, W["decoder.final_layer_norm.weight"], None, c.layer_norm_epsilon, kind=1)
        if getattr(c, "tie_word_embeddings", False):
            h = (h.float() * (d ** -0.5)).to(dt)                                  # src/modeling_t5.py:1642-1645
        logits = ops.linear(h, W["lm_head.weight"])
        return logits.view(B, T, -1)

    # ---- training path (autograd through the kernels, grad_ops.py) ---------------------------------------
    def _train_weights(self):
        """16-bit views of the live parameters built with DIFFERENTIABLE casts / concatenations, so that autograd routes
        the kernels' weight gradients back to the reference-named parameters (q / k / v, wi_0 / wi_1 ...)."""
        dt = self._dtype()
        W = {n: (p if p.dtype == dt else p.to(dt)) for n, p in self.named_parameters()}
        G = {}
        for name in list(W.keys()):
            if name.endswith("DenseReluDense.wi_0.weight"):
                w0, w1 = W[name], W[name.replace("wi_0", "wi_1")]
                G[name.replace("wi_0.weight", "wi_01")] = torch.stack([w0, w1], dim=1).reshape(-1, w0.shape[1])
            elif name.endswith("SelfAttention.q.weight"):
                G[name.replace("q.weight", "qkv")] = torch.cat(
                    [W[name], W[name.replace("q.weight", "k.weight")], W[name.replace("q.weight", "v.weight")]], 0)
            elif name.endswith("EncDecAttention.k.weight"):
                G[name.replace("k.weight", "kv")] = torch.cat([W[name], W[name.replace("k.weight", "v.weight")]], 0)
        return W, G, dt

    def _check_trainable(self):
        return None

    def _dropout_p(self):
        """config.dropout_rate in training mode (nn.Dropout / F.dropout of src/modeling_t5.py:266,286,310,515,561,597,
        964,1059), else 0: masks are generated inside the kernels and re-derived in the backward (csrc/dropout.cuh)."""
        return float(getattr(self.config, "dropout_rate", 0.0) or 0.0) if self.training else 0.0

    def _maybe_ckpt(self, fn, *args):
        if getattr(self, "_grad_ckpt", False):
            from torch.utils.checkpoint import checkpoint

            return checkpoint(fn, *args, use_reentrant=False)
        return fn(*args)

    def _encode_train(self, WG, input_ids, attention_mask):
        """`encode` with autograd (same kernels forward, un-fused norms; backward = grad_ops)."""
        g = grad_ops
        c = self.config
        W, G, dt = WG
        n_ctx, bsz = self.encoder.config.n_context, self.encoder.config.bsz
        ids = input_ids.reshape(input_ids.size(0) * n_ctx, -1)
        mask = attention_mask.reshape(attention_mask.size(0) * n_ctx, -1)
        S, L = ids.shape
        d, H = c.d_model, c.num_heads
        eps = c.layer_norm_epsilon
        pdrop = self._dropout_p()
        h = g.dropout(g.embedding(W["shared.weight"], ids), pdrop)
        add_mask = (1.0 - mask.to(torch.float32)) * -10000.0
        add_mask._atlas_block_live = ops.key_block_live(add_mask)     # read by grad_ops._SelfAttention (all layers share it)

        def lin_res(x, w, res):
            # h + dropout(linear(x)) (T5LayerSelfAttention / T5LayerFF); without dropout the add is the GEMM's epilogue;
            # then the fp16 overflow clamp of the block (a no-op for bf16)
            if pdrop:
                return g.clamp_inf(g.dropout(g.linear(x, w), pdrop, residual=res))
            return g.clamp_inf(g.linear(x, w, None, residual=res))

        bias = bias_by_delta(W["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"], L, L, True,
                             c.relative_attention_num_buckets)

        def block(i, h, bias):
            p = f"encoder.block.{i}.layer.0."
            n = g.layernorm(h, W[p + "layer_norm.weight"], None, eps, kind=1)
            qkv = g.linear(n, G[p + "SelfAttention.qkv"])
            ctx = g.self_attention(qkv, S, H, L, add_mask=add_mask, bias_delta=bias, scale=1.0, dropout_p=pdrop)
            h = lin_res(ctx, W[p + "SelfAttention.o.weight"], h)
            p = f"encoder.block.{i}.layer.1."
            n = g.layernorm(h, W[p + "layer_norm.weight"], None, eps, kind=1)
            u = g.linear(n, G[p + "DenseReluDense.wi_01"])
            return lin_res(g.dropout(g.gated_gelu(u), pdrop), W[p + "DenseReluDense.wo.weight"], h)

        for i in range(c.num_layers):
            h = self._maybe_ckpt(block, i, h, bias)
        h = g.dropout(g.layernorm(h, W["encoder.final_layer_norm.weight"], None, eps, kind=1), pdrop)
        return h.view(bsz, -1, d)

    def _decode_train(self, WG, decoder_input_ids, enc, enc_mask):
        g = grad_ops
        c = self.config
        W, G, dt = WG
        B, T = decoder_input_ids.shape
        d, H = c.d_model, c.num_heads
        Lk = enc.shape[1]
        eps = c.layer_norm_epsilon
        pdrop = self._dropout_p()
        split = None
        if pdrop:    # the dropout mask is addressed in 32-key groups over the whole key range (csrc/dropout.cuh)
            split = next((s for s in range(min(512, Lk) // 32 * 32, 63, -32) if Lk % s == 0), None)
        split = split or next((s for s in range(min(384, Lk), 63, -1) if Lk % s == 0), None) or \
            next(s for s in range(min(512, Lk), 0, -1) if Lk % s == 0)
        if split < 64 and Lk > 512:
            raise AtlasB200Error(f"n_context*text_maxlength = {Lk} has no divisor in [64, 512] for the split-KV kernel")
        flat = enc.reshape(-1, d)
        if flat.dtype != dt:
            flat = flat.to(dt)
        if pdrop and split % 32 != 0 and split != Lk:
            raise AtlasB200Error(f"dropout on the cross-attention probabilities needs a key split that is a multiple of 32 "
                                 f"(n_context*text_maxlength = {Lk} gave {split})")
        h = g.dropout(g.embedding(W["shared.weight"], decoder_input_ids), pdrop)

        def lin_res(x, w, res):
            if pdrop:
                return g.clamp_inf(g.dropout(g.linear(x, w), pdrop, residual=res))
            return g.clamp_inf(g.linear(x, w, None, residual=res))

        bias = bias_by_delta(W["decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"], T, T, False,
                             c.relative_attention_num_buckets)
        neg = -1e4 if dt == torch.float16 else -1e9
        cross_mask = (1.0 - enc_mask.reshape(B, Lk).to(torch.float32)) * neg
        cross_mask._atlas_block_live = ops.key_block_live(cross_mask)     # padded 64-key tiles: skipped forward and backward

        def block(i, h, bias, flat):
            p = f"decoder.block.{i}.layer.0."
            n = g.layernorm(h, W[p + "layer_norm.weight"], None, eps, kind=1)
            qkv = g.linear(n, G[p + "SelfAttention.qkv"])
            ctx = g.self_attention(qkv, B, H, T, bias_delta=bias, scale=1.0, causal_value=-10000.0, dropout_p=pdrop)
            h = lin_res(ctx, W[p + "SelfAttention.o.weight"], h)
            p = f"decoder.block.{i}.layer.1."
            n = g.layernorm(h, W[p + "layer_norm.weight"], None, eps, kind=1)
            q = g.linear(n, W[p + "EncDecAttention.q.weight"])
            kv = g.linear(flat, G[p + "EncDecAttention.kv"])
            ctx, lse = g.cross_attention(q, kv, B, H, T, Lk, add_mask=cross_mask, scale=1.0, split=split, return_lse=True,
                                         dropout_p=pdrop)
            self._record_xattn(q, kv, B, H, T, Lk, lse, cross_mask, layer=i)
            h = lin_res(ctx, W[p + "EncDecAttention.o.weight"], h)
            p = f"decoder.block.{i}.layer.2."
            n = g.layernorm(h, W[p + "layer_norm.weight"], None, eps, kind=1)
            u = g.linear(n, G[p + "DenseReluDense.wi_01"])
            return lin_res(g.dropout(g.gated_gelu(u), pdrop), W[p + "DenseReluDense.wo.weight"], h)

        for i in range(c.num_decoder_layers):
            h = self._maybe_ckpt(block, i, h, bias, flat)
        h = g.dropout(g.layernorm(h, W["decoder.final_layer_norm.weight"], None, eps, kind=1), pdrop)
        if getattr(c, "tie_word_embeddings", False):
            raise AtlasB200Error("tied LM head (T5 v1.0) is not supported on the training path (T5 v1.1 is untied)")
        logits = g.linear(h, W["lm_head.weight"])
        return logits.view(B, T, -1)

    def _forward_train(self, input_ids, attention_mask, decoder_input_ids, labels, encoder_outputs):
        self._check_trainable()
        WG = self._train_weights()
        if encoder_outputs is not None:
            enc = encoder_outputs[0]
        else:
            enc = self._encode_train(WG, input_ids, attention_mask.to(torch.bool))
        B = enc.shape[0]
        logits = self._decode_train(WG, decoder_input_ids, enc, attention_mask.reshape(B, -1).to(torch.bool))
        loss = None
        pd = self.shared.weight.dtype
        if labels is not None:
            loss = grad_ops.cross_entropy(logits, labels).to(pd)      # CrossEntropyLoss(ignore_index=-100)
        if logits.dtype != pd:
            logits, enc = logits.to(pd), enc.to(pd)
        return FiDOutput(loss, logits, enc)

    # ---- CUDA-graph cache --------------------------------------------------------------------
    def _run(self, tag, fn, inputs):
        """Run `fn(*inputs)` (static shapes, device tensors in / out): replay a captured graph when enabled."""
        if not self.cuda_graphs or getattr(self, "_capture", False):   # recording forwards append to a python list
            return fn(*inputs)
        self._weights()                                      # make sure the 16-bit weight copies exist / are current
        dt = self._dtype()
        # the 16-bit weight buffers are refreshed in place (HalfCache): the graph key only holds their allocation generation
        key = (tag, dt, self.fuse_norm, ops._ENC_PACKED, self._half.sets[dt]["gen"],
               tuple((tuple(t.shape), t.dtype, t.device) for t in inputs))
        runner = self._graphs.get(key)
        if runner is None:
            if len(self._graphs) >= 8:                       # shapes changed often: drop the oldest graph + its pool
                self._graphs.pop(next(iter(self._graphs)))
            runner = _GraphRunner(fn, inputs)
            self._graphs[key] = runner
        out = runner(*inputs)
        # results live in the graph's static buffers: hand out copies
        return out.clone() if torch.is_tensor(out) else tuple(o.clone() for o in out)

    # ---- public forward / generate -----------------------------------------------------------
    def forward(self, input_ids=None, attention_mask=None, decoder_input_ids=None, labels=None, encoder_outputs=None,
                use_cache=False, **unused):
        if decoder_input_ids is None and labels is not None:
            decoder_input_ids = self._shift_right(labels)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training step (train.py): the same kernels forward, autograd through grad_ops.py backward
            return self._forward_train(input_ids, attention_mask, decoder_input_ids, labels, encoder_outputs)
        if encoder_outputs is not None:
            enc = encoder_outputs[0]
            B = enc.shape[0]
            mask2 = attention_mask.reshape(B, -1)
            logits = self._run("dec", lambda e, m, d: self.decode(d, e, m), (enc, mask2.to(torch.bool), decoder_input_ids))
        else:
            n_ctx, bsz = self.encoder.config.n_context, self.encoder.config.bsz

            def full(ids, mask, dec):
                r = self._encode_rows(ids, mask, packed=ops._ENC_PACKED)
                packed = r if isinstance(r, tuple) else None
                e = (ops.expand_packed_tiles(r[0], r[2]) if packed else r).view(bsz, -1, self.config.d_model)
                return self.decode(dec, e, mask.reshape(e.shape[0], -1), enc_packed=packed), e

            logits, enc = self._run(("full", n_ctx, bsz), full,
                                    (input_ids, attention_mask.to(torch.bool), decoder_input_ids))
        loss = None
        if labels is not None:
            # CrossEntropyLoss(ignore_index=-100), src/modeling_t5.py:1650-1652
            loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.size(-1)).float(), labels.reshape(-1),
                                                     ignore_index=-100).to(logits.dtype)
        pd = self.shared.weight.dtype
        if logits.dtype != pd:
            logits, enc = logits.to(pd), enc.to(pd)
        return FiDOutput(loss, logits, enc)

    # ---- greedy generation: KV-cached single-token decode (csrc/decode.cu) ----------------------------------------
    class _DecodeState:
        """Static buffers of one (batch, keys, max_length, dtype) generation shape: per-layer self K|V caches, the cross
        K|V cache, the token / sequence / done / step tensors the step reads and writes on the device, and (optionally)
        the captured CUDA graph of ONE decode step that every step replays (the step index lives in device memory)."""

        def __init__(self, model, B, Lk, Tmax, dt, dev):
            c = model.config
            H, d = c.num_heads, c.d_model
            nl = c.num_decoder_layers
            self.B, self.Lk, self.Tmax = B, Lk, Tmax
            self.self_kv = [torch.zeros((B, Tmax, 2 * H * 64), dtype=dt, device=dev) for _ in range(nl)]
            self.cross_kv = [torch.empty((B * Lk, 2 * H * 64), dtype=dt, device=dev) for _ in range(nl)]
            self.cross_mask = torch.zeros((B, Lk), dtype=torch.float32, device=dev)
            self.tok_in = torch.zeros(B, dtype=torch.int64, device=dev)
            self.seq = torch.zeros((B, Tmax), dtype=torch.int64, device=dev)
            self.done = torch.zeros(B, dtype=torch.uint8, device=dev)
            self.t_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self.logits = None
            self.graph = None
            self.cross_live = None

    @torch.no_grad()
    def _decode_step_logits(self, st):
        """Decoder stack for ONE new token per sequence (`st.tok_in`, position `st.t_dev`) -> logits [B, vocab].
        Self-attention appends to / reads the per-layer K|V cache; cross-attention streams the cached cross K|V once."""
        c = self.config
        W, G, dt = self._weights()
        H = c.num_heads
        eps = c.layer_norm_epsilon
        B, Lk = st.B, st.Lk
        h = W["shared.weight"][st.tok_in]                                          # [B, d]
        for i in range(c.num_decoder_layers):
            p = f"decoder.block.{i}.layer.0."
            n = ops.layernorm(h, W[p + "layer_norm.weight"], None, eps, kind=1)
            qkv = ops.linear(n, G[p + "SelfAttention.qkv"])
            ctx = ops.decode_self_attention(qkv, st.self_kv[i], st.t_dev, H, bias_delta=st.bias, scale=1.0)
            h = ops.clamp_inf_(ops.linear(ctx, W[p + "SelfAttention.o.weight"], None, residual=h, epilogue=ops.EPI_RESIDUAL))
            p = f"decoder.block.{i}.layer.1."
            n = ops.layernorm(h, W[p + "layer_norm.weight"], None, eps, kind=1)
            q = ops.linear(n, W[p + "EncDecAttention.q.weight"])
            ctx = ops.decode_cross_attention(q, st.cross_kv[i], B, H, Lk, add_mask=st.cross_mask, scale=1.0,
                                             chunk=st.chunk, tile_live=st.cross_live)
            h = ops.clamp_inf_(ops.linear(ctx, W[p + "EncDecAttention.o.weight"], None, residual=h,
                                          epilogue=ops.EPI_RESIDUAL))
            h = self._ff(W, G, f"decoder.block.{i}.layer.2.", h, eps)
        h = ops.layernorm(h, W["decoder.final_layer_norm.weight"], None, eps, kind=1)
        if getattr(c, "tie_word_embeddings", False):
            h = (h.float() * (c.d_model ** -0.5)).to(dt)
        return ops.linear(h, W["lm_head.weight"])

    def _decode_state(self, B, Lk, Tmax, dt, dev):
        key = (B, Lk, Tmax, dt, dev)
        states = self.__dict__.setdefault("_decode_states", {})
        st = states.get(key)
        if st is None:
            if len(states) >= 2:
                states.pop(next(iter(states)))
            st = FiD._DecodeState(self, B, Lk, Tmax, dt, dev)
            # keys per block of the cross-attention sweep: ~4 blocks per SM keep every SM streaming
            st.chunk = 256 if Lk >= 2048 else 64
            states[key] = st
        return st

    @torch.no_grad()
    def generate(self, input_ids=None, attention_mask=None, max_length=32, min_length=1, num_beams=1,
                 num_return_sequences=1, length_penalty=1.0, forced_bos_token_id=None, prefix_allowed_tokens_fn=None,
                 use_cache=True, **unused):
        """Greedy decoding (what `Atlas.generate` requests with num_beams=1, src/atlas.py:592-619; transformers 4.18
        `greedy_search` with `use_cache`): the encoder and the cross-attention K|V projections run once; every step
        decodes ONE token per sequence against the per-layer self K|V cache and the cached cross K|V (csrc/decode.cu),
        picks the next token on the device and - without a `prefix_allowed_tokens_fn` - is replayed from a single CUDA
        graph with no host synchronisation except an all-done poll every 8 steps.  use_cache=False runs the
        first-generation path (the decoder prefix re-run every step), kept as the parity reference of the tests."""
        if num_beams != 1 or num_return_sequences != 1:
            raise AtlasB200Error("only greedy generation (num_beams=1) is implemented")
        if not use_cache:
            return self._generate_prefix_rerun(input_ids, attention_mask, max_length, min_length, prefix_allowed_tokens_fn)
        c = self.config
        enc = self.encode(input_ids, attention_mask)
        B, Lk = enc.shape[0], enc.shape[1]
        W, G, dt = self._weights()
        st = self._decode_state(B, Lk, int(max_length), dt, enc.device)
        self.cross_kv(enc, out=st.cross_kv)
        neg = -1e4 if dt == torch.float16 else -1e9                                 # invert_attention_mask (4.18)
        st.cross_mask.copy_((1.0 - attention_mask.reshape(B, Lk).to(torch.float32)) * neg)
        live = ops.key_block_live(st.cross_mask)              # padded 64-key tiles: not read by the decode steps
        if live is not None:
            if st.cross_live is None:
                st.cross_live = live                          # fixed buffer: the captured step graph reads it
            else:
                st.cross_live.copy_(live)
        wkey = (self._half.sets[dt]["gen"], self._half.sets[dt]["key"])
        if getattr(st, "wkey", None) != wkey:                                       # weights changed: the bias table is stale
            fresh = bias_by_delta(W["decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"], st.Tmax,
                                  st.Tmax, False, c.relative_attention_num_buckets)
            if getattr(st, "bias", None) is None:
                st.bias = fresh
            else:
                st.bias.copy_(fresh)            # in place: the captured step graph reads this buffer
            if getattr(st, "wgen", None) != wkey[0]:                                # buffers re-allocated: re-capture
                st.graph, st.wgen = None, wkey[0]
            st.wkey = wkey
        st.seq.fill_(c.pad_token_id)
        st.seq[:, 0] = c.decoder_start_token_id
        st.tok_in.fill_(c.decoder_start_token_id)
        st.done.zero_()
        st.t_dev.zero_()
        n_steps = int(max_length) - 1

        def step():
            logits = self._decode_step_logits(st)
            ops.decode_argmax(logits, st.seq, st.tok_in, st.done, st.t_dev, c.eos_token_id, c.pad_token_id, int(min_length))

        use_graph = self.cuda_graphs and prefix_allowed_tokens_fn is None and n_steps > 0
        if use_graph and (st.graph is None or getattr(st, "min_length", None) != int(min_length)):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                    # eager warm-up step on scratch state (then reset below)
                step()
            torch.cuda.current_stream().wait_stream(side)
            st.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(st.graph):
                step()
            st.min_length = int(min_length)
            st.seq.fill_(c.pad_token_id)
            st.seq[:, 0] = c.decoder_start_token_id
            st.tok_in.fill_(c.decoder_start_token_id)
            st.done.zero_()
            st.t_dev.zero_()
        ran = 0
        for s_i in range(n_steps):
            if use_graph:
                st.graph.replay()
            elif prefix_allowed_tokens_fn is None:
                step()
            else:
                logits = self._decode_step_logits(st).float()
                seq_now = st.seq[:, : s_i + 1]
                allowed = torch.full_like(logits, -float("inf"))
                for b in range(B):
                    allowed[b, prefix_allowed_tokens_fn(b, seq_now[b])] = 0
                logits = (logits + allowed).to(dt)
                ops.decode_argmax(logits, st.seq, st.tok_in, st.done, st.t_dev, c.eos_token_id, c.pad_token_id,
                                  int(min_length))
            ran = s_i + 1
            if (s_i % 8 == 7 or s_i == n_steps - 1) and bool(st.done.all()):      # the only host synchronisation
                break
        seq = st.seq[:, : ran + 1].clone()
        # transformers stops right after the step in which the last sequence finished: drop all-pad trailing columns
        is_eos = seq[:, 1:] == c.eos_token_id
        if bool(is_eos.any(dim=1).all()):
            first = is_eos.float().argmax(dim=1) + 1
            seq = seq[:, : int(first.max()) + 1]
        return seq

    @torch.no_grad()
    def _generate_prefix_rerun(self, input_ids, attention_mask, max_length, min_length, prefix_allowed_tokens_fn):
        """First-generation greedy loop: the encoder and the cross-attention K/V projections run once, every step
        re-runs the whole decoder prefix (no self-attention cache).  Parity reference for the cached path."""
        c = self.config
        enc = self.encode(input_ids, attention_mask)
        B = enc.shape[0]
        kv = self.cross_kv(enc)
        mask = attention_mask.reshape(B, -1)
        seq = torch.full((B, 1), c.decoder_start_token_id, dtype=torch.long, device=enc.device)
        done = torch.zeros(B, dtype=torch.bool, device=enc.device)
        for step in range(max_length - 1):
            logits = self.decode(seq, enc, mask, cross_kv=kv)[:, -1].float()
            if step + 1 < min_length:
                logits[:, c.eos_token_id] = -float("inf")
            if prefix_allowed_tokens_fn is not None:
                allowed = torch.full_like(logits, -float("inf"))
                for b in range(B):
                    allowed[b, prefix_allowed_tokens_fn(b, seq[b])] = 0
                logits = logits + allowed
            nxt = logits.argmax(dim=-1)
            nxt = torch.where(done, torch.full_like(nxt, c.pad_token_id), nxt)
            seq = torch.cat([seq, nxt[:, None]], dim=1)
            done = done | (nxt == c.eos_token_id)
            if bool(done.all()):
                break
        return seq
