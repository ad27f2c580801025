"""Thin torch-tensor wrappers over the C ABI (`include/atlas_b200.h`).

torch owns the device memory and the stream; the arithmetic happens in `lib/libatlas_b200.so`.
Nothing here falls back to torch ops: a missing library or a CPU tensor raises.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import AtlasB200Error, check, current_stream_ptr, lib, require_cuda

DIM = _lib.EMBEDDINGS_DIM


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class Workspace:
    """Grow-only device scratch buffer reused across calls (kernels borrow it per launch)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return self.buf


_default_ws = Workspace()
_wgrad_ws = Workspace()


def _check_bank(bank):
    require_cuda(bank, "bank")
    if bank.dim() != 2 or bank.shape[1] != DIM or bank.stride(1) != 1:
        raise AtlasB200Error(f"bank must be [n, {DIM}] with unit inner stride, got {tuple(bank.shape)} {bank.stride()}")
    if bank.dtype not in (torch.float16, torch.bfloat16):
        raise AtlasB200Error("bank must be fp16 (reference dtype, src/index.py:51) or bf16")
    return 1 if bank.dtype == torch.bfloat16 else 0


def cast_queries(queries, dtype):
    """`allqueries.half()` (src/index.py:117) on the device through the library's cast kernel."""
    require_cuda(queries, "queries")
    if queries.dtype == dtype:
        return queries.contiguous()
    q32 = queries.float().contiguous()
    out = torch.empty(q32.shape, dtype=dtype, device=q32.device)
    check(lib().atlas_b200_cast_f32(_ptr(q32), _ptr(out), q32.numel(), 1 if dtype == torch.bfloat16 else 0,
                                    current_stream_ptr()))
    return out


def mips_topk(bank, queries, k, id_base=0, id_stride=1, workspace=None, exhaustive=False):
    """Exact top-k inner-product search of `queries` [nq, 768] in `bank` [n, 768] (one shard).

    Returns (scores [nq,k] bank dtype desc, ids [nq,k] int64, status int32[1] device tensor).
    `status != 0` means the fast path overflowed and the caller must retry with exhaustive=True
    (see `search_shard`, which does that)."""
    is_bf16 = _check_bank(bank)
    n = bank.shape[0]
    q = cast_queries(queries.reshape(-1, DIM), bank.dtype)
    nq = q.shape[0]
    ws = workspace or _default_ws
    nbytes = lib().atlas_b200_mips_workspace_bytes(n, nq, k)
    buf = ws.get(nbytes, bank.device)
    out_s = torch.empty((nq, k), dtype=bank.dtype, device=bank.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=bank.device)
    status = torch.zeros(1, dtype=torch.int32, device=bank.device)
    if exhaustive:
        check(lib().atlas_b200_mips_topk_exhaustive(
            _ptr(bank), n, bank.stride(0), is_bf16, _ptr(q), nq, k, _ptr(out_s), _ptr(out_i), id_base, id_stride,
            _ptr(buf), buf.numel(), current_stream_ptr()))
    else:
        check(lib().atlas_b200_mips_topk(
            _ptr(bank), n, bank.stride(0), is_bf16, _ptr(q), nq, k, _ptr(out_s), _ptr(out_i), id_base, id_stride,
            _ptr(status), _ptr(buf), buf.numel(), current_stream_ptr()))
    return out_s, out_i, status


def search_shard(bank, queries, k, id_base=0, id_stride=1, workspace=None):
    """mips_topk + the overflow fallback (one host sync to read the status word)."""
    s, i, status = mips_topk(bank, queries, k, id_base, id_stride, workspace)
    if queries.numel() and int(status.item()) != 0:
        s, i, _ = mips_topk(bank, queries, k, id_base, id_stride, workspace, exhaustive=True)
    return s, i


def topk_merge(scores_in, ids_in, world, nq_total, k, q_begin, nq_out, stride_s=None, stride_i=None):
    """Merge W per-shard lists [W, nq_total, k] -> rows [q_begin, q_begin+nq_out) merged [nq_out, k]."""
    require_cuda(scores_in, "scores_in")
    out_s = torch.empty((nq_out, k), dtype=scores_in.dtype, device=scores_in.device)
    out_i = torch.empty((nq_out, k), dtype=torch.int64, device=scores_in.device)
    ss = nq_total * k if stride_s is None else stride_s
    si = nq_total * k if stride_i is None else stride_i
    check(lib().atlas_b200_topk_merge(_ptr(scores_in), _ptr(ids_in), ss, si,
                                      1 if scores_in.dtype == torch.bfloat16 else 0, world, nq_total, k, q_begin,
                                      nq_out, _ptr(out_s), _ptr(out_i), current_stream_ptr()))
    return out_s, out_i


def search_host(bank, queries_host, k, id_base=0, id_stride=1, workspace=None):
    """The C-ABI end-to-end call with HOST buffers (H2D + cast + scan + select + D2H, synchronous).

    queries_host: CPU float32 tensor [nq, 768] (pinned or pageable).  Returns CPU tensors
    (scores float32 [nq,k] holding the 16-bit values, ids int64 [nq,k])."""
    is_bf16 = _check_bank(bank)
    if queries_host.is_cuda or queries_host.dtype != torch.float32:
        raise AtlasB200Error("search_host takes a CPU float32 query tensor")
    qh = queries_host.reshape(-1, DIM).contiguous()
    nq = qh.shape[0]
    n = bank.shape[0]
    ws = workspace or _default_ws
    nbytes = lib().atlas_b200_mips_workspace_bytes(n, nq, k)
    buf = ws.get(nbytes, bank.device)
    out_s = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    out_i = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    check(lib().atlas_b200_search_host(_ptr(bank), n, bank.stride(0), is_bf16, _ptr(qh), nq, k, _ptr(out_s),
                                       _ptr(out_i), id_base, id_stride, _ptr(buf), buf.numel(),
                                       current_stream_ptr()))
    return out_s, out_i


def pack_results(scores, ids, status=None):
    """One contiguous byte blob [scores (16-bit) | pad to 8 | ids (int64) | status (int64)] so that a SINGLE all-gather
    ships both lists AND the shard's overflow flag (replaces the 2*W gathers of src/index.py:138-141; the flag lets every
    rank learn, at the host sync it has anyway, whether any shard must be re-run exhaustively).
    Returns (blob uint8, ids offset, status offset)."""
    nbytes_s = scores.numel() * 2
    ids_off = (nbytes_s + 7) // 8 * 8
    st_off = ids_off + ids.numel() * 8
    blob = torch.empty(st_off + 8, dtype=torch.uint8, device=scores.device)
    blob[:nbytes_s].view(scores.dtype).copy_(scores.reshape(-1))
    blob[ids_off:st_off].view(torch.int64).copy_(ids.reshape(-1))
    if status is None:
        blob[st_off:].zero_()
    else:
        blob[st_off:].view(torch.int64).copy_(status.reshape(1))
    return blob, ids_off, st_off


def topk_merge_blob(blob_all, ids_off, world, nq_total, k, q_begin, nq_out, dtype):
    """Merge directly out of the all-gathered blobs `[W, blob_bytes]` (see pack_results)."""
    require_cuda(blob_all, "blob_all")
    blob_bytes = blob_all.shape[1]
    assert blob_bytes % 8 == 0 and blob_all.is_contiguous()
    out_s = torch.empty((nq_out, k), dtype=dtype, device=blob_all.device)
    out_i = torch.empty((nq_out, k), dtype=torch.int64, device=blob_all.device)
    base = blob_all.data_ptr()
    check(lib().atlas_b200_topk_merge(ctypes.c_void_p(base), ctypes.c_void_p(base + ids_off), blob_bytes // 2,
                                      blob_bytes // 8, 1 if dtype == torch.bfloat16 else 0, world, nq_total, k,
                                      q_begin, nq_out, _ptr(out_s), _ptr(out_i), current_stream_ptr()))
    return out_s, out_i


# ---------------------------------------------------------------------------------------------
# dense layers (csrc/gemm.cu)
# ---------------------------------------------------------------------------------------------
EPI_NONE, EPI_BIAS, EPI_GELU, EPI_RESIDUAL, EPI_GATED = 0, 1, 2, 3, 4


def linear(x, weight, bias=None, residual=None, epilogue=None, out=None, row_ss=None, out_ss=None, rs_eps=1e-6, rows=None):
    """y = epilogue(x @ weight.T) on tcgen05.  x [..., K], weight [N, K] (nn.Linear layout), 16-bit.

    epilogue: EPI_NONE / EPI_BIAS / EPI_GELU / EPI_RESIDUAL / EPI_GATED (see include/atlas_b200.h);
    default: EPI_RESIDUAL if `residual` is given, else EPI_BIAS if `bias` is given, else EPI_NONE.
    Fused T5 RMSNorm (atlas_b200_linear_ex): `row_ss` fp32 [M] = sum of squares of the rows of x -> accumulator rows are
    scaled by rsqrt(row_ss / K + rs_eps) (x is the UN-normalised hidden state, the norm weight is folded into `weight`);
    `out_ss` fp32 [M] (pre-zeroed) receives the sum of squares of the stored output rows.
    `rows`: int32 [1] device tensor - only the first *rows rows are computed (atlas_b200_linear_rows; static launch shape)."""
    require_cuda(x, "x")
    if x.dtype not in (torch.float16, torch.bfloat16) or weight.dtype != x.dtype:
        raise AtlasB200Error(f"linear: x and weight must both be fp16 or bf16 (got {x.dtype}, {weight.dtype})")
    K = x.shape[-1]
    N = weight.shape[0]
    if weight.shape[1] != K:
        raise AtlasB200Error(f"linear: weight {tuple(weight.shape)} does not match K={K}")
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    w = weight if weight.stride(-1) == 1 else weight.contiguous()
    M = x2.shape[0]
    if epilogue is None:
        epilogue = EPI_RESIDUAL if residual is not None else (EPI_BIAS if bias is not None else EPI_NONE)
    n_out = N // 2 if epilogue == EPI_GATED else N
    if out is None:
        out = torch.empty((M, n_out), dtype=x.dtype, device=x.device)
    r2 = None
    if residual is not None:
        r2 = residual.reshape(-1, residual.shape[-1])
        if r2.stride(-1) != 1:
            r2 = r2.contiguous()
    for t in (row_ss, out_ss):
        if t is not None and (t.dtype != torch.float32 or t.numel() != M or not t.is_contiguous() or not t.is_cuda):
            raise AtlasB200Error("linear: row_ss / out_ss must be contiguous CUDA fp32 tensors with one element per row")
    check(lib().atlas_b200_linear_rows(
        _ptr(x2), x2.stride(0), _ptr(w), w.stride(0), _ptr(bias) if bias is not None else None,
        _ptr(r2) if r2 is not None else None, r2.stride(0) if r2 is not None else 0, _ptr(out), out.stride(0),
        M, N, K, epilogue, 1 if x.dtype == torch.bfloat16 else 0, _ptr(row_ss) if row_ss is not None else None,
        _ptr(out_ss) if out_ss is not None else None, float(rs_eps), _ptr(rows) if rows is not None else None,
        current_stream_ptr()))
    return out.reshape(*x.shape[:-1], n_out)


# ---------------------------------------------------------------------------------------------
# normalisation / embedding / pooling (csrc/elementwise.cu) and attention (csrc/attention.cu)
# ---------------------------------------------------------------------------------------------
def _bf(t):
    return 1 if t.dtype == torch.bfloat16 else 0


def layernorm(x, weight, bias=None, eps=1e-12, kind=0, out=None):
    """kind 0: BertLayerNorm (uncentred 2nd moment, src/modeling_bert.py:104-114); kind 1: T5 RMSNorm."""
    require_cuda(x, "x")
    H = x.shape[-1]
    x2 = x.reshape(-1, H)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    if out is None:
        out = torch.empty_like(x2)
    check(lib().atlas_b200_layernorm(_ptr(x2), x2.stride(0), _ptr(weight), _ptr(bias) if bias is not None else None,
                                     _ptr(out), out.stride(0), x2.shape[0], H, float(eps), kind, _bf(x),
                                     current_stream_ptr()))
    return out.reshape(x.shape)


def bert_embed_ln(input_ids, token_type_ids, word_emb, type_emb, pos_emb, ln_weight, ln_bias, eps):
    require_cuda(input_ids, "input_ids")
    B, L = input_ids.shape
    H = word_emb.shape[1]
    y = torch.empty((B, L, H), dtype=word_emb.dtype, device=word_emb.device)
    ids = input_ids.contiguous()
    tt = token_type_ids.contiguous() if token_type_ids is not None else None
    check(lib().atlas_b200_bert_embed_ln(_ptr(ids), _ptr(tt) if tt is not None else None, _ptr(word_emb), _ptr(type_emb),
                                         _ptr(pos_emb), _ptr(ln_weight), _ptr(ln_bias), _ptr(y), B, L, H, float(eps),
                                         _bf(word_emb), current_stream_ptr()))
    return y


def masked_mean_pool(x, mask, out=None):
    """x [B, L, H] 16-bit, mask [B, L] int64 -> [B, H] (optionally written into `out` rows, e.g. the bank)."""
    require_cuda(x, "x")
    B, L, H = x.shape
    xc = x.contiguous()
    m = mask.to(torch.int64).contiguous()
    if out is None:
        out = torch.empty((B, H), dtype=x.dtype, device=x.device)
    check(lib().atlas_b200_masked_mean_pool(_ptr(xc), _ptr(m), _ptr(out), out.stride(0), B, L, H, _bf(x),
                                            current_stream_ptr()))
    return out


def attention(q, q_col0, k, k_col0, v, v_col0, B, H, Lq, Lk, add_mask=None, bias_delta=None, scale=1.0,
              causal_value=0.0, out=None, return_lse=False, dropout=None, block_live=None):
    """Fused attention reading Q/K/V in place from [B*L, ld] projection buffers (head h at col0 + 64h).
    return_lse: also return the row log-sum-exp [B, H, Lq] fp32 (saved for the backward pass).
    block_live = key_block_live(add_mask): fully masked 64-key blocks are skipped by the encoder kernel (identical results).
    dropout = (p, seed, offset): dropout on the probabilities (training), mask re-derived by attention_bwd."""
    require_cuda(q, "q")
    if out is None:
        out = torch.empty((B * Lq, H * 64), dtype=q.dtype, device=q.device)
    am = add_mask.float().contiguous() if add_mask is not None else None
    bd = bias_delta.float().contiguous() if bias_delta is not None else None
    lse = torch.empty((B, H, Lq), dtype=torch.float32, device=q.device) if return_lse else None
    dp, dseed, doff = dropout if dropout is not None else (0.0, 0, 0)
    check(lib().atlas_b200_attention_train(_ptr(q), q.stride(0), q_col0, _ptr(k), k.stride(0), k_col0, _ptr(v), v.stride(0),
                                           v_col0, _ptr(out), out.stride(0), _ptr(am) if am is not None else None,
                                           _ptr(bd) if bd is not None else None, B, H, Lq, Lk, float(scale),
                                           float(causal_value), 1, None, None, _ptr(lse) if lse is not None else None,
                                           float(dp), int(dseed), int(doff),
                                           _ptr(block_live) if block_live is not None else None, _bf(q),
                                           current_stream_ptr()))
    return (out, lse) if return_lse else out


_SKIP_MASKED = os.environ.get("ATLAS_B200_ATTN_SKIP_MASKED", "1") != "0"    # A/B switch of the masked-key-block skipping


def key_block_live(add_mask, block=64):
    """uint8 [B, ceil(Lk / block)]: 1 where a block of `block` consecutive keys holds at least one key whose additive mask is
    > -5000 (a live key).  The reference masks with -10000 (`get_extended_attention_mask`, transformers 4.18) or -1e4 / -1e9
    (`invert_attention_mask`): the softmax weight of such a key is exp(-10000 + s - max) = 0.0 exactly in fp32, so a block
    without a live key contributes nothing and the attention kernels skip its loads, MMAs and exponentials - bit-identical
    results.  (Assumes |score| < ~4000, like the reference's own fp32 softmax needs to stay finite.)  A batch element without
    any live key keeps every block live: its uniform-over-masked-keys softmax is computed like the reference's.  Returns None
    when skipping is switched off (ATLAS_B200_ATTN_SKIP_MASKED=0) or there is no mask."""
    if add_mask is None or not _SKIP_MASKED:
        return None
    B, Lk = add_mask.shape
    nb = (Lk + block - 1) // block
    m = add_mask
    if nb * block != Lk:
        m = torch.nn.functional.pad(m, (0, nb * block - Lk), value=float("-inf"))
    live = (m.view(B, nb, block) > -5000.0).any(-1)
    live = live | ~live.any(-1, keepdim=True)
    return live.to(torch.uint8).contiguous()


_XKV_COMPACT = os.environ.get("ATLAS_B200_XKV_COMPACT", "1") != "0"        # A/B switch of the compacted cross K | V


def compact_live_tiles(x, tile_live):
    """x [n_tiles * 64, d] 16-bit, tile_live uint8 [n_tiles] (flattened key_block_live) -> (x_live: the rows of the live 64-row
    tiles packed to the front of a buffer of the same shape, tile_off int32 [n_tiles]: position of every live tile in it (-1 =
    dead), count_rows int32 [1] on the device = 64 x #live).  Static shapes, no host synchronisation."""
    require_cuda(x, "x")
    x2 = _rows2d(x)
    flags = tile_live.reshape(-1)
    n_tiles = flags.numel()
    if x2.shape[0] != n_tiles * 64:
        raise AtlasB200Error(f"compact_live_tiles: {x2.shape[0]} rows for {n_tiles} tiles of 64")
    dst = torch.empty_like(x2)
    tile_off = torch.empty(n_tiles, dtype=torch.int32, device=x.device)
    count = torch.empty(1, dtype=torch.int32, device=x.device)
    check(lib().atlas_b200_compact_live_tiles(_ptr(x2), x2.stride(0), _ptr(flags), n_tiles, x2.shape[1], _ptr(dst),
                                              dst.stride(0), _ptr(tile_off), _ptr(count), current_stream_ptr()))
    return dst, tile_off, count


_ENC_PACKED = os.environ.get("ATLAS_B200_ENC_PACKED", "1") != "0"          # A/B switch of the padding-compacted FiD encoder
_BERT_PACKED = os.environ.get("ATLAS_B200_BERT_PACKED", "1") != "0"        # ... of the Contriever encoder (needs _ENC_PACKED too)
_PACKED_BALANCE = os.environ.get("ATLAS_B200_PACKED_BALANCE", "1") != "0"  # A/B: packed attention CTAs split by work, not count


def segment_tile_scan(live):
    """live uint8 [S, nb] (key_block_live) -> (keep uint8 [S, nb], tile_off int32 [S * nb], tile_src int32 [S * nb], count_rows
    int32 [1]): the packed layout of the padding-compacted encoder (include/atlas_b200.h).  No host synchronisation.
    `keep` carries the work prefix of the segments (`keep._atlas_work`, int32 [S + 1]) that attention_packed balances its CTAs by."""
    require_cuda(live, "live")
    S, nb = live.shape
    live = live.contiguous()
    keep = torch.empty_like(live)
    tile_off = torch.empty(S * nb, dtype=torch.int32, device=live.device)
    tile_src = torch.empty(S * nb, dtype=torch.int32, device=live.device)
    count = torch.empty(1, dtype=torch.int32, device=live.device)
    work = torch.empty(S + 1, dtype=torch.int32, device=live.device)
    check(lib().atlas_b200_segment_tile_scan(_ptr(live), S, nb, _ptr(keep), _ptr(tile_off), _ptr(tile_src), _ptr(count),
                                             _ptr(work), current_stream_ptr()))
    keep._atlas_work = work
    return keep, tile_off, tile_src, count


def embed_packed_tiles(ids, table, tile_src):
    """Embedding rows of the token ids (int64 [n_tiles * 64]) of the kept tiles, in packed order; zeros past the end."""
    require_cuda(table, "table")
    ids = ids.reshape(-1)
    if ids.dtype != torch.int64:                   # torch's own embedding lookup takes int32 ids too
        if ids.dtype not in (torch.int32, torch.int16, torch.uint8):
            raise AtlasB200Error(f"embed_packed_tiles: integer token ids expected (got {ids.dtype})")
        ids = ids.to(torch.int64)
    ids = ids.contiguous()
    n_tiles = tile_src.numel()
    if ids.numel() != n_tiles * 64:
        raise AtlasB200Error(f"embed_packed_tiles: need {n_tiles * 64} token ids (got {ids.numel()})")
    out = torch.empty((n_tiles * 64, table.shape[1]), dtype=table.dtype, device=table.device)
    check(lib().atlas_b200_embed_packed_tiles(_ptr(ids), _ptr(table), table.stride(0), table.shape[0], _ptr(tile_src), n_tiles,
                                              _ptr(out), out.stride(0), table.shape[1], current_stream_ptr()))
    return out


def attention_packed(qkv, keep, tile_off, S, H, L, add_mask, bias_delta, scale=1.0, out=None):
    """Encoder self-attention on the packed rows (segment s = its kept tiles, back to back): atlas_b200_attention_packed."""
    require_cuda(qkv, "qkv")
    if out is None:
        out = torch.empty((qkv.shape[0], H * 64), dtype=qkv.dtype, device=qkv.device)
    am = add_mask.float().contiguous()
    bd = bias_delta.float().contiguous() if bias_delta is not None else None
    work = getattr(keep, "_atlas_work", None) if _PACKED_BALANCE else None
    check(lib().atlas_b200_attention_packed(_ptr(qkv), qkv.stride(0), 0, H * 64, 2 * H * 64, _ptr(out), out.stride(0), _ptr(am),
                                            _ptr(bd) if bd is not None else None, _ptr(keep), _ptr(tile_off),
                                            _ptr(work) if work is not None else None, S, H, L, float(scale), _bf(qkv),
                                            current_stream_ptr()))
    return out


def expand_packed_tiles(x, tile_off):
    """Packed rows -> the padded layout [n_tiles * 64, d]; dropped tiles are zero."""
    require_cuda(x, "x")
    n_tiles = tile_off.numel()
    out = torch.empty((n_tiles * 64, x.shape[1]), dtype=x.dtype, device=x.device)
    check(lib().atlas_b200_expand_packed_tiles(_ptr(x), x.stride(0), _ptr(tile_off), n_tiles, _ptr(out), out.stride(0),
                                               x.shape[1], current_stream_ptr()))
    return out


def linear_dynm(x, weight, m_dev, out=None):
    """y[m] = x[m] @ weight.T for m < *m_dev (device int32): row blocks past the device-side count are skipped."""
    require_cuda(x, "x")
    M, K = x.shape
    N = weight.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=x.dtype, device=x.device)
    check(lib().atlas_b200_linear_dynm(_ptr(x), x.stride(0), _ptr(weight), weight.stride(0), _ptr(out), out.stride(0), M, N, K,
                                       _ptr(m_dev), _bf(x), current_stream_ptr()))
    return out


def cross_attention_stream_compact(q, kv_live, tile_live, tile_off, B, H, Lq, Lk_total, add_mask, scale=1.0, return_lse=False):
    """`cross_attention_split` on the stream kernel with K | V given as the compacted live tiles (`compact_live_tiles` of the
    encoder output, projected by `linear_dynm`)."""
    require_cuda(q, "q")
    am = add_mask.float().contiguous()
    chunk = _XATTN_CHUNK
    chunks = (Lk_total + chunk - 1) // chunk
    o_part = torch.empty((B * chunks * Lq, H * 64), dtype=torch.float32, device=q.device)
    ml = torch.empty((B * chunks * Lq, H, 2), dtype=torch.float32, device=q.device)
    check(lib().atlas_b200_cross_attention_stream_compact(_ptr(q), q.stride(0), 0, _ptr(kv_live), kv_live.stride(0), 0, H * 64,
                                                          _ptr(am), _ptr(tile_live), _ptr(tile_off), B, H, Lq, Lk_total,
                                                          chunk, float(scale), _ptr(o_part), _ptr(ml), _bf(q),
                                                          current_stream_ptr()))
    out = torch.empty((B * Lq, H * 64), dtype=q.dtype, device=q.device)
    lse = torch.empty((B, H, Lq), dtype=torch.float32, device=q.device) if return_lse else None
    check(lib().atlas_b200_attention_combine_ex(_ptr(o_part), _ptr(ml), B, chunks, Lq, H, _ptr(out), out.stride(0),
                                                _ptr(lse) if lse is not None else None, _bf(q), current_stream_ptr()))
    return (out, lse) if return_lse else out


_XATTN_STREAM = os.environ.get("ATLAS_B200_XATTN_STREAM", "1") != "0"      # A/B switch of the stream kernel
_XATTN_CHUNK = int(os.environ.get("ATLAS_B200_XATTN_CHUNK", "1024"))        # keys per CTA (multiple of 64)


def cross_attention_split(q, q_col0, kv, k_col0, v_col0, B, H, Lq, Lk_total, add_mask=None, scale=1.0, split=512,
                          return_lse=False, dropout=None, tile_live=None):
    """Attention of Lq (<= 128) queries per batch element over Lk_total keys (FiD decoder cross-attention,
    Lk_total = n_ctx * L): split-KV over segments of `split` keys + combine.  q [B*Lq, ld], kv [B*Lk_total, ld]."""
    require_cuda(q, "q")
    am = add_mask.float().contiguous() if add_mask is not None else None
    if Lq <= 64 and dropout is None and Lk_total >= 1024 and _XATTN_STREAM:
        # few target tokens against the concatenated encoder keys: the K / V stream kernel (csrc/attention_stream.cu)
        chunk = _XATTN_CHUNK
        chunks = (Lk_total + chunk - 1) // chunk
        o_part = torch.empty((B * chunks * Lq, H * 64), dtype=torch.float32, device=q.device)
        ml = torch.empty((B * chunks * Lq, H, 2), dtype=torch.float32, device=q.device)
        if tile_live is None:
            tile_live = key_block_live(am)          # callers with many layers compute it once (fid.py)
        check(lib().atlas_b200_cross_attention_stream(_ptr(q), q.stride(0), q_col0, _ptr(kv), kv.stride(0), k_col0, v_col0,
                                                      _ptr(am) if am is not None else None,
                                                      _ptr(tile_live) if tile_live is not None else None, B, H, Lq, Lk_total,
                                                      chunk, float(scale), _ptr(o_part), _ptr(ml), _bf(q),
                                                      current_stream_ptr()))
        out = torch.empty((B * Lq, H * 64), dtype=q.dtype, device=q.device)
        lse = torch.empty((B, H, Lq), dtype=torch.float32, device=q.device) if return_lse else None
        check(lib().atlas_b200_attention_combine_ex(_ptr(o_part), _ptr(ml), B, chunks, Lq, H, _ptr(out), out.stride(0),
                                                    _ptr(lse) if lse is not None else None, _bf(q), current_stream_ptr()))
        return (out, lse) if return_lse else out
    if Lk_total % split != 0:
        raise AtlasB200Error(f"cross_attention_split: Lk_total={Lk_total} must be a multiple of {split}")
    splits = Lk_total // split
    o_part = torch.empty((B * splits * Lq, H * 64), dtype=torch.float32, device=q.device)
    ml = torch.empty((B * splits * Lq, H, 2), dtype=torch.float32, device=q.device)
    dummy = torch.empty((8,), dtype=q.dtype, device=q.device)
    dp, dseed, doff = dropout if dropout is not None else (0.0, 0, 0)
    check(lib().atlas_b200_attention_train(_ptr(q), q.stride(0), q_col0, _ptr(kv), kv.stride(0), k_col0, _ptr(kv),
                                           kv.stride(0), v_col0, _ptr(dummy), 8, _ptr(am) if am is not None else None, None,
                                           B * splits, H, Lq, split, float(scale), 0.0, splits, _ptr(o_part), _ptr(ml),
                                           None, float(dp), int(dseed), int(doff), None, _bf(q), current_stream_ptr()))
    out = torch.empty((B * Lq, H * 64), dtype=q.dtype, device=q.device)
    lse = torch.empty((B, H, Lq), dtype=torch.float32, device=q.device) if return_lse else None
    check(lib().atlas_b200_attention_combine_ex(_ptr(o_part), _ptr(ml), B, splits, Lq, H, _ptr(out), out.stride(0),
                                                _ptr(lse) if lse is not None else None, _bf(q), current_stream_ptr()))
    return (out, lse) if return_lse else out


# ---------------------------------------------------------------------------------------------
# backward pass (csrc/backward.cu, csrc/attention_bwd.cu): raw kernel wrappers; the autograd wiring is grad_ops.py
# ---------------------------------------------------------------------------------------------
def _rows2d(t):
    t2 = t.reshape(-1, t.shape[-1])
    return t2 if t2.stride(-1) == 1 else t2.contiguous()


def transpose(x, pad_to=8):
    """[R, C] 16-bit -> [C, Rpad] with Rpad = R rounded up to `pad_to` and the pad columns zero."""
    require_cuda(x, "x")
    x2 = _rows2d(x)
    R, C = x2.shape
    Rpad = (R + pad_to - 1) // pad_to * pad_to
    out = torch.empty((C, Rpad), dtype=x.dtype, device=x.device)
    check(lib().atlas_b200_transpose(_ptr(x2), x2.stride(0), _ptr(out), out.stride(0), R, C, Rpad, current_stream_ptr()))
    return out


def colsum(x):
    """fp32 [N] column sums of a 16-bit [M, N] matrix (bias gradient)."""
    require_cuda(x, "x")
    x2 = _rows2d(x)
    out = torch.zeros(x2.shape[1], dtype=torch.float32, device=x.device)
    check(lib().atlas_b200_colsum(_ptr(x2), x2.stride(0), _ptr(out), x2.shape[0], x2.shape[1], _bf(x), current_stream_ptr()))
    return out


def linear_dgrad(dy, weight):
    """dX [M, K] = dY [M, N] . W [N, K]: the same tcgen05 GEMM on W^T (a weight-sized transpose per call)."""
    return linear(dy, transpose(weight))


def linear_wgrad(dy, x):
    """dW [N, K] = dY^T [N, M] . X [M, K] (contraction over the M tokens, fp32 accumulation, 16-bit result): the tcgen05
    GEMM with both operands MN-major (no transposes).  ATLAS_B200_WGRAD_TRANSPOSE=1 selects the first-generation path
    (two explicit transposes + the K-major GEMM), kept for A/B measurements."""
    require_cuda(dy, "dy")
    if os.environ.get("ATLAS_B200_WGRAD_TRANSPOSE") == "1":
        return linear(transpose(dy), transpose(x))
    dy2, x2 = _rows2d(dy), _rows2d(x)
    if dy2.shape[0] != x2.shape[0] or dy2.dtype != x2.dtype:
        raise AtlasB200Error(f"linear_wgrad: dy {tuple(dy2.shape)} and x {tuple(x2.shape)} do not match")
    M, N = dy2.shape
    K = x2.shape[1]
    dw = torch.empty((N, K), dtype=dy.dtype, device=dy.device)
    nbytes = lib().atlas_b200_linear_wgrad_workspace_bytes(M, N, K)        # split-K partial tiles (0 = un-split)
    ws = _wgrad_ws.get(nbytes, dy.device) if nbytes else None
    check(lib().atlas_b200_linear_wgrad(_ptr(dy2), dy2.stride(0), _ptr(x2), x2.stride(0), _ptr(dw), dw.stride(0), M, N, K,
                                        _bf(dy), _ptr(ws) if ws is not None else None, ws.numel() if ws is not None else 0,
                                        current_stream_ptr()))
    return dw


def layernorm_bwd(x, dy, weight, eps, kind, dres=None, need_bias=False):
    """-> (dx like x, dweight fp32 [H], dbias fp32 [H] or None)."""
    require_cuda(x, "x")
    H = x.shape[-1]
    x2, dy2 = _rows2d(x), _rows2d(dy)
    r2 = _rows2d(dres) if dres is not None else None
    dx = torch.empty_like(x2)
    dw = torch.zeros(H, dtype=torch.float32, device=x.device)
    db = torch.zeros(H, dtype=torch.float32, device=x.device) if (need_bias or kind == 0) else None
    check(lib().atlas_b200_layernorm_bwd(_ptr(x2), x2.stride(0), _ptr(dy2), dy2.stride(0), _ptr(weight),
                                         _ptr(r2) if r2 is not None else None, r2.stride(0) if r2 is not None else 0,
                                         _ptr(dx), dx.stride(0), _ptr(dw), _ptr(db) if db is not None else None,
                                         x2.shape[0], H, float(eps), kind, _bf(x), current_stream_ptr()))
    return dx.reshape(x.shape), dw, db


def gated_gelu(u, dg=None):
    """u [M, 2F] interleaved (wi_0 | wi_1) pre-activations -> g [M, F]; with dg [M, F] -> du [M, 2F]."""
    require_cuda(u, "u")
    u2 = _rows2d(u)
    M, F2 = u2.shape
    F = F2 // 2
    if dg is None:
        out = torch.empty((M, F), dtype=u.dtype, device=u.device)
        check(lib().atlas_b200_gated_gelu(_ptr(u2), u2.stride(0), None, 0, _ptr(out), out.stride(0), M, F, _bf(u),
                                          current_stream_ptr()))
        return out.reshape(*u.shape[:-1], F)
    d2 = _rows2d(dg)
    out = torch.empty((M, F2), dtype=u.dtype, device=u.device)
    check(lib().atlas_b200_gated_gelu(_ptr(u2), u2.stride(0), _ptr(d2), d2.stride(0), _ptr(out), out.stride(0), M, F,
                                      _bf(u), current_stream_ptr()))
    return out.reshape(u.shape)


def gelu_erf(z, dy=None):
    """erf GELU of z (dy None) or its backward dy * gelu'(z)."""
    require_cuda(z, "z")
    z2 = _rows2d(z)
    d2 = _rows2d(dy) if dy is not None else None
    out = torch.empty_like(z2)
    check(lib().atlas_b200_gelu_erf(_ptr(z2), z2.stride(0), _ptr(d2) if d2 is not None else None,
                                    d2.stride(0) if d2 is not None else 0, _ptr(out), out.stride(0), z2.shape[0],
                                    z2.shape[1], _bf(z), current_stream_ptr()))
    return out.reshape(z.shape)


def dropout(x, p, seed, offset, residual=None):
    """out = (residual +) dropout(x): keep decisions are a pure function of (seed, offset, position), so calling this on
    dy with the same (seed, offset) is the backward of the forward call (csrc/dropout.cu)."""
    require_cuda(x, "x")
    x2 = _rows2d(x)
    r2 = _rows2d(residual) if residual is not None else None
    out = torch.empty_like(x2)
    check(lib().atlas_b200_dropout(_ptr(x2), x2.stride(0), _ptr(r2) if r2 is not None else None,
                                   r2.stride(0) if r2 is not None else 0, _ptr(out), out.stride(0), x2.shape[0],
                                   x2.shape[1], float(p), int(seed), int(offset), _bf(x), current_stream_ptr()))
    return out.reshape(x.shape)


def dropout_mask(M, N, p, seed, offset, device):
    """The keep mask (uint8 [M, N]) `dropout` applies to an [M, N] tensor with this (seed, offset)."""
    out = torch.empty((M, N), dtype=torch.uint8, device=device)
    check(lib().atlas_b200_dropout_mask(_ptr(out), M, N, float(p), int(seed), int(offset), current_stream_ptr()))
    return out


def attention_dropout_mask(B, H, Lq, Lk, p, seed, offset, device):
    """The keep mask (uint8 [B, H, Lq, Lk]) the attention kernels apply to the probabilities with this (seed, offset)."""
    out = torch.empty((B, H, Lq, Lk), dtype=torch.uint8, device=device)
    check(lib().atlas_b200_attention_dropout_mask(_ptr(out), B * H * Lq, Lk, float(p), int(seed), int(offset),
                                                  current_stream_ptr()))
    return out


_inf_flags = {}


def clamp_inf_(x, row_ss=None):
    """The reference's per-sub-layer fp16 overflow clamp (src/modeling_t5.py:657-708), in place, decided on the device
    (no host synchronisation).  A no-op for any other dtype, like the reference's `dtype == torch.float16` test."""
    if x.dtype != torch.float16:
        return x
    require_cuda(x, "x")
    x2 = _rows2d(x)
    flag = _inf_flags.get(x.device)
    if flag is None:
        flag = _inf_flags[x.device] = torch.zeros(1, dtype=torch.int32, device=x.device)
    check(lib().atlas_b200_clamp_inf_fp16(_ptr(x2), x2.stride(0), x2.shape[0], x2.shape[1], _ptr(flag),
                                          _ptr(row_ss) if row_ss is not None else None, current_stream_ptr()))
    return x


def bert_embed_sum(input_ids, token_type_ids, word_emb, type_emb, pos_emb):
    require_cuda(input_ids, "input_ids")
    B, L = input_ids.shape
    H = word_emb.shape[1]
    y = torch.empty((B, L, H), dtype=word_emb.dtype, device=word_emb.device)
    ids = input_ids.contiguous()
    tt = token_type_ids.contiguous() if token_type_ids is not None else None
    check(lib().atlas_b200_bert_embed_sum(_ptr(ids), _ptr(tt) if tt is not None else None, _ptr(word_emb), _ptr(type_emb),
                                          _ptr(pos_emb), _ptr(y), B, L, H, _bf(word_emb), current_stream_ptr()))
    return y


def scatter_add_rows(src, table_rows, index=None, modulo=0, skip_index=-1):
    """fp32 [table_rows, H] with dst[index[r]] += src[r] (index None: r % modulo)."""
    require_cuda(src, "src")
    s2 = _rows2d(src)
    H = s2.shape[1]
    dst = torch.zeros((table_rows, H), dtype=torch.float32, device=src.device)
    idx = index.reshape(-1).contiguous() if index is not None else None
    check(lib().atlas_b200_scatter_add_rows(_ptr(idx) if idx is not None else None, int(modulo), _ptr(s2), s2.stride(0),
                                            _ptr(dst), s2.shape[0], H, int(skip_index), int(table_rows), _bf(src),
                                            current_stream_ptr()))
    return dst


def masked_mean_pool_bwd(demb, mask, L, H):
    require_cuda(demb, "demb")
    d2 = demb if demb.stride(-1) == 1 else demb.contiguous()
    B = d2.shape[0]
    m = mask.to(torch.int64).contiguous()
    dx = torch.empty((B, L, H), dtype=demb.dtype, device=demb.device)
    if B > 1 and d2.stride(0) % 2:
        d2 = d2.clone(memory_format=torch.contiguous_format)
    ld = d2.stride(0) if B > 1 else H          # a single row: its (arbitrary) stride is never used
    check(lib().atlas_b200_masked_mean_pool_bwd(_ptr(d2), ld, _ptr(m), _ptr(dx), B, L, H, _bf(demb),
                                                current_stream_ptr()))
    return dx


def cross_entropy_fwd(logits, labels):
    """-> (lse fp32 [rows], per-row loss fp32 [rows], 0 where label == -100)."""
    require_cuda(logits, "logits")
    l2 = _rows2d(logits)
    rows, V = l2.shape
    y = labels.reshape(-1).to(torch.int64).contiguous()
    lse = torch.empty(rows, dtype=torch.float32, device=logits.device)
    loss = torch.empty(rows, dtype=torch.float32, device=logits.device)
    check(lib().atlas_b200_cross_entropy_fwd(_ptr(l2), l2.stride(0), _ptr(y), _ptr(lse), _ptr(loss), rows, V, _bf(logits),
                                             current_stream_ptr()))
    return lse, loss


def cross_entropy_bwd(logits, labels, lse, gscale):
    """dlogits = (softmax - onehot) * gscale[0] for the valid rows (gscale: fp32 device scalar)."""
    l2 = _rows2d(logits)
    rows, V = l2.shape
    y = labels.reshape(-1).to(torch.int64).contiguous()
    out = torch.empty_like(l2)
    gs = gscale.reshape(1).to(torch.float32).contiguous()
    check(lib().atlas_b200_cross_entropy_bwd(_ptr(l2), l2.stride(0), _ptr(y), _ptr(lse), _ptr(gs), _ptr(out),
                                             out.stride(0), rows, V, _bf(logits), current_stream_ptr()))
    return out.reshape(logits.shape)


def attention_bwd(q, q_col0, k, k_col0, v, v_col0, out, dout, dq, dq_col0, dk, dk_col0, dv, dv_col0, B, H, Lq, Lk,
                  add_mask=None, bias_delta=None, need_dbias=False, scale=1.0, causal_value=0.0, lse=None,
                  split_keys=False, dropout=None, block_live=None):
    """Backward of `attention` / `cross_attention_split` (un-split key range).  dq / dk / dv are written in place at
    their column offsets; returns dbias_delta fp32 [H, Lq+Lk-1] (or None).  `lse` = the forward's log-sum-exp
    (return_lse=True) saves the recomputation pass; split_keys (needs lse, no dbias; dq must be a whole contiguous
    [B*Lq, H*64] tensor) spreads a long key range over many CTAs (FiD cross-attention)."""
    require_cuda(q, "q")
    am = add_mask.float().contiguous() if add_mask is not None else None
    bd = bias_delta.float().contiguous() if bias_delta is not None else None
    dbias = torch.zeros((H, Lq + Lk - 1), dtype=torch.float32, device=q.device) if (need_dbias and bd is not None) else None
    dsum = torch.empty((B, H, Lq), dtype=torch.float32, device=q.device)
    lse_buf = lse.contiguous() if lse is not None else torch.empty((B, H, Lq), dtype=torch.float32, device=q.device)
    o2 = out if out.stride(-1) == 1 else out.contiguous()
    do2 = _rows2d(dout)
    accum = None
    if split_keys:
        if lse is None or dbias is not None or dq_col0 != 0 or not dq.is_contiguous() or dq.shape[1] != H * 64:
            raise AtlasB200Error("attention_bwd: split_keys needs the forward's lse, no dbias and a contiguous dq")
        accum = torch.zeros((B * Lq, H * 64), dtype=torch.float32, device=q.device)
    dp, dseed, doff = dropout if dropout is not None else (0.0, 0, 0)
    check(lib().atlas_b200_attention_bwd_train(
        _ptr(q), q.stride(0), q_col0, _ptr(k), k.stride(0), k_col0, _ptr(v), v.stride(0), v_col0, _ptr(o2), o2.stride(0),
        _ptr(do2), do2.stride(0), _ptr(dq), dq.stride(0), dq_col0, _ptr(dk), dk.stride(0), dk_col0, _ptr(dv),
        dv.stride(0), dv_col0, _ptr(am) if am is not None else None, _ptr(bd) if bd is not None else None,
        _ptr(dbias) if dbias is not None else None, _ptr(lse_buf), 1 if lse is not None else 0, _ptr(dsum),
        _ptr(accum) if accum is not None else None, B, H, Lq, Lk, float(scale), float(causal_value), float(dp),
        int(dseed), int(doff), _ptr(block_live) if block_live is not None else None, _bf(q), current_stream_ptr()))
    if accum is not None:
        check(lib().atlas_b200_cast_f32(_ptr(accum), _ptr(dq), accum.numel(), _bf(q), current_stream_ptr()))
    return dbias


def cross_attention_stats(q, q_col0, kv, k_col0, v_col0, B, H, T, Lk, lse, add_mask=None, scale=1.0):
    """Means over heads of the cross-attention logits, probabilities and ||V||-weighted probabilities -> three fp32
    [B, T, Lk] tensors (`score_storage`, `prob_storage`, `normalized_score_storage` of src/fid.py:333-343)."""
    require_cuda(q, "q")
    am = add_mask.float().contiguous() if add_mask is not None else None
    out = torch.empty((3, B, T, Lk), dtype=torch.float32, device=q.device)
    l2 = lse.contiguous()
    check(lib().atlas_b200_cross_attention_stats(_ptr(q), q.stride(0), q_col0, _ptr(kv), kv.stride(0), k_col0, v_col0,
                                                 _ptr(am) if am is not None else None, _ptr(l2), _ptr(out[0]), _ptr(out[1]),
                                                 _ptr(out[2]), B, H, T, Lk, float(scale), _bf(q), current_stream_ptr()))
    return out[0], out[1], out[2]


# ---------------------------------------------------------------------------------------------
# single-token decode (csrc/decode.cu): KV-cached greedy generation
# ---------------------------------------------------------------------------------------------
def decode_self_attention(qkv, cache, t_dev, H, bias_delta=None, scale=1.0, out=None):
    """qkv [B, 3*H*64] of the new token; cache [B, Tmax, 2*H*64] (k | v) updated in place at row t_dev[0]; -> ctx [B, H*64]."""
    require_cuda(qkv, "qkv")
    B, Tmax = cache.shape[0], cache.shape[1]
    if out is None:
        out = torch.empty((B, H * 64), dtype=qkv.dtype, device=qkv.device)
    bd = bias_delta.float().contiguous() if bias_delta is not None else None
    check(lib().atlas_b200_decode_self_attention(_ptr(qkv), qkv.stride(0), _ptr(cache), Tmax, _ptr(t_dev),
                                                 _ptr(bd) if bd is not None else None, float(scale), _ptr(out),
                                                 out.stride(0), B, H, _bf(qkv), current_stream_ptr()))
    return out


def decode_cross_attention(q, kv, B, H, Lk, add_mask=None, scale=1.0, chunk=256, out=None, tile_live=None):
    """q [B, H*64] against kv [B*Lk, 2*H*64] (k | v): chunked partials + the split-KV combine -> ctx [B, H*64].
    tile_live = key_block_live(add_mask): 64-key tiles of masked keys only are not read (identical result)."""
    require_cuda(q, "q")
    if tile_live is not None and chunk % 64 != 0:
        tile_live = None
    chunks = (Lk + chunk - 1) // chunk
    o_part = torch.empty((B * chunks, H * 64), dtype=torch.float32, device=q.device)
    ml = torch.empty((B * chunks, H, 2), dtype=torch.float32, device=q.device)
    am = add_mask.float().contiguous() if add_mask is not None else None
    check(lib().atlas_b200_decode_cross_attention_live(_ptr(q), q.stride(0), _ptr(kv), kv.stride(0), 0, H * 64,
                                                       _ptr(am) if am is not None else None,
                                                       _ptr(tile_live) if tile_live is not None else None, B, H, Lk, chunk,
                                                       float(scale), _ptr(o_part), _ptr(ml), _bf(q), current_stream_ptr()))
    if out is None:
        out = torch.empty((B, H * 64), dtype=q.dtype, device=q.device)
    check(lib().atlas_b200_attention_combine_ex(_ptr(o_part), _ptr(ml), B, chunks, 1, H, _ptr(out), out.stride(0), None,
                                                _bf(q), current_stream_ptr()))
    return out


def decode_argmax(logits, seq, tok_in, done, t_dev, eos_id, pad_id, min_length):
    """Greedy pick on the device (see include/atlas_b200.h); advances t_dev."""
    require_cuda(logits, "logits")
    B, V = logits.shape
    check(lib().atlas_b200_decode_argmax(_ptr(logits), logits.stride(0), V, _ptr(seq), seq.stride(0), _ptr(tok_in),
                                         _ptr(done), _ptr(t_dev), int(eos_id), int(pad_id), int(min_length), B, _bf(logits),
                                         current_stream_ptr()))
