"""Seeded weights / inputs for the model-level goldens (Contriever, FiD).

TEST INFRASTRUCTURE.  Both the golden generator (which fills the REFERENCE modules) and the tests (which
fill atlas_b200's modules) call `fill_state_dict` on a state dict with the same keys, so the two models
hold identical parameters without shipping 30 MB of weights; a sha256 of the values is stored next to the
golden outputs."""
import hashlib

import numpy as np
import torch


def _canonical(key):
    """Tied tensors share one name so every alias receives the same values."""
    if key.endswith("embed_tokens.weight"):
        return "shared.weight"
    return key


def fill_state_dict(sd, seed):
    """Deterministic values for every floating tensor of `sd`.  Each tensor has its own stream seeded by
    (seed, crc32(name)), so models whose state dicts differ in auxiliary keys still get identical weights."""
    import zlib

    h = hashlib.sha256()
    out = {}
    for k in sorted(sd.keys()):
        t = sd[k]
        if not torch.is_floating_point(t):
            out[k] = t
            continue
        ck = _canonical(k)
        rng = np.random.default_rng([seed, zlib.crc32(ck.encode())])
        shape = tuple(t.shape)
        x = rng.standard_normal(shape, dtype=np.float32)
        if ck.endswith("LayerNorm.weight") or ck.endswith("layer_norm.weight"):
            x = 1.0 + 0.1 * x
        elif ck.endswith(".bias"):
            x = 0.02 * x
        elif "relative_attention_bias" in ck:
            x = 0.5 * x
        elif "embeddings" in ck or ck == "shared.weight":
            x = 0.5 * x
        elif len(shape) == 2:
            x = x * (0.6 / np.sqrt(shape[1]))
        h.update(ck.encode())
        h.update(np.ascontiguousarray(x).tobytes())
        out[k] = torch.from_numpy(x)
    return out, h.hexdigest()


CONTRIEVER_CFG = dict(vocab_size=2000, hidden_size=768, num_hidden_layers=2, num_attention_heads=12,
                      intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12)


def contriever_inputs(seed=21, B=6, L=48, vocab=2000):
    rng = np.random.default_rng(seed)
    lens = rng.integers(L // 3, L + 1, size=B)
    lens[0] = L
    ids = rng.integers(1, vocab, size=(B, L))
    mask = (np.arange(L)[None, :] < lens[:, None]).astype(np.int64)
    ids = ids * mask  # pad token 0
    return torch.from_numpy(ids.astype(np.int64)), torch.from_numpy(mask)


T5_CFG = dict(vocab_size=512, d_model=768, d_kv=64, d_ff=2048, num_layers=2, num_decoder_layers=2, num_heads=12,
              relative_attention_num_buckets=32, dropout_rate=0.1, layer_norm_epsilon=1e-6,
              feed_forward_proj="gated-gelu", tie_word_embeddings=False, decoder_start_token_id=0, pad_token_id=0,
              eos_token_id=1, is_encoder_decoder=True, use_cache=False)


def fid_inputs(seed=33, B=2, n_ctx=3, L=64, T=8, vocab=512):
    rng = np.random.default_rng(seed)
    lens = rng.integers(L // 2, L + 1, size=(B, n_ctx))
    ids = rng.integers(2, vocab, size=(B, n_ctx, L))
    mask = (np.arange(L)[None, None, :] < lens[:, :, None])
    ids = ids * mask
    tlen = rng.integers(3, T + 1, size=B)
    tlen[0] = T
    labels = rng.integers(2, vocab, size=(B, T))
    tmask = np.arange(T)[None, :] < tlen[:, None]
    labels = np.where(tmask, labels, -100)
    return (torch.from_numpy(ids.reshape(B, n_ctx * L).astype(np.int64)),
            torch.from_numpy(mask.reshape(B, n_ctx * L)),
            torch.from_numpy(labels.astype(np.int64)))


def fid_inputs_with_sep(seed=33, B=2, n_ctx=3, L=64, T=8, vocab=512, n_query=5):
    """`fid_inputs` with an EOS (id 1) closing every passage, plus the reader's query mask (first `n_query` tokens of a
    passage are the question, src/atlas.py:407-413): inputs of the cross-attention score goldens."""
    ids, mask, labels = fid_inputs(seed, B, n_ctx, L, T, vocab)
    ids = ids.clone().view(B, n_ctx, L)
    m = mask.view(B, n_ctx, L)
    last = m.sum(-1) - 1
    ids.scatter_(2, last[..., None], 1)
    mask_query = torch.zeros(B, n_query + 3, dtype=torch.bool)
    mask_query[:, :n_query] = True
    return ids.view(B, n_ctx * L), mask, labels, mask_query
