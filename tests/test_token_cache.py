"""CPU: the retriever-side passage token cache (atlas_b200/token_cache.py) hands `Atlas.build_index` exactly the tensors the
reference's tokenizer call produces (`padding="longest"`, `truncation=True`, src/atlas.py:69-76), for any batching."""
import pytest
import torch

import atlas_synth
from atlas_b200 import token_cache


def _tokenise(tok, passages, fmt, max_len):
    return tok([fmt.format(**p) for p in passages], padding="longest", return_tensors="pt", max_length=max_len,
               truncation=True)


@pytest.mark.parametrize("max_len", [64, 12])
def test_batches_equal_tokenizer_output(max_len):
    _, tok = atlas_synth.tokenizers()
    passages = atlas_synth.make_corpus()
    fmt = "{title} {text}"
    cache = token_cache.RetrieverTokenCache(len(passages), max_len)
    for a in range(0, len(passages), 32):                       # recorded with one batching ...
        enc = _tokenise(tok, passages[a:a + 32], fmt, max_len)
        cache.append(enc["input_ids"], enc["attention_mask"])
    assert cache.complete
    for bs in (32, 7, 96, 1):                                   # ... served with any other
        for a in range(0, len(passages), bs):
            b = min(len(passages), a + bs)
            ids, mask = cache.batch(a, b)
            enc = _tokenise(tok, passages[a:b], fmt, max_len)
            assert torch.equal(ids, enc["input_ids"]) and torch.equal(mask, enc["attention_mask"].to(torch.int64))
    with pytest.raises(ValueError):
        cache.append(torch.zeros(1, 1, dtype=torch.long), torch.ones(1, 1, dtype=torch.long))   # already full


def test_key_and_budget():
    p = atlas_synth.make_corpus()
    k = token_cache.cache_key(p, 64, "{title} {text}")
    assert k == token_cache.cache_key(p, 64, "{title} {text}")
    assert k != token_cache.cache_key(list(p), 64, "{title} {text}") and k != token_cache.cache_key(p, 32, "{title} {text}")
    assert token_cache.fits(4 << 20, 384, 8 << 30) and not token_cache.fits(32 << 20, 512, 8 << 30)


def test_build_index_refresh_uses_the_record(monkeypatch):
    """Second `build_index` over the same shard: no tokenizer call, identical (ids, mask) batches reach the encoder."""
    from types import SimpleNamespace

    import atlas_b200.atlas as A

    monkeypatch.setattr(A, "_to_cuda", lambda d: d)
    monkeypatch.setattr(A, "_device", lambda: torch.device("cpu"))
    reader_tok, retr_tok = atlas_synth.tokenizers()
    calls = {"tok": 0}

    class CountingTok:
        def __call__(self, *a, **k):
            calls["tok"] += 1
            return retr_tok(*a, **k)

    seen = []

    class Tower(torch.nn.Module):
        def embed_into(self, ids, mask, rows, dtype=None):
            seen.append((ids.clone(), mask.clone()))
            rows.copy_(ids.float().sum(1, keepdim=True).expand_as(rows).to(rows.dtype))

    retriever = SimpleNamespace(contriever=Tower())
    opt = atlas_synth.make_opt()
    model = A.Atlas(opt, torch.nn.Linear(1, 1), retriever, reader_tok, CountingTok())
    passages = atlas_synth.make_corpus()
    index = SimpleNamespace(_bank=torch.zeros(len(passages), 8, dtype=torch.float16), is_index_trained=lambda: True)
    model.build_index(index, passages, 32)
    first, n_tok, bank1 = list(seen), calls["tok"], index._bank.clone()
    assert n_tok == 3 and model._token_cache is not None and model._token_cache.complete
    seen.clear()
    index._bank.zero_()
    model.build_index(index, passages, 32)
    assert calls["tok"] == n_tok, "the refresh tokenised again"
    assert len(seen) == len(first) and all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(seen, first))
    assert torch.equal(index._bank, bank1)
    # another shard (or cache_retriever_tokens = False) goes back to the tokenizer
    model.build_index(index, passages[:40], 32)
    assert calls["tok"] == n_tok + 2


@pytest.mark.parametrize("fmt,max_len", [("{query} title: {title} context: {text}", 64),
                                         ("{query} title: {title} context: {text}", 20),
                                         ("question: {query} passage: {text}", 48)])
def test_reader_rows_equal_full_tokenisation(monkeypatch, fmt, max_len):
    """`Atlas.tokenize_passages` with `cache_reader_tokens`: the spliced rows are tensor-equal to tokenising the
    concatenated "query + passage" strings (src/atlas.py:261-280), incl. ragged passage counts (EOS-only padding
    examples), truncation, and repeated retrieval of the same passages."""
    import atlas_b200.atlas as A

    monkeypatch.setattr(A, "_to_cuda", lambda d: d)
    reader_tok, retr_tok = atlas_synth.tokenizers()
    corpus = atlas_synth.make_corpus()
    query, _ = atlas_synth.make_batch()
    passages = [corpus[0:4], corpus[2:5], corpus[10:11]]
    plain = A.Atlas(atlas_synth.make_opt(encoder_format=fmt, text_maxlength=max_len), torch.nn.Linear(1, 1), None,
                    reader_tok, retr_tok)
    cached = A.Atlas(atlas_synth.make_opt(encoder_format=fmt, text_maxlength=max_len, cache_reader_tokens=True),
                     torch.nn.Linear(1, 1), None, reader_tok, retr_tok)
    for _ in range(2):                                         # second round: every passage part comes from the cache
        want, _ = plain.tokenize_passages(query, passages)
        got, _ = cached.tokenize_passages(query, passages)
        assert torch.equal(got["input_ids"], want["input_ids"]) and torch.equal(got["attention_mask"], want["attention_mask"])
    rc = cached._reader_cache
    assert rc.usable and rc.hits >= rc.misses > 0


def test_encoder_format_split():
    assert token_cache.split_encoder_format("{query} title: {title} context: {text}") == ("{query} title: ", "{title} context: {text}")
    assert token_cache.split_encoder_format("{title} {query} {text}") is None          # query after a passage field
    assert token_cache.split_encoder_format("{query}:{text}") is None                   # cut not on white space
    assert token_cache.split_encoder_format("{query}") is None
