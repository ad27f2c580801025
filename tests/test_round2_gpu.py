"""GPU tests of the round-2 additions: the device token bank + splice kernel, the KV-cached greedy decode
(§8f-2), the in-place 16-bit weight cache, the sync-free search entry points."""
import numpy as np
import pytest
import torch

import atlas_synth
import model_synth
from fake_tokenizer import FakeTokenizer

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    from atlas_b200._lib import lib

    lib()
    return torch.device("cuda:0")


def _tiny_fid(dev, dtype=torch.bfloat16, **over):
    from atlas_b200.fid import FiD, T5ConfigLite

    cfg = {k: v for k, v in dict(model_synth.T5_CFG, **over).items() if k not in ("dropout_rate", "is_encoder_decoder", "use_cache")}
    model = FiD(T5ConfigLite(**cfg))
    sd, _ = model_synth.fill_state_dict(model.state_dict(), 202)
    model.load_state_dict(sd)
    return model.to(dtype).to(dev).eval()


def test_token_bank_splice_equals_host_tokenisation(dev):
    """`DeviceTokenBank.splice` == `encode_passages([[encoder_format.format(query=q, **p)]])` (src/atlas.py:26-39,261-270) for a
    white-space tokenizer: ragged passages, truncation at text_maxlength, fewer passages than n (EOS-only padding rows)."""
    from atlas_b200.atlas import encode_passages
    from atlas_b200.token_bank import DeviceTokenBank

    tok = FakeTokenizer("t5", 512)
    passages = atlas_synth.make_corpus()
    fmt = "{query} title: {title} context: {text}"
    bank = DeviceTokenBank.build(passages, tok, fmt, 48, dev)
    rng = np.random.default_rng(3)
    queries = ["w1 w2 w3", "w9 w8 w7 w6 w5 w4 w3 w2 w1 w0 w11 w12", "w5"]
    picks = [list(rng.choice(len(passages), 4, replace=False)) for _ in queries]
    picks[2] = picks[2][:2]                                           # a query with fewer passages
    for L in (24, 64):
        text = [[fmt.format(query=q, **passages[i]) for i in row] for q, row in zip(queries, picks)]
        want = encode_passages(text, tok, L)
        gids = torch.tensor([row + [-1] * (4 - len(row)) for row in picks], dtype=torch.int64, device=dev)
        q_ids, q_lens = bank.query_tokens(tok, queries, dev)
        got = bank.splice(gids, L, q_ids, q_lens)
        assert torch.equal(got["input_ids"].cpu(), want["input_ids"])
        assert torch.equal(got["attention_mask"].cpu().long(), want["attention_mask"])


def test_atlas_uses_the_token_bank(dev):
    from atlas_b200.atlas import Atlas
    from atlas_b200.retrievers import BertConfigLite, Contriever, DualEncoderRetriever
    from atlas_b200.token_bank import DeviceTokenBank

    opt = atlas_synth.make_opt()
    reader_tok, retriever_tok = atlas_synth.tokenizers()
    reader = _tiny_fid(dev)
    model = Atlas(opt, reader, DualEncoderRetriever(opt, Contriever(BertConfigLite(**model_synth.CONTRIEVER_CFG)).to(dev)),
                  reader_tok, retriever_tok).eval()
    passages = atlas_synth.make_corpus()
    query, _ = atlas_synth.make_batch()
    chosen = [[passages[(7 * b + j) % len(passages)] for j in range(atlas_synth.TOPK)] for b in range(len(query))]
    host, _ = model.tokenize_passages(query, chosen)
    model.set_token_bank(DeviceTokenBank.build(passages, reader_tok, opt.encoder_format, opt.text_maxlength, dev))
    banked, _ = model.tokenize_passages(query, chosen)
    assert torch.equal(banked["input_ids"], host["input_ids"])
    assert torch.equal(banked["attention_mask"].long(), host["attention_mask"].long())


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_generate_kv_cached_matches_prefix_rerun(dev, dtype):
    """The KV-cached single-token decode (csrc/decode.cu, graph-replayed) against the first-generation path that re-runs the
    whole decoder prefix every step, and against the teacher-forced argmax chain: identical tokens wherever the top-2
    logit gap of the teacher-forced run exceeds the 16-bit noise of the two attention formulations."""
    model = _tiny_fid(dev, dtype)
    ids, mask, _ = model_synth.fid_inputs()
    model.encoder.config.n_context, model.encoder.config.bsz = 3, 2
    with torch.no_grad():
        a = model.generate(input_ids=ids.to(dev), attention_mask=mask.to(dev), max_length=9, min_length=3)
        b = model.generate(input_ids=ids.to(dev), attention_mask=mask.to(dev), max_length=9, min_length=3, use_cache=False)
        a2 = model.generate(input_ids=ids.to(dev), attention_mask=mask.to(dev), max_length=9, min_length=3)   # graph replay
    assert a.shape[0] == 2 and a.shape[1] <= 9 and int(a[0, 0]) == 0
    with torch.no_grad():
        tf = model(input_ids=ids.to(dev), attention_mask=mask.to(dev), decoder_input_ids=b[:, :-1], use_cache=False).logits.float()
    top2 = tf.topk(2, dim=-1)[0]
    gap_ok = (top2[..., 0] - top2[..., 1]) > (0.05 if dtype == torch.bfloat16 else 0.01)
    # the encoder is not bit-reproducible run to run (the fused RMSNorm accumulates each row's sum of squares with fp32
    # atomics, csrc/gemm.cu), so the first run and the graph replay are held to the same near-tie rule as the two paths
    for got in (a, a2):
        n = min(got.shape[1], b.shape[1])
        for bi in range(2):
            for t in range(1, n):
                if not bool(gap_ok[bi, t - 1]):
                    break                              # a near-tie may legitimately fork the greedy chains from here on
                assert int(got[bi, t]) == int(b[bi, t]), (bi, t, got[bi].tolist(), b[bi].tolist())


def test_generate_with_prefix_constraint(dev):
    """`prefix_allowed_tokens_fn` (decoder prompt, src/atlas.py:621-636) through the cached path (eager, host callback)."""
    model = _tiny_fid(dev)
    ids, mask, _ = model_synth.fid_inputs()
    model.encoder.config.n_context, model.encoder.config.bsz = 3, 2
    forced = [[17, 23], [31, 5]]

    def allowed(b, seq):
        pos = seq.shape[-1]
        return forced[b][pos - 1] if pos <= 2 else list(range(512))

    with torch.no_grad():
        out = model.generate(input_ids=ids.to(dev), attention_mask=mask.to(dev), max_length=6, prefix_allowed_tokens_fn=allowed)
        ref = model.generate(input_ids=ids.to(dev), attention_mask=mask.to(dev), max_length=6, prefix_allowed_tokens_fn=allowed,
                             use_cache=False)
    assert out[:, 1:3].tolist() == forced and ref[:, 1:3].tolist() == forced


def test_generate_base_size_kv_cached(dev):
    """Base-size decoder (12 layers, 32 128-entry vocabulary, 15 360 cross keys, B = 2): the cached greedy chain equals the
    teacher-forced argmax of the full forward wherever the top-2 gap exceeds the 16-bit noise."""
    model = _tiny_fid(dev, torch.bfloat16, vocab_size=32128, num_layers=2, num_decoder_layers=12)
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(2, 32000, (2, 40 * 384), generator=g).to(dev)
    mask = torch.ones(2, 40 * 384, dtype=torch.bool, device=dev)
    model.encoder.config.n_context, model.encoder.config.bsz = 40, 2
    with torch.no_grad():
        seq = model.generate(input_ids=ids, attention_mask=mask, max_length=12, min_length=12)
        tf = model(input_ids=ids, attention_mask=mask, decoder_input_ids=seq[:, :-1], use_cache=False).logits.float()
    assert seq.shape == (2, 12)
    top2 = tf.topk(2, dim=-1)[0]
    gap_ok = (top2[..., 0] - top2[..., 1]) > 0.05
    pred = tf.argmax(-1)
    for bi in range(2):
        for t in range(11):
            if not bool(gap_ok[bi, t]):
                break
            assert int(pred[bi, t]) == int(seq[bi, t + 1]), (bi, t)


def test_halfcache_refreshes_in_place(dev):
    """writes through `param.data` (no version bump) are picked up after `invalidate()` and after any
    optimizer.step(); the 16-bit buffers (and therefore captured CUDA graphs) keep their addresses."""
    from atlas_b200.retrievers import BertConfigLite, Contriever

    model = Contriever(BertConfigLite(**model_synth.CONTRIEVER_CFG)).to(dev).eval()      # fp32 parameters -> fp16 copies
    ids, mask = model_synth.contriever_inputs()
    ids, mask = ids.to(dev), mask.to(dev)
    with torch.no_grad():
        e0 = model(input_ids=ids, attention_mask=mask).clone()
        ptr0 = model._half.sets[torch.float16]["store"]["encoder.layer.0.output.dense.weight"].data_ptr()
        w = model.encoder.layer[1].output.dense.weight
        w.data.mul_(1.5)                                            # bypasses the autograd version counter
        model._half.invalidate()
        e1 = model(input_ids=ids, attention_mask=mask).clone()
        assert float((e1 - e0).abs().max()) > 1e-3
        assert model._half.sets[torch.float16]["store"]["encoder.layer.0.output.dense.weight"].data_ptr() == ptr0
    # an optimizer step marks every cache stale through the global post-step hook
    opt = torch.optim.SGD([w], lr=0.5)
    w.grad = torch.ones_like(w) * 0.01
    opt.step()
    with torch.no_grad():
        e2 = model(input_ids=ids, attention_mask=mask)
    assert float((e2 - e1).abs().max()) > 1e-4


def test_fid_graph_survives_weight_update(dev):
    """The no-grad forward replays ONE captured graph across weight updates (scoring passes of a training loop)."""
    model = _tiny_fid(dev)
    ids, mask, labels = model_synth.fid_inputs()
    model.encoder.config.n_context, model.encoder.config.bsz = 3, 2
    kw = dict(input_ids=ids.to(dev), attention_mask=mask.to(dev), labels=labels.to(dev))
    with torch.no_grad():
        l0 = float(model(**kw)[0])
        n_graphs = len(model._graphs)
        with torch.no_grad():
            model.lm_head.weight.mul_(0.5)                          # in-place update (version bump)
        l1 = float(model(**kw)[0])
        assert len(model._graphs) == n_graphs, "a weight update must not trigger a re-capture"
        assert abs(l1 - l0) > 1e-3
        model.cuda_graphs = False
        l1_eager = float(model(**kw)[0])
    assert abs(l1 - l1_eager) < 1e-3


def test_search_device_deferred_status(dev):
    """`search_device(return_status=True)` never synchronises; an all-zero bank (massive ties) raises the flag and the
    exhaustive retry through `search_knn` is exact."""
    import synth
    from atlas_b200.index import DistributedIndex

    n = 70000
    index = DistributedIndex()
    index.init_embeddings(synth.make_passages(n))
    q = torch.from_numpy(synth.make_queries(5, seed=3)).to(dev)
    s, i, status = index.search_device(q, 40, return_status=True)
    assert int(status) != 0                                          # every score ties at 0: the candidate lists overflow
    docs, scores = index.search_knn(q, 40)
    assert [int(d["id"]) for d in docs[0]] == list(range(40)) and scores[0] == [0.0] * 40
