"""Generate tests/golden/atlas_tiny.npz by running the UNMODIFIED reference `src.atlas.Atlas` (with the reference
`DistributedIndex`, `Contriever`/`DualEncoderRetriever` and `FiD`) on CPU under oracle/ref_shims.py, with seeded
weights (oracle/model_synth.py), a seeded corpus / batch (oracle/atlas_synth.py) and the deterministic fake
tokenizers (no vocabulary files offline).  `Tensor.cuda` is patched to a no-op: the reference hard-codes `.cuda()`.

Stored per stage, so the GPU test can check each §8(a) row separately:
  bank_fp16       [N,768]   index.embeddings.T after `Atlas.build_index`           (src/atlas.py:61-88)
  bank_fp32       [N,768]   the same passages embedded by the fp32 retriever (accuracy budget)
  ret_ids/scores  [B,k]     `Atlas.retrieve`                                         (src/atlas.py:90-182)
  rerank_ids/scores         `retrieve_with_rerank`
  reader_loss, retriever_loss, gold_score [B,k]  `Atlas.forward(train_retriever=True)`, ppmean, eval mode
  loop_gold       [B,k]     `Atlas.loop_score`
  eval_loss, eval_logits    `compute_reader_loss_and_logits`
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import atlas_synth  # noqa: E402
import model_synth  # noqa: E402
import ref_shims  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    ref_shims.install()
    torch.Tensor.cuda = lambda self, *a, **kw: self
    torch.set_num_threads(8)
    from transformers import BertConfig, T5Config
    from src.atlas import Atlas
    from src.fid import FiD
    from src.index import DistributedIndex
    from src.retrievers import Contriever, DualEncoderRetriever

    opt = atlas_synth.make_opt()
    reader_tok, retriever_tok = atlas_synth.tokenizers()
    cfg = T5Config(**model_synth.T5_CFG)
    cfg.tie_word_embeddings = False
    reader = FiD(cfg).eval()
    sd, sha_r = model_synth.fill_state_dict(reader.state_dict(), seed=202)
    reader.load_state_dict(sd)
    contriever = Contriever(BertConfig(**model_synth.CONTRIEVER_CFG)).eval()
    sd, sha_c = model_synth.fill_state_dict(contriever.state_dict(), seed=101)
    contriever.load_state_dict(sd)
    retriever = DualEncoderRetriever(opt, contriever)
    model = Atlas(opt, reader, retriever, reader_tok, retriever_tok).eval()

    passages = atlas_synth.make_corpus()
    query, target = atlas_synth.make_batch()
    index = DistributedIndex()
    index.is_in_gpu = False
    index.init_embeddings(passages)

    class _Log:
        def info(self, *a):
            pass

    out = {"reader_sha256": sha_r, "retriever_sha256": sha_c}
    with torch.no_grad():
        model.build_index(index, passages, opt.per_gpu_embedder_batch_size, _Log())
        out["bank_fp16"] = index.embeddings.T.contiguous().numpy()
        # fp32 embeddings of the same passages (what the fp16 copy approximates)
        embs = []
        for i in range(0, len(passages), opt.per_gpu_embedder_batch_size):
            chunk = [opt.retriever_format.format(**p) for p in passages[i:i + opt.per_gpu_embedder_batch_size]]
            enc = retriever_tok(chunk, padding="longest", return_tensors="pt",
                                max_length=min(opt.text_maxlength, opt.per_gpu_embedder_batch_size), truncation=True)
            embs.append(retriever(**enc, is_passages=True))
        out["bank_fp32"] = torch.cat(embs).numpy()

        query_enc = model.retriever_tokenize(query)
        ps, sc = model.retrieve(index, atlas_synth.TOPK, query, query_enc["input_ids"], query_enc["attention_mask"])
        out["ret_ids"] = np.array([[int(p["id"]) for p in row] for row in ps])
        out["ret_scores"] = np.array(sc, dtype=np.float32)
        out["query_emb"] = retriever(**query_enc, is_passages=False).numpy()
        out["all_scores_fp16"] = torch.matmul(torch.from_numpy(out["query_emb"]).half(), index.embeddings).float().numpy()

        opt.retrieve_with_rerank = True
        ps2, sc2 = model.retrieve(index, atlas_synth.TOPK, query, query_enc["input_ids"], query_enc["attention_mask"])
        opt.retrieve_with_rerank = False
        out["rerank_ids"] = np.array([[int(p["id"]) for p in row] for row in ps2])
        out["rerank_scores"] = np.array(sc2, dtype=np.float32)

        # the whole step (retrieve -> read -> losses); the reference retrieves `ps` again (deterministic on CPU)
        stats = {}
        reader_loss, retriever_loss = model(index, query, target, train_retriever=True, iter_stats=stats)
        out["reader_loss"] = np.array(float(reader_loss))
        out["retriever_loss"] = np.array(float(retriever_loss))
        reader_tokens, retriever_tokens = model.tokenize_passages(query, ps)
        _, labels, dec_in = model.tokenize(query, target, None)
        cfgm = reader.encoder.config
        rid, rmask = reader_tokens["input_ids"], reader_tokens["attention_mask"].bool()
        out["gold_ppmean"] = model.perplexity_score(rid, rmask, dec_in, labels, cfgm, len(query)).numpy()
        out["gold_loop"] = model.loop_score(rid, rmask, dec_in, labels, cfgm, len(query)).numpy()
        el, logits = model.compute_reader_loss_and_logits(reader_tokens, dec_in, labels)
        out["eval_loss"] = np.array(el)
        out["eval_logits"] = logits.float().numpy()
        out["labels"] = labels.numpy()
        # emdr objective on the same batch
        opt.gold_score_mode = "emdr"
        _, emdr_loss = model(index, query, target, train_retriever=True, iter_stats={})
        out["emdr_loss"] = np.array(float(emdr_loss))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "atlas_tiny.npz"), **out)
    for k, v in out.items():
        if isinstance(v, np.ndarray):
            print(k, v.shape, v.dtype, (v.reshape(-1)[:4] if v.size else v))
    print("bank fp16-vs-fp32 max abs", np.abs(out["bank_fp16"].astype(np.float32) - out["bank_fp32"]).max())


if __name__ == "__main__":
    main()
