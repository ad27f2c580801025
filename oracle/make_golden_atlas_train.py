"""Generate tests/golden/atlas_train_tiny.npz: one TRAINING step of the UNMODIFIED reference `src.atlas.Atlas` on CPU
(`reader_loss, retriever_loss = model(...)`, `(reader_loss + retriever_loss).backward()` as train.py:90-101 does) under
oracle/ref_shims.py, seeded weights / corpus / batch, fake tokenizers.

TEST INFRASTRUCTURE.  Deterministic set-up: modules in eval mode (dropout off) with gradients enabled, passages given
explicitly (`retrieve` stubbed to return the reference's own retrieval result of atlas_tiny.npz) so the batch cannot depend on
16-bit score ties, ppmean gold scores, temperatures 0.1 (at the default 0.01 the KL target is a near one-hot whose
winner flips with 16-bit noise in the reference itself).  Stored: both losses, gold scores, and per parameter the
gradient's L2 norm / seeded random projection (fp32 run and the reference's own bf16-reader run as accuracy budget)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import atlas_synth  # noqa: E402
import model_synth  # noqa: E402
import ref_shims  # noqa: E402
from grad_oracle import direction  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
OPT_OVER = dict(temperature_gold=0.1, temperature_score=0.1)


def run(reader_dtype, out, tag):
    from transformers import BertConfig, T5Config
    from src.atlas import Atlas
    from src.fid import FiD
    from src.retrievers import Contriever, DualEncoderRetriever

    opt = atlas_synth.make_opt(**OPT_OVER)
    reader_tok, retriever_tok = atlas_synth.tokenizers()
    cfg = T5Config(**model_synth.T5_CFG)
    cfg.tie_word_embeddings = False
    reader = FiD(cfg).eval()
    sd, _ = model_synth.fill_state_dict(reader.state_dict(), seed=202)
    reader.load_state_dict(sd)
    reader = reader.to(reader_dtype)
    contriever = Contriever(BertConfig(**model_synth.CONTRIEVER_CFG, hidden_dropout_prob=0.0,
                                       attention_probs_dropout_prob=0.0)).eval()
    sd, _ = model_synth.fill_state_dict(contriever.state_dict(), seed=101)
    contriever.load_state_dict(sd)
    retriever = DualEncoderRetriever(opt, contriever)
    model = Atlas(opt, reader, retriever, reader_tok, retriever_tok).eval()
    corpus = atlas_synth.make_corpus()
    G = np.load(os.path.join(GOLDEN_DIR, "atlas_tiny.npz"))
    passages = [[corpus[int(i)] for i in row] for row in G["ret_ids"]]
    query, target = atlas_synth.make_batch()
    stats = {}
    model.retrieve = lambda *a, **k: (passages, None)
    reader_loss, retriever_loss = model(None, query, target, train_retriever=True, iter_stats=stats)
    (reader_loss + retriever_loss).backward()
    out[f"{tag}/reader_loss"] = np.float32(reader_loss.item())
    out[f"{tag}/retriever_loss"] = np.float32(retriever_loss.item())
    for prefix, module in (("reader", reader), ("retriever", retriever)):
        for name, p in module.named_parameters():
            if p.grad is None:
                continue
            g = p.grad.float().numpy()
            out[f"{tag}/norm/{prefix}.{name}"] = np.float32(np.linalg.norm(g))
            out[f"{tag}/proj/{prefix}.{name}"] = np.float32((g * direction(name, g.shape)).sum())
    print(tag, "reader_loss", reader_loss.item(), "retriever_loss", retriever_loss.item())


if __name__ == "__main__":
    ref_shims.install()
    torch.Tensor.cuda = lambda self, *a, **kw: self
    torch.set_num_threads(8)
    out = {}
    run(torch.float32, out, "fp32")
    try:
        run(torch.bfloat16, out, "bf16")
    except Exception as e:
        print("bf16 reader failed on CPU:", repr(e)[:300])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "atlas_train_tiny.npz"), **out)
    names = [k[10:] for k in out if k.startswith("fp32/norm/")]
    worst = max((abs(float(out[f"fp32/norm/{n}"]) - float(out.get(f"bf16/norm/{n}", out[f"fp32/norm/{n}"])))
                 / max(float(out[f"fp32/norm/{n}"]), 1e-30), n) for n in names)
    print(len(names), "gradients; worst relative norm drift of the reference's bf16-reader run:", worst)
