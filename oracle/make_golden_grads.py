"""Generate tests/golden/grads_tiny.npz: gradients of the UNMODIFIED reference modules (`src.fid.FiD`,
`src.retrievers.Contriever`) on CPU under oracle/ref_shims.py, seeded weights / inputs (oracle/model_synth.py).

TEST INFRASTRUCTURE.  The training step of the reference is `loss.backward()` through these modules
(train.py -> Atlas.forward, src/atlas.py:399-550); dropout is off (eval mode, gradients enabled) so the result
is deterministic.  Stored per parameter: the fp32 gradient's L2 norm and its projection on a seeded random
direction (cheap, order-independent fingerprints of every gradient), the full gradient of the small parameters
(norm weights, biases, relative-attention-bias tables), and the same fingerprints of the reference's own run
with bf16 parameters (how far the reference itself drifts at 16 bits: the accuracy budget of the GPU tests)."""
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import model_synth  # noqa: E402
import ref_shims  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
SMALL = 4096   # parameters up to this many elements are stored in full


def direction(name, shape):
    rng = np.random.default_rng([7, zlib.crc32(name.encode())])
    return rng.standard_normal(shape, dtype=np.float32)


def fingerprints(model, prefix, out, full=True):
    for name, p in model.named_parameters():
        g = p.grad
        if g is None:
            continue
        g = g.float().numpy()
        out[f"{prefix}/norm/{name}"] = np.float32(np.linalg.norm(g))
        out[f"{prefix}/proj/{name}"] = np.float32((g * direction(name, g.shape)).sum())
        if full and g.size <= SMALL:
            out[f"{prefix}/full/{name}"] = g


def fid(out):
    from transformers import T5Config
    from src.fid import FiD

    cfg = T5Config(**model_synth.T5_CFG)
    cfg.tie_word_embeddings = False
    ids, mask, labels = model_synth.fid_inputs()
    B, n_ctx = 2, 3
    sd = None
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        try:
            model = FiD(cfg).eval()     # eval: dropout off; gradients stay enabled
            if sd is None:
                sd, sha = model_synth.fill_state_dict(model.state_dict(), seed=202)
                out["fid/weights_sha256"] = np.array(sha)
            model.load_state_dict(sd)
            model = model.to(dt)
            model.encoder.config.n_context = n_ctx
            model.encoder.config.bsz = B
            res = model(input_ids=ids, attention_mask=mask, decoder_input_ids=model._shift_right(labels), labels=labels,
                        use_cache=False)
            res[0].backward()
            out[f"fid_{name}/loss"] = np.float32(float(res[0]))
            fingerprints(model, f"fid_{name}", out, full=(name == "fp32"))
            print("fid", name, "loss", float(res[0]))
        except Exception as e:
            print("fid", name, "failed on CPU:", repr(e)[:300])


def contriever(out):
    from transformers import BertConfig
    from src.retrievers import Contriever

    cfg = BertConfig(**model_synth.CONTRIEVER_CFG, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    ids, mask = model_synth.contriever_inputs()
    sd = None
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        try:
            model = Contriever(cfg).eval()
            if sd is None:
                sd, sha = model_synth.fill_state_dict(model.state_dict(), seed=101)
                out["contriever/weights_sha256"] = np.array(sha)
            model.load_state_dict(sd)
            model = model.to(dt)
            emb = model(input_ids=ids, attention_mask=mask)
            w = torch.from_numpy(direction("emb", tuple(emb.shape)))
            loss = (emb.float() * w).sum() / emb.shape[0]     # a fixed linear functional of the embeddings
            loss.backward()
            out[f"contriever_{name}/loss"] = np.float32(float(loss))
            fingerprints(model, f"contriever_{name}", out, full=(name == "fp32"))
            print("contriever", name, "loss", float(loss))
        except Exception as e:
            print("contriever", name, "failed on CPU:", repr(e)[:300])


if __name__ == "__main__":
    ref_shims.install()
    torch.manual_seed(0)
    out = {}
    fid(out)
    contriever(out)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "grads_tiny.npz"), **out)
    print(len(out), "entries")
    for k in ("fid", "contriever"):
        names = [n[len(k) + 11:] for n in out if n.startswith(f"{k}_fp32/norm/")]
        worst = 0.0
        for n in names:
            a, b = out[f"{k}_fp32/norm/{n}"], out.get(f"{k}_bf16/norm/{n}")
            if b is not None and a > 0:
                worst = max(worst, abs(float(a) - float(b)) / float(a))
        print(k, "params", len(names), "worst relative norm drift of the reference's bf16 run:", worst)
