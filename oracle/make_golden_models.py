"""Generate tests/golden/{contriever,fid}_tiny.npz by running the UNMODIFIED reference modules
(`src.retrievers.Contriever`, `src.fid.FiD`) on CPU under oracle/ref_shims.py with seeded weights/inputs
(oracle/model_synth.py).  Stored: fp32 outputs and, for the accuracy budget, the reference's own outputs
when its parameters are cast to bf16 / fp16 (how far the reference itself drifts at 16 bits)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import model_synth  # noqa: E402
import ref_shims  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


def contriever():
    from transformers import BertConfig
    from src.retrievers import Contriever

    cfg = BertConfig(**model_synth.CONTRIEVER_CFG)
    model = Contriever(cfg).eval()
    sd, sha = model_synth.fill_state_dict(model.state_dict(), seed=101)
    model.load_state_dict(sd)
    ids, mask = model_synth.contriever_inputs()
    out = {}
    with torch.no_grad():
        out["emb_fp32"] = model(input_ids=ids, attention_mask=mask).float().numpy()
        for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
            try:
                m16 = Contriever(cfg).eval()
                m16.load_state_dict(sd)
                m16 = m16.to(dt)
                out[f"emb_{name}"] = m16(input_ids=ids, attention_mask=mask).float().numpy()
            except Exception as e:  # some half ops may be missing on CPU
                print("contriever", name, "failed on CPU:", repr(e)[:200])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "contriever_tiny.npz"), weights_sha256=sha, **out)
    print("contriever:", {k: v.shape for k, v in out.items()},
          {k: float(np.abs(v - out["emb_fp32"]).max()) for k, v in out.items()})


def fid():
    from transformers import T5Config
    from src.fid import FiD

    cfg = T5Config(**model_synth.T5_CFG)
    # transformers 5.x drops the kwarg; T5 v1.1 checkpoints (google/t5-*-lm-adapt) have untied heads
    cfg.tie_word_embeddings = False
    assert cfg.tie_word_embeddings is False
    ids, mask, labels = model_synth.fid_inputs()
    B, n_ctx = 2, 3
    out = {}
    sd = sha = None
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16), ("fp16", torch.float16)):
        try:
            model = FiD(cfg).eval()
            if sd is None:
                sd, sha = model_synth.fill_state_dict(model.state_dict(), seed=202)
            model.load_state_dict(sd)
            model = model.to(dt)
            model.encoder.config.n_context = n_ctx
            model.encoder.config.bsz = B
            with torch.no_grad():
                dec_in = model._shift_right(labels)
                res = model(input_ids=ids, attention_mask=mask, decoder_input_ids=dec_in, labels=labels, use_cache=False)
            out[f"loss_{name}"] = np.array(float(res[0]))
            out[f"logits_{name}"] = res[1].float().numpy()
            if name == "fp32":  # stored in fp16 to keep the fixture small (values are O(1))
                out["enc_fp32"] = res.encoder_last_hidden_state.float().numpy().astype(np.float16)
        except Exception as e:
            print("fid", name, "failed on CPU:", repr(e)[:300])
    np.savez_compressed(os.path.join(GOLDEN_DIR, "fid_tiny.npz"), weights_sha256=sha, **out)
    print("fid:", {k: v.shape for k, v in out.items()})
    for name in ("bf16", "fp16"):
        if f"logits_{name}" in out:
            print(f"  reference {name} vs fp32: logits max abs diff",
                  float(np.abs(out[f"logits_{name}"] - out["logits_fp32"]).max()), "loss", float(out[f"loss_{name}"]),
                  float(out["loss_fp32"]))


if __name__ == "__main__":
    ref_shims.install()
    torch.manual_seed(0)
    contriever()
    fid()
