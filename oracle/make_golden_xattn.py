"""Generate tests/golden/xattn_tiny.npz: the cross-attention score capture of the UNMODIFIED reference `src.fid.FiD`
(`overwrite_forward_crossattention`, `get_crossattention_scores`, src/fid.py:126-235,333-343) on CPU under
oracle/ref_shims.py, seeded weights / inputs (oracle/model_synth.py).

TEST INFRASTRUCTURE.  Stored: the three recorded maps of decoder layer 0 and of the last layer ([B, T, n*L]: head-means
of the logits, probabilities, ||V||-weighted probabilities), every aggregate of `get_crossattention_scores(mode="all")`
([B, n] each: {scores,probs,norms} x {top5,top10,top20,nosep,first,sum,avg,woquery}), in fp32 and for the reference's own
bf16-parameter run (accuracy budget)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import model_synth  # noqa: E402
import ref_shims  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    from transformers import T5Config
    from src.fid import FiD

    cfg = T5Config(**model_synth.T5_CFG)
    cfg.tie_word_embeddings = False
    ids, mask, labels, mask_query = model_synth.fid_inputs_with_sep()
    B, n_ctx = 2, 3
    out = {}
    sd = None
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        model = FiD(cfg).eval()
        if sd is None:
            sd, sha = model_synth.fill_state_dict(model.state_dict(), seed=202)
            out["weights_sha256"] = np.array(sha)
        model.load_state_dict(sd)
        model = model.to(dt)
        model.overwrite_forward_crossattention()
        model.create_crossattention_storage()
        model.encoder.config.n_context, model.encoder.config.bsz = n_ctx, B
        with torch.no_grad():
            res = model(input_ids=ids, attention_mask=mask, decoder_input_ids=model._shift_right(labels), labels=labels,
                        use_cache=False)
            agg = model.get_crossattention_scores(n_ctx, mask, labels=labels, ids=ids, mode="all", mask_query=mask_query)
        out[f"{name}/loss"] = np.float32(float(res[0]))
        for k, v in agg.items():
            out[f"{name}/agg/{k}"] = v.float().numpy()
        if name == "fp32":
            for li in (0, cfg.num_decoder_layers - 1):
                x = model.decoder.block[li].layer[1].EncDecAttention
                out[f"fp32/layer{li}/scores"] = x.score_storage.float().numpy()
                out[f"fp32/layer{li}/probs"] = x.prob_storage.float().numpy()
                out[f"fp32/layer{li}/norms"] = x.normalized_score_storage.float().numpy()
        print(name, "loss", float(res[0]), "keys", len(agg))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "xattn_tiny.npz"), **out)
    worst = 0.0
    for k in [k for k in out if k.startswith("fp32/agg/")]:
        a, b = out[k], out["bf16" + k[4:]]
        worst = max(worst, float(np.abs(a - b).max() / (np.abs(a).max() + 1e-12)))
    print("worst relative drift of the reference's bf16 aggregates:", worst)


if __name__ == "__main__":
    ref_shims.install()
    torch.manual_seed(0)
    main()
