"""CPU: the numpy oracle (oracle/mips_oracle.py) against the fixtures produced by running the
UNMODIFIED reference index (oracle/make_golden.py -> tests/golden/mips_*.npz)."""
import glob
import os

import numpy as np
import pytest

import mips_oracle
from conftest import GOLDEN_DIR, golden_inputs, load_golden

CASES = sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "mips_*.npz")))


def test_goldens_present():
    assert {"c1_grid", "c1_gauss", "w4_grid_empty_rank"} <= set(CASES)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    g = load_golden(name)
    bank, q, nq_per_rank = golden_inputs(g)
    k = int(g["k"])
    off = np.cumsum([0] + nq_per_rank)
    res = mips_oracle.search_knn_oracle(bank, [q[off[i]:off[i + 1]] for i in range(len(nq_per_rank))], k)
    vals = np.concatenate([v for v, _ in res])
    ids = np.concatenate([i for _, i in res])
    ref_scores, ref_ids, canon_ids = g["ref_scores"], g["ref_ids"], g["canon_ids"]
    if str(g["dist"]) == "grid":
        # exact-grid inputs: accumulation order cannot matter -> bit-exact values, canonical ids
        assert np.array_equal(vals.view(np.uint16), ref_scores.view(np.uint16))
        assert np.array_equal(ids, canon_ids)
    else:
        # realistic inputs: BLAS accumulation order may flip an fp16 rounding -> <= 1 ulp
        a = vals.astype(np.float32)
        b = ref_scores.astype(np.float32)
        assert np.all(np.abs(a - b) <= np.spacing(np.abs(b).astype(np.float16)).astype(np.float32))
    # the reference's own pick agrees wherever the tie order / k-boundary cannot matter
    for r in range(vals.shape[0]):
        if str(g["dist"]) == "grid":
            assert mips_oracle.ids_match_tie_aware(ref_scores[r], ref_ids[r], ids[r])


def test_canonical_topk_tie_rule():
    s = np.array([[1.0, 2.0, 2.0, -0.0, 0.0, 2.0]], dtype=np.float16)
    v, i = mips_oracle.canonical_topk(s, 5)
    assert i.tolist() == [[1, 2, 5, 0, 3]]
    with pytest.raises(RuntimeError):
        mips_oracle.canonical_topk(s, 7)


def test_sharded_equals_single():
    import synth

    bank = synth.make_bank(999, seed=3)
    q = synth.make_queries(6, seed=4)
    single = mips_oracle.search_knn_oracle(bank, [q], 11)[0]
    multi = mips_oracle.search_knn_oracle(bank, [q[:2], q[2:2], q[2:]], 11)
    assert np.array_equal(np.concatenate([m[1] for m in multi]), single[1])
    assert np.array_equal(np.concatenate([m[0] for m in multi]).view(np.uint16), single[0].view(np.uint16))


def test_ref_cpu_path_matches_golden():
    """The torch-CPU restatement used for the timed CPU baseline returns the reference's values."""
    import torch

    import ref_cpu_path
    import synth

    g = load_golden("c1_grid")
    bank, q, _ = golden_inputs(g)
    docs, scores = ref_cpu_path.reference_search_cpu(
        torch.from_numpy(bank).T.contiguous(), ref_cpu_path.LazyDocMap(bank.shape[0]), torch.from_numpy(q), int(g["k"]))
    got = np.array(scores, dtype=np.float32).astype(np.float16)
    assert np.array_equal(got.view(np.uint16), g["ref_scores"].view(np.uint16))
    ids = np.array([[int(d["id"]) for d in row] for row in docs])
    for r in range(ids.shape[0]):
        assert mips_oracle.ids_match_tie_aware(g["ref_scores"][r], g["ref_ids"][r], ids[r])


def test_atlas_golden_retrieval_is_the_oracle_topk():
    """The Atlas-level golden (reference `Atlas.retrieve` on the reference `DistributedIndex`, oracle/make_golden_atlas.py)
    pins the oracle once more: its returned scores are the canonical top-k of its own fp16 score matrix."""
    g = np.load(os.path.join(GOLDEN_DIR, "atlas_tiny.npz"))
    full = g["all_scores_fp16"].astype(np.float16)
    vals, ids = mips_oracle.canonical_topk(full, g["ret_ids"].shape[1])
    assert np.array_equal(vals.astype(np.float32), g["ret_scores"])
    for r in range(ids.shape[0]):
        uniq = np.array([np.sum(full[r] == v) == 1 for v in vals[r]])
        assert np.array_equal(ids[r][uniq], g["ret_ids"][r][uniq])
    # and the reference bank is the fp16 rounding of embeddings close to the fp32 retriever's
    assert np.abs(g["bank_fp16"].astype(np.float32) - g["bank_fp32"]).max() < 1e-2


def test_fake_tokenizer_surface():
    import atlas_synth

    rt, bt = atlas_synth.tokenizers()
    enc = bt(["a b c", "d"], padding="longest", return_tensors="pt", max_length=8, truncation=True)
    assert enc["input_ids"].shape == (2, 5) and enc["attention_mask"].sum().item() == 8
    assert enc["input_ids"][0, 0].item() == bt.cls_token_id and enc["input_ids"][1, 2].item() == bt.sep_token_id
    enc = rt(["x y </s>"], padding="max_length", max_length=6, truncation=True, return_tensors="pt", add_special_tokens=False)
    assert enc["input_ids"].shape == (1, 6) and enc["input_ids"][0, 2].item() == rt.eos_token_id
    assert rt(["p q r s t u v"], max_length=4, truncation=True, return_tensors="pt")["input_ids"][0, -1].item() == 1
    assert len(rt.vocab) == atlas_synth.READER_VOCAB


def _named_state(kind, seed):
    """The seeded weights of oracle/model_synth.py under the HF parameter names (no module needed)."""
    import torch

    import model_synth
    from atlas_b200.fid import FiD, T5ConfigLite
    from atlas_b200.retrievers import BertConfigLite, Contriever

    if kind == "fid":
        m = FiD(T5ConfigLite(**{k: v for k, v in model_synth.T5_CFG.items()
                                if k not in ("dropout_rate", "is_encoder_decoder", "use_cache")}))
    else:
        m = Contriever(BertConfigLite(**model_synth.CONTRIEVER_CFG))
    sd, sha = model_synth.fill_state_dict(m.state_dict(), seed)
    return {k: v.float() if torch.is_floating_point(v) else v for k, v in sd.items()}, sha


def test_fid_cpu_restatement_matches_reference():
    """oracle/fid_cpu.py (the CPU checker / timed CPU baseline on the GPU box) reproduces the UNMODIFIED reference
    FiD's fp32 logits, loss and encoder states on the golden case (oracle/make_golden_models.py)."""
    import torch

    import fid_cpu
    import model_synth

    g = np.load(os.path.join(GOLDEN_DIR, "fid_tiny.npz"))
    sd, sha = _named_state("fid", 202)
    assert sha == str(g["weights_sha256"])
    ids, mask, labels = model_synth.fid_inputs()
    dec = labels.clone()
    dec = torch.cat([torch.zeros(dec.shape[0], 1, dtype=dec.dtype), dec[:, :-1]], dim=1).masked_fill_(
        torch.cat([torch.zeros(dec.shape[0], 1, dtype=torch.bool), labels[:, :-1] == -100], dim=1), 0)   # _shift_right
    with torch.no_grad():
        loss, logits, enc = fid_cpu.fid_forward(sd, model_synth.T5_CFG, ids, mask, dec, labels, n_context=3)
    assert np.abs(logits.numpy() - g["logits_fp32"]).max() < 2e-4
    assert abs(float(loss) - float(g["loss_fp32"])) < 1e-4
    assert np.abs(enc.numpy() - g["enc_fp32"].astype(np.float32)).max() < 2e-3   # golden stored in fp16


def test_contriever_cpu_restatement_matches_reference():
    import torch

    import fid_cpu
    import model_synth

    g = np.load(os.path.join(GOLDEN_DIR, "contriever_tiny.npz"))
    sd, sha = _named_state("contriever", 101)
    assert sha == str(g["weights_sha256"])
    ids, mask = model_synth.contriever_inputs()
    with torch.no_grad():
        emb = fid_cpu.contriever_forward(sd, model_synth.CONTRIEVER_CFG, ids, mask)
    assert np.abs(emb.numpy() - g["emb_fp32"]).max() < 2e-4


# ---- gradients: the autograd of the CPU restatements against the reference's own loss.backward() -------------------
def _shift_right(labels):
    import torch

    s = labels.new_zeros(labels.shape)
    s[..., 1:] = labels[..., :-1].clone()
    s[..., 0] = 0
    return s.masked_fill(s == -100, 0)


@pytest.mark.parametrize("which", ["fid", "contriever"])
def test_gradient_oracle_matches_reference(which):
    import torch

    import fid_cpu
    import grad_oracle
    import model_synth

    g = np.load(os.path.join(GOLDEN_DIR, "grads_tiny.npz"))
    if which == "fid":
        from atlas_b200.fid import FiD, T5ConfigLite

        cfg = {k: v for k, v in model_synth.T5_CFG.items() if k not in ("dropout_rate", "is_encoder_decoder", "use_cache")}
        sd, sha = model_synth.fill_state_dict(FiD(T5ConfigLite(**cfg)).state_dict(), 202)
        ids, mask, labels = model_synth.fid_inputs()
        loss, grads = grad_oracle.fid_grads(sd, model_synth.T5_CFG, ids, mask, labels, 3, _shift_right)
    else:
        from atlas_b200.retrievers import BertConfigLite, Contriever

        sd, sha = model_synth.fill_state_dict(Contriever(BertConfigLite(**model_synth.CONTRIEVER_CFG)).state_dict(), 101)
        ids, mask = model_synth.contriever_inputs()
        loss, grads = grad_oracle.contriever_grads(sd, model_synth.CONTRIEVER_CFG, ids, mask)
    assert sha == str(g[f"{which}/weights_sha256"])
    assert abs(loss - float(g[f"{which}_fp32/loss"])) <= 2e-5 * abs(loss)
    names = [n[len(which) + 11:] for n in g.files if n.startswith(f"{which}_fp32/norm/")]
    assert len(names) >= 30 and set(names) <= set(grads)
    top = max(float(g[f"{which}_fp32/norm/{n}"]) for n in names)
    for n in names:
        gr = grads[n].numpy()
        ref_norm = float(g[f"{which}_fp32/norm/{n}"])
        if ref_norm < 1e-6 * top:            # mathematically zero gradients (BERT key bias): noise in the reference too
            assert np.linalg.norm(gr) < 1e-5 * top
            continue
        assert abs(np.linalg.norm(gr) - ref_norm) <= 2e-4 * ref_norm, n
        proj = float((gr * grad_oracle.direction(n, gr.shape)).sum())
        assert abs(proj - float(g[f"{which}_fp32/proj/{n}"])) <= 2e-3 * ref_norm, n
        key = f"{which}_fp32/full/{n}"
        if key in g.files:
            assert np.abs(gr - g[key]).max() <= 2e-4 * np.abs(g[key]).max() + 1e-7, n


def test_crossattention_aggregates_match_reference():
    """Host logic of `FiD.get_crossattention_scores` / `aggregate_value` / `get_topk_score` / `get_woquery_score`
    (atlas_b200/fid.py, pure tensor reductions) fed with the maps the UNMODIFIED reference recorded
    (tests/golden/xattn_tiny.npz, oracle/make_golden_xattn.py): all 24 aggregates."""
    import torch

    import model_synth
    from atlas_b200.fid import FiD, T5ConfigLite

    g = np.load(os.path.join(GOLDEN_DIR, "xattn_tiny.npz"))
    cfg = {k: v for k, v in model_synth.T5_CFG.items() if k not in ("dropout_rate", "is_encoder_decoder", "use_cache")}
    model = FiD(T5ConfigLite(**cfg))
    ids, mask, labels, mask_query = model_synth.fid_inputs_with_sep()
    model._xattn = [tuple(torch.from_numpy(g[f"fp32/layer{li}/{k}"]) for k in ("scores", "probs", "norms"))
                    for li in range(cfg["num_decoder_layers"])]
    agg = model.get_crossattention_scores(3, mask, labels=labels, ids=ids, mode="all", mask_query=mask_query)
    keys = [k[9:] for k in g.files if k.startswith("fp32/agg/")]
    assert len(keys) == 24 and set(keys) == set(agg)
    for k in keys:
        ref = g[f"fp32/agg/{k}"]
        assert np.abs(agg[k].numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), k
