"""Seeded synthetic corpus / batch / options for the Atlas-level goldens (tests/golden/atlas_tiny.npz).

TEST INFRASTRUCTURE shared by `oracle/make_golden_atlas.py` (drives the UNMODIFIED reference `src.atlas.Atlas`
on CPU) and `tests/test_atlas_gpu.py` (drives `atlas_b200.atlas.Atlas` on the GPU)."""
from types import SimpleNamespace

import numpy as np

import model_synth
from fake_tokenizer import FakeTokenizer

N_PASSAGES = 96
N_QUERIES = 3
TOPK = 4
WORDS = [f"w{i}" for i in range(400)]

READER_VOCAB = model_synth.T5_CFG["vocab_size"]          # 512
RETRIEVER_VOCAB = model_synth.CONTRIEVER_CFG["vocab_size"]  # 2000


def make_opt(**over):
    """The option fields `Atlas` reads (defaults of src/options.py, sizes shrunk)."""
    d = dict(retriever_format="{title} {text}", encoder_format="{query} title: {title} context: {text}",
             text_maxlength=64, target_maxlength=8, retriever_n_context=TOPK, n_context=TOPK,
             filtering_overretrieve_ratio=2, retrieve_with_rerank=False, n_to_rerank_with_retrieve_with_rerank=8,
             per_gpu_embedder_batch_size=32, decoder_prompt_format=None, decoder_format=None, use_file_passages=False,
             gold_score_mode="ppmean", use_gradient_checkpoint_retriever=False, use_gradient_checkpoint_reader=False,
             compute_crossattention_stats=False, temperature_gold=0.01, temperature_score=0.01,
             generation_max_length=8, generation_min_length=1, generation_num_beams=1, generation_length_penalty=1.0,
             query_side_retriever_training=False)
    d.update(over)
    return SimpleNamespace(**d)


def make_corpus(seed=77):
    rng = np.random.default_rng(seed)
    passages = []
    for i in range(N_PASSAGES):
        n = int(rng.integers(8, 40))
        passages.append({"id": str(i), "title": " ".join(rng.choice(WORDS, 2)), "text": " ".join(rng.choice(WORDS, n))})
    return passages


def make_batch(seed=78):
    rng = np.random.default_rng(seed)
    query = [" ".join(rng.choice(WORDS, int(rng.integers(4, 9)))) for _ in range(N_QUERIES)]
    target = [" ".join(rng.choice(WORDS, int(rng.integers(2, 6)))) for _ in range(N_QUERIES)]
    return query, target


def tokenizers():
    return FakeTokenizer("t5", READER_VOCAB), FakeTokenizer("bert", RETRIEVER_VOCAB)
