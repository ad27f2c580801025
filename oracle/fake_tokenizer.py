"""Deterministic stand-in for the HF tokenizers Atlas uses (BertTokenizer for Contriever, T5Tokenizer for FiD).

TEST INFRASTRUCTURE.  There is no network in the build container or on the GPU box, so the real vocabularies
are unavailable; both the golden generator (which drives the UNMODIFIED reference `src.atlas.Atlas`) and the
tests (which drive `atlas_b200.atlas.Atlas`) use this class, so the two sides see identical token ids.  It
implements the call surface the reference uses (src/atlas.py:26-39,69-76,187-246,412-426,621-636):
`tokenizer(texts, padding=, max_length=, truncation=, return_tensors="pt", add_special_tokens=)`,
`batch_encode_plus`, `.vocab`.
"""
import zlib

import torch


class FakeTokenizer:
    def __init__(self, kind, vocab_size):
        assert kind in ("bert", "t5")
        self.kind = kind
        self.vocab_size = vocab_size
        self.pad_token_id = 0
        self.eos_token_id = 1      # t5 "</s>"
        self.cls_token_id = 2      # bert [CLS]
        self.sep_token_id = 3      # bert [SEP]
        self.vocab = {f"tok{i}": i for i in range(vocab_size)}

    def _word(self, w):
        return 10 + zlib.crc32(w.encode()) % (self.vocab_size - 10)

    def _encode(self, text, add_special_tokens):
        ids = []
        for w in text.replace("</s>", " </s> ").split():
            ids.append(self.eos_token_id if w == "</s>" else self._word(w.lower()))
        if add_special_tokens:
            if self.kind == "bert":
                ids = [self.cls_token_id] + ids + [self.sep_token_id]
            else:
                ids = ids + [self.eos_token_id]
        return ids

    def __call__(self, texts, padding=False, max_length=None, truncation=False, return_tensors=None,
                 add_special_tokens=True):
        single = isinstance(texts, str)
        if single:
            texts = [texts]
        rows = [self._encode(t, add_special_tokens) for t in texts]
        if truncation and max_length is not None:
            out = []
            for r in rows:
                if len(r) > max_length:
                    if add_special_tokens and self.kind == "bert":
                        r = r[:max_length - 1] + [self.sep_token_id]
                    elif add_special_tokens:
                        r = r[:max_length - 1] + [self.eos_token_id]
                    else:
                        r = r[:max_length]
                out.append(r)
            rows = out
        if return_tensors is None:
            return {"input_ids": rows[0] if single else rows,
                    "attention_mask": [[1] * len(r) for r in rows][0] if single else [[1] * len(r) for r in rows]}
        if padding == "max_length":
            width = max_length
        elif padding in ("longest", True):
            width = max((len(r) for r in rows), default=0)
        else:
            width = max((len(r) for r in rows), default=0)
        ids = torch.full((len(rows), width), self.pad_token_id, dtype=torch.long)
        mask = torch.zeros((len(rows), width), dtype=torch.long)
        for i, r in enumerate(rows):
            ids[i, :len(r)] = torch.tensor(r, dtype=torch.long)
            mask[i, :len(r)] = 1
        return {"input_ids": ids, "attention_mask": mask}

    def batch_encode_plus(self, texts, **kw):
        return self(texts, **kw)

    def decode(self, ids, skip_special_tokens=True):
        return " ".join(f"tok{int(i)}" for i in ids if not (skip_special_tokens and int(i) < 10))
