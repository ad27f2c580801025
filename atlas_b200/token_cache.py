"""Retriever-side passage token cache for index refresh.

`Atlas.build_index` (src/atlas.py:61-88) formats and tokenises EVERY local passage again on every index refresh
(`fstr.format(**p)` + a tokenizer call per 512-passage batch, src/atlas.py:69-76): 4 Mi tokenizer calls' worth of host
work per refresh per GPU at BASELINE configs[2], while the GPU embeds a batch in milliseconds.  Passages do not change
between refreshes - only the retriever's weights do - so the token ids are computed once and kept as one int32 matrix
`ids [N, Lmax]` + `lens [N]` (pinned host memory by default, or on the device); a refresh then slices batches out of
it.  The batches are IDENTICAL to what the tokenizer call of the reference produces (`padding="longest"`: a batch is cut
at its longest passage; `truncation=True, max_length`), which `tests/test_token_cache.py` checks tensor for tensor.
"""
import torch


class RetrieverTokenCache:
    def __init__(self, n, max_len, pad_id=0, device="cpu"):
        self.n, self.max_len, self.pad_id = int(n), int(max_len), int(pad_id)
        self.device = torch.device(device)
        pin = self.device.type == "cpu" and torch.cuda.is_available()
        self.ids = torch.full((self.n, self.max_len), self.pad_id, dtype=torch.int32, device=self.device,
                              pin_memory=pin)
        self.lens = torch.zeros(self.n, dtype=torch.int32, device=self.device, pin_memory=pin)
        self.filled = 0

    def nbytes(self):
        return self.ids.numel() * 4 + self.lens.numel() * 4

    def append(self, input_ids, attention_mask):
        """Store one tokenised batch (`padding="longest"` output of the tokenizer, any device)."""
        b, l = input_ids.shape
        if self.filled + b > self.n or l > self.max_len:
            raise ValueError(f"token cache overflow: batch {tuple(input_ids.shape)} into [{self.n}, {self.max_len}] at {self.filled}")
        rows = slice(self.filled, self.filled + b)
        self.ids[rows, :l] = input_ids.to(self.device, torch.int32)
        self.lens[rows] = attention_mask.sum(dim=1).to(self.device, torch.int32)
        # the cache assumes right padding (what BERT / T5 tokenizers do): the mask must be a prefix of ones
        if not bool((attention_mask.long().cumprod(dim=1).sum(dim=1) == attention_mask.long().sum(dim=1)).all()):
            raise ValueError("token cache needs right-padded batches")
        self.filled += b

    @property
    def complete(self):
        return self.filled == self.n

    def batch(self, start, stop, device=None):
        """(input_ids int64 [b, L], attention_mask int64 [b, L]) of passages [start, stop), padded to the longest
        passage of the batch - the tensors `tokenizer(texts, padding="longest", truncation=True, max_length=...)` returns."""
        lens = self.lens[start:stop]
        longest = int(lens.max()) if stop > start else 0
        ids = self.ids[start:stop, :longest]
        if device is not None:
            ids, lens = ids.to(device, non_blocking=True), lens.to(device, non_blocking=True)
        mask = (torch.arange(longest, device=ids.device)[None, :] < lens[:, None]).to(torch.int64)
        return ids.to(torch.int64), mask


def cache_key(passages, max_len, fmt):
    """Identity of a passage shard for cache reuse: the list object, its length, its first / last passage ids, the format
    and the truncation length."""
    def pid(p):
        return p.get("id") if isinstance(p, dict) else None

    ends = (pid(passages[0]), pid(passages[-1])) if len(passages) else (None, None)
    return (id(passages), len(passages), ends, int(max_len), fmt)


def fits(n, max_len, max_bytes):
    """Whether an [n, max_len] int32 record stays under the configured budget (default 8 GiB of pinned host memory:
    4 Mi passages x 384 tokens = 6.4 GB)."""
    return n * max_len * 4 + n * 4 <= max_bytes
