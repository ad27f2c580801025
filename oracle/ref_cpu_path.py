"""The reference's CPU search path restated with the reference's own torch-CPU ops, for TIMING.

TEST/BENCH INFRASTRUCTURE ONLY (imported by bench.py's `cpu_baseline` / `--impl reference` legs and
by tests).  `/root/reference` cannot travel to the GPU box, so this file restates, line for line in
behaviour, what `DistributedIndex.search_knn` executes on one process with a CPU-resident bank
(src/index.py:113-157 with `torch.distributed` not initialised):

    scores = torch.matmul(allqueries.half(), self.embeddings)        # :117   [nq, n] fp16
    scores, indices = torch.topk(scores, topk, dim=1)                # :118
    indices = indices.tolist(); docs = [[doc_map[x] ...]]            # :132-133
    _, subindices = torch.topk(scores, topk, dim=1)                  # :151
    scores = scores.tolist(); subindices = subindices.tolist()       # :152-153
    scores/docs re-indexed in Python                                 # :155-156

`tests/test_oracle_golden.py::test_ref_cpu_path_matches_golden` pins it to the reference goldens.
"""
import torch


@torch.no_grad()
def reference_search_cpu(embeddings, doc_map, queries, topk):
    """embeddings: [768, n] fp16 CPU tensor (the reference layout); queries: [nq, 768] float."""
    scores = torch.matmul(queries.half(), embeddings)
    scores, indices = torch.topk(scores, topk, dim=1)
    indices = indices.tolist()
    docs = [[doc_map[x] for x in sample_indices] for sample_indices in indices]
    _, subindices = torch.topk(scores, topk, dim=1)
    scores = scores.tolist()
    subindices = subindices.tolist()
    scores = [[scores[k][j] for j in idx] for k, idx in enumerate(subindices)]
    docs = [[docs[k][j] for j in idx] for k, idx in enumerate(subindices)]
    return docs, scores


class LazyDocMap:
    """id -> synthetic passage dict without materialising millions of dicts (bench only)."""

    def __init__(self, n, base=0, stride=1):
        self.n, self.base, self.stride = n, base, stride

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = self.base + self.stride * int(i)
        return {"id": str(g), "title": f"t{g}", "text": f"passage {g}"}
