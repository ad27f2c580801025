"""torchrun worker for tests/test_search_gpu.py::test_distributed_search_nccl (one rank per GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import mips_oracle  # noqa: E402
import synth  # noqa: E402
from atlas_b200.index import DistributedIndex  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    n, k = 50001, 40
    sizes = [5, 0, 9, 3, 4, 4, 4, 4][:world]
    bank = synth.make_bank(n, seed=91)
    qs = [synth.make_queries(m, seed=100 + r) for r, m in enumerate(sizes)]
    want = mips_oracle.search_knn_oracle(bank, qs, k)
    index = DistributedIndex()
    index.init_embeddings(synth.make_passages(n, rank, world))
    rows = mips_oracle.shard_rows(n, rank, world)
    index.embeddings[:, :] = torch.from_numpy(bank[rows]).T.cuda()
    for capacity in (None, 16):          # size-exchange protocol, then the fixed per-rank capacity (no host sync)
        index.max_queries_per_rank = capacity
        for _ in range(2):
            docs, scores = index.search_knn(torch.from_numpy(qs[rank]).cuda(), k)
        ids = np.array([[int(d["id"]) for d in row] for row in docs], dtype=np.int64).reshape(-1, k)
        vals = np.array(scores, dtype=np.float32).reshape(-1, k).astype(np.float16)
        assert np.array_equal(ids, want[rank][1]), f"rank {rank}: ids differ (capacity {capacity})"
        assert np.array_equal(vals.view(np.uint16), want[rank][0].view(np.uint16)), f"rank {rank}: scores differ"
    index._reset_store()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("nccl distributed search ok")


if __name__ == "__main__":
    main()
