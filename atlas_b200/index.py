"""Drop-in for `src.index.DistributedIndex` (reference src/index.py:43-160) on B200.

Same public surface: `embeddings` ([768, N] fp16 CUDA, slice-assignable on dim 1),
`doc_map`, `is_in_gpu`, `init_embeddings`, `search_knn`, `save_index`, `load_index`,
`is_index_trained`, `train_index`.  What changes underneath:

  * the bank is stored passage-major, `[N, 768]` row-major (1536 contiguous bytes per passage, what
    TMA / tcgen05 want); `embeddings` is the transposed VIEW, so `index.embeddings[:, a:b] = emb.T`
    (src/atlas.py:79) and `embeddings[:, s:e]` (src/index.py:85) keep working and write through;
  * `_compute_scores_and_indices` is ONE fused scan (csrc/mips.cu) instead of matmul + topk: the
    [nq, N] score matrix is never materialised;
  * `search_knn` exchanges (fp16 score, int64 global id) pairs - one fixed-shape all-gather - instead
    of pickled passage dicts through 4*W var-size gathers; passage text is resolved from a node-shared
    store (atlas_b200/passage_store.py);
  * ties are ordered canonically: score descending, then global passage id ascending.

There is no CPU path: `is_in_gpu = False` (the reference's FAISS-only mode) raises.
"""
import ctypes
import math
import os
import pickle
from typing import Optional

import numpy as np
import torch

from . import dist_utils, ops
from ._lib import MAX_TOPK, AtlasB200Error, check, lib
from .passage_store import make_store

EMBEDDINGS_DIM: int = 768  # src/retrievers.py:13


class DistributedIndex(object):
    def __init__(self):
        self._bank = None  # [N_local, 768] fp16, CUDA
        self.doc_map = dict()
        self.is_in_gpu = True
        self._store = None
        self._workspace = ops.Workspace()
        # global id of local row l is _id_base + _id_stride * l
        self._id_base, self._id_stride = 0, 1
        self._offsets = None  # contiguous layout (after load_index): first global id of every rank
        # Upper bound of the queries ANY rank passes to one search (e.g. --per_gpu_batch_size, identical on all ranks).
        # When set, a distributed search pads every rank's block to this many rows: no size exchange, no host
        # synchronisation in front of the query all-gather (src/index.py:127-129 does a var-size gather + a size gather).
        self.max_queries_per_rank = None

    # ------------------------------------------------------------------ bank / reference view
    @property
    def embeddings(self):
        """The reference's `[768, N]` tensor (src/index.py:51), as a view of the passage-major bank."""
        return None if self._bank is None else self._bank.t()

    @embeddings.setter
    def embeddings(self, value):
        if value is None:
            self._bank = None
            return
        if value.dim() != 2 or value.shape[0] != EMBEDDINGS_DIM:
            raise ValueError(f"embeddings must be [{EMBEDDINGS_DIM}, N], got {tuple(value.shape)}")
        self._bank = value.t().contiguous().to(device=self._device(), dtype=torch.float16)

    @staticmethod
    def _device():
        if not torch.cuda.is_available():
            raise AtlasB200Error("atlas_b200.DistributedIndex needs a CUDA device (no CPU fallback)")
        return torch.device("cuda", torch.cuda.current_device())

    def init_embeddings(self, passages, dim: Optional[int] = EMBEDDINGS_DIM):
        """src/index.py:49-53.  Local row l holds global passage line l*W + rank (src/index_io.py:41)."""
        if dim != EMBEDDINGS_DIM:
            raise ValueError(f"only dim={EMBEDDINGS_DIM} is supported (src/retrievers.py:13)")
        if not self.is_in_gpu:
            raise AtlasB200Error("is_in_gpu=False (CPU-resident bank) is a FAISS-only mode and is not supported")
        self.doc_map = {i: doc for i, doc in enumerate(passages)}
        self._bank = torch.zeros(len(passages), dim, dtype=torch.float16, device=self._device())
        self._id_base, self._id_stride = dist_utils.get_rank(), dist_utils.get_world_size()
        self._offsets = None
        self._reset_store()

    def _reset_store(self):
        if self._store is not None:
            self._store.close()
        self._store = None

    def _get_store(self):
        if self._store is None:
            self._store = make_store(self.doc_map, dist_utils.get_rank(), dist_utils.get_world_size())
        return self._store

    # ------------------------------------------------------------------ persistence (reference format)
    def _get_saved_embedding_path(self, save_dir: str, shard: int) -> str:
        return os.path.join(save_dir, f"embeddings.{shard}.pt")

    def _get_saved_passages_path(self, save_dir: str, shard: int) -> str:
        return os.path.join(save_dir, f"passages.{shard}.pt")

    def save_index(self, path: str, total_saved_shards: int, overwrite_saved_passages: bool = False) -> None:
        """Same files as src/index.py:61-87: `embeddings.{s}.pt` = [768, n_s] fp16, `passages.{s}.pt`."""
        assert self._bank is not None
        rank = dist_utils.get_rank()
        ws = dist_utils.get_world_size()
        assert total_saved_shards % ws == 0, f"N workers must be a multiple of shards to save"
        shards_per_worker = total_saved_shards // ws
        n_embeddings = self._bank.shape[0]
        embeddings_per_shard = math.ceil(n_embeddings / shards_per_worker)
        assert n_embeddings == len(self.doc_map), len(self.doc_map)
        for shard_ind, shard_start in enumerate(range(0, n_embeddings, embeddings_per_shard)):
            shard_end = min(shard_start + embeddings_per_shard, n_embeddings)
            shard_id = shard_ind + rank * shards_per_worker
            passage_shard_path = self._get_saved_passages_path(path, shard_id)
            if not os.path.exists(passage_shard_path) or overwrite_saved_passages:
                passage_shard = [self.doc_map[i] for i in range(shard_start, shard_end)]
                with open(passage_shard_path, "wb") as fobj:
                    pickle.dump(passage_shard, fobj, protocol=pickle.HIGHEST_PROTOCOL)
            embeddings_shard = self._bank[shard_start:shard_end].t().contiguous()  # [768, n_s]
            torch.save(embeddings_shard, self._get_saved_embedding_path(path, shard_id))

    def load_index(self, path: str, total_saved_shards: int):
        """src/index.py:89-111.  Rank r holds shard files r*S/W .. (r+1)*S/W-1 concatenated; global ids
        are then contiguous per rank (offsets exchanged once here)."""
        rank = dist_utils.get_rank()
        ws = dist_utils.get_world_size()
        assert total_saved_shards % ws == 0, f"N workers must be a multiple of shards to save"
        shards_per_worker = total_saved_shards // ws
        passages = []
        shard_ids = list(range(rank * shards_per_worker, (rank + 1) * shards_per_worker))
        for shard_id in shard_ids:
            with open(self._get_saved_passages_path(path, shard_id), "rb") as fobj:
                passages.append(pickle.load(fobj))
        self.doc_map = {}
        n_passages = 0
        for chunk in passages:
            for p in chunk:
                self.doc_map[n_passages] = p
                n_passages += 1
        device = self._device()
        files = [self._get_saved_embedding_path(path, s) for s in shard_ids]
        if device.type == "cuda":
            self._bank = self._load_bank_streamed(files, device)
        else:       # CPU test doubles of this class (tests/test_index_gloo.py)
            rows = [torch.load(f, map_location="cpu").t().to(torch.float16) for f in files]
            self._bank = torch.cat(rows, dim=0).contiguous().to(device)
        sizes = dist_utils.get_varsize(self._bank)
        self._offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        self._id_base, self._id_stride = int(self._offsets[rank]), 1
        self._reset_store()

    @staticmethod
    def _load_bank_streamed(files, device, chunk_cols=262144):
        """The shard files (`embeddings.{s}.pt` = [768, n_s] fp16, src/index.py:75-87) straight into the device bank
        [sum n_s, 768]: every file is memory-mapped (`torch.load(mmap=True)`: no host copy of the whole tensor), column
        chunks are staged through two pinned buffers, copied H2D asynchronously and transposed into their bank rows by
        `atlas_b200_transpose` on the GPU.  The reference (and round 1 here) reads every shard into host memory,
        transposes it on ONE CPU core (`.t()` + `cat` -> a strided 2-byte copy, ~1 GB/s) and only then uploads it."""
        shards = []
        for f in files:
            try:
                t = torch.load(f, map_location="cpu", mmap=True)
            except (RuntimeError, ValueError, TypeError):     # legacy (non-zip) checkpoint: no mmap
                t = torch.load(f, map_location="cpu")
            if t.dim() != 2 or t.shape[0] != EMBEDDINGS_DIM:
                raise AtlasB200Error(f"{f}: expected a [{EMBEDDINGS_DIM}, n] embedding shard, got {tuple(t.shape)}")
            shards.append(t)
        total = sum(int(t.shape[1]) for t in shards)
        bank = torch.empty((total, EMBEDDINGS_DIM), dtype=torch.float16, device=device)
        width = max(8, min(chunk_cols, max((int(t.shape[1]) for t in shards), default=8)))
        width = (width + 7) // 8 * 8
        stage = [torch.empty((EMBEDDINGS_DIM, width), dtype=torch.float16).pin_memory() for _ in range(2)]
        dstage = [torch.empty((EMBEDDINGS_DIM, width), dtype=torch.float16, device=device) for _ in range(2)]
        done = [torch.cuda.Event(), torch.cuda.Event()]
        used = [False, False]
        stream = torch.cuda.current_stream(device)
        row0, it = 0, 0
        for t in shards:
            n = int(t.shape[1])
            for a in range(0, n, width):
                w = min(width, n - a)
                k = it & 1
                if used[k]:
                    done[k].synchronize()                      # the pinned buffer is free again
                src = t[:, a:a + w]
                stage[k][:, :w].copy_(src if src.dtype == torch.float16 else src.to(torch.float16))   # the disk read
                dstage[k][:, :w].copy_(stage[k][:, :w], non_blocking=True)
                done[k].record(stream)
                used[k] = True
                view = dstage[k][:, :w]
                if w % 2 == 0:
                    dst = bank.data_ptr() + (row0 + a) * bank.stride(0) * 2
                    check(lib().atlas_b200_transpose(view.data_ptr(), view.stride(0), dst, bank.stride(0), EMBEDDINGS_DIM, w,
                                                     EMBEDDINGS_DIM, ctypes.c_void_p(stream.cuda_stream)))
                else:   # odd tail width: the transpose kernel moves 2-column words
                    bank[row0 + a:row0 + a + w].copy_(view.t())
                it += 1
            row0 += n
        torch.cuda.current_stream(device).synchronize()
        return bank

    # ------------------------------------------------------------------ search
    def _compute_scores_and_indices(self, allqueries: torch.Tensor, topk: int):
        """src/index.py:113-120 semantics (fp16 scores of `allqueries.half() @ embeddings`, row top-k),
        fused.  Returns (scores [nq, k] fp16 desc, indices [nq, k] int64 LOCAL row numbers)."""
        return ops.search_shard(self._bank, allqueries, topk, 0, 1, self._workspace)

    def _merge(self, blob_all, ids_off, world, nq_total, topk, q_begin, nq_out):
        return ops.topk_merge_blob(blob_all, ids_off, world, nq_total, topk, q_begin, nq_out, torch.float16)

    def _owner_local(self, gid, world):
        if self._offsets is None:
            return gid % world, gid // world
        r = int(np.searchsorted(self._offsets, gid, side="right")) - 1
        return r, gid - int(self._offsets[r])

    def _local_search(self, allqueries, topk, exhaustive=False):
        """Local shard scan returning GLOBAL ids + the device status word (no host synchronisation)."""
        return ops.mips_topk(self._bank, allqueries, topk, self._id_base, self._id_stride, self._workspace,
                             exhaustive=exhaustive)

    @torch.no_grad()
    def search_device(self, queries, topk, return_status=False, exhaustive=False):
        """Device half of `search_knn`: (scores [nq_local, k] fp16, global ids [nq_local, k] int64) as CUDA tensors.

        Collective when world_size > 1: ONE all-gather of the query rows + ONE all-gather of the packed per-shard
        (scores | ids | status) blobs (the reference: 3 + 4*W collectives with pickled passages, src/index.py:127-143).
        With `max_queries_per_rank` set there is no host synchronisation at all; otherwise one scalar size exchange.

        return_status=False: the overflow flag of the fast path is read here (one host sync) and the exhaustive path is
        re-run if needed, so the result is final.  return_status=True: nothing synchronises; the third return value is a
        device int64 scalar (max over shards) that the caller checks at its own synchronisation point
        (`search_knn` folds it into the D2H copy it needs anyway) and, if non-zero, calls again with exhaustive=True."""
        if self._bank is None:
            raise AtlasB200Error("search_knn before init_embeddings/load_index")
        if topk > MAX_TOPK:
            raise AtlasB200Error(f"topk={topk} > {MAX_TOPK}")
        if topk > self._bank.shape[0]:
            # torch.topk in the reference raises the same way (src/index.py:118)
            raise RuntimeError(f"selected index k out of range: topk={topk} > local bank size {self._bank.shape[0]}")
        world = dist_utils.get_world_size()
        rank = dist_utils.get_rank()
        queries = queries.reshape(-1, EMBEDDINGS_DIM)
        nq_local = queries.shape[0]
        if world == 1:
            s, i, status = self._local_search(queries, topk, exhaustive)
            status = status.to(torch.int64).reshape(())
        else:
            q16 = queries.to(self._bank.device).to(torch.float16)                  # `.half()`, src/index.py:117
            cap = self.max_queries_per_rank
            if cap:
                if nq_local > cap:
                    raise AtlasB200Error(f"search_knn: {nq_local} queries on this rank but max_queries_per_rank={cap} "
                                         "(set it to the per-GPU batch size, identical on every rank, or to None)")
                if nq_local < cap:
                    q16 = torch.cat([q16, q16.new_zeros((cap - nq_local, EMBEDDINGS_DIM))], dim=0)
                allq = dist_utils.all_gather_fixed(q16).reshape(world * cap, EMBEDDINGS_DIM)     # all_gather #1, no sync
                nq_total, q_begin = world * cap, rank * cap
            else:
                sizes = dist_utils.get_varsize(queries)                            # tiny all_gather (+ host sync)
                allq = dist_utils.varsize_all_gather(q16, sizes)                   # all_gather #1
                nq_total, q_begin = int(sum(sizes)), int(sum(sizes[:rank]))
            s_loc, i_loc, st_loc = self._local_search(allq, topk, exhaustive)
            blob, ids_off, st_off = ops.pack_results(s_loc, i_loc, st_loc.to(torch.int64))
            blob_all = dist_utils.all_gather_fixed(blob)                           # all_gather #2
            s, i = self._merge(blob_all, ids_off, world, nq_total, topk, q_begin, nq_local)
            status = blob_all[:, st_off:].contiguous().view(torch.int64).max()
        if return_status:
            return s, i, status
        if not exhaustive and int(status.item()) != 0:      # identical on every rank (max over shards): collective retry
            return self.search_device(queries, topk, exhaustive=True)
        return s, i

    @torch.no_grad()
    def search_knn(self, queries, topk):
        """Exhaustive k-nearest-neighbour search by inner product (src/index.py:122-157).

        Collective: every rank must call it, also with 0 queries (src/atlas.py:103-106).
        Returns (docs: List[nq][k] passage dicts, scores: List[nq][k] floats, descending)."""
        world = dist_utils.get_world_size()
        nq_local = queries.reshape(-1, EMBEDDINGS_DIM).shape[0]
        scores, ids, status = self.search_device(queries, topk, return_status=True)
        # the overflow flag (max over shards, identical on every rank) is read at the host sync the results need anyway
        scores_host = scores.float().cpu()
        ids_host = ids.cpu()
        if int(status.item()) != 0:                         # massive ties (e.g. an all-zero bank): exact chunked path
            scores, ids, _ = self.search_device(queries, topk, return_status=True, exhaustive=True)
            scores_host, ids_host = scores.float().cpu(), ids.cpu()
        ids_host = ids_host.tolist()
        flat = [self._owner_local(g, world) for row in ids_host for g in row]
        docs_flat = self._get_store().lookup(flat)
        docs = [docs_flat[r * topk:(r + 1) * topk] for r in range(nq_local)]
        return docs, scores_host.tolist()

    def is_index_trained(self) -> bool:  # src/index.py:159-160
        return True

    def train_index(self):
        return None


class DistributedFAISSIndex(DistributedIndex):
    """src/index.py:163-381 wraps faiss-gpu IVF/PQ indices.  Out of scope here (north_star: "No FAISS");
    the exact flat index above is the supported mode and is what the README recommends."""

    def __init__(self, index_type: str = "flat", code_size: Optional[int] = None):
        raise AtlasB200Error(
            "DistributedFAISSIndex is not provided by atlas_b200: use --index_mode flat "
            "(exact search, atlas_b200.index.DistributedIndex)"
        )
