"""Passage-dict lookup by (owner rank, local index) for the ids-only search exchange.

The reference ships pickled passage dicts between ranks on every search (src/index.py:33-40,
135-143).  Here only (score, id) pairs travel; the text of the k winners is resolved locally:

  * world_size == 1      : the index's own `doc_map`.
  * one node, W ranks    : `SharedPassageStore` - every rank serialises its shard ONCE into
                           /dev/shm (pickle records + an int64 offset table); all ranks mmap all
                           W shard files, so a lookup is a slice + `pickle.loads`, no communication.
  * ranks on >1 node     : `ExchangePassageStore` - two `all_gather_object` calls per search that
                           carry only the requested winners (slow path, kept for correctness).
"""
import mmap
import os
import pickle
import shutil

import numpy as np
import torch.distributed as dist


class LocalPassageStore:
    def __init__(self, doc_map):
        self.doc_map = doc_map

    def lookup(self, owners_locals):
        return [self.doc_map[l] for _, l in owners_locals]

    def close(self):
        pass


class SharedPassageStore:
    """Node-shared, read-only passage store in /dev/shm.  Construction is collective."""

    def __init__(self, doc_map, rank, world, root="/dev/shm"):
        self.rank, self.world = rank, world
        tag = [None]
        if rank == 0:
            tag[0] = f"atlas_b200_{os.getpid()}_{id(self) & 0xFFFFFF:x}"
        dist.broadcast_object_list(tag, src=0)
        self.dir = os.path.join(root, tag[0])
        os.makedirs(self.dir, exist_ok=True)
        n = len(doc_map)
        offsets = np.zeros(n + 1, dtype=np.int64)
        with open(self._data_path(rank), "wb") as f:
            pos = 0
            for i in range(n):
                blob = pickle.dumps(doc_map[i], protocol=pickle.HIGHEST_PROTOCOL)
                f.write(blob)
                pos += len(blob)
                offsets[i + 1] = pos
        np.save(self._off_path(rank), offsets)
        dist.barrier()
        self._maps = {}
        self._offs = {}
        self._files = {}
        if rank == 0:       # the directory must not outlive the job when a process exits without close()
            import atexit

            atexit.register(shutil.rmtree, self.dir, True)

    def _data_path(self, r):
        return os.path.join(self.dir, f"passages.{r}.bin")

    def _off_path(self, r):
        return os.path.join(self.dir, f"offsets.{r}.npy")

    def _open(self, r):
        if r not in self._maps:
            self._offs[r] = np.load(self._off_path(r), mmap_mode="r")
            f = open(self._data_path(r), "rb")
            self._files[r] = f
            size = os.fstat(f.fileno()).st_size
            self._maps[r] = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) if size > 0 else b""
        return self._maps[r], self._offs[r]

    def lookup(self, owners_locals):
        out = []
        for r, l in owners_locals:
            m, off = self._open(r)
            out.append(pickle.loads(m[int(off[l]): int(off[l + 1])]))
        return out

    def close(self):
        for m in self._maps.values():
            if not isinstance(m, bytes):
                m.close()
        for f in self._files.values():
            f.close()
        self._maps, self._files = {}, {}
        if dist.is_initialized():
            dist.barrier()
        if self.rank == 0:
            shutil.rmtree(self.dir, ignore_errors=True)


class ExchangePassageStore:
    """Multi-node fallback: resolve remote winners with two object all-gathers per search."""

    def __init__(self, doc_map, rank, world):
        self.doc_map, self.rank, self.world = doc_map, rank, world

    def lookup(self, owners_locals):
        wanted = [None] * self.world
        dist.all_gather_object(wanted, [(r, l) for r, l in owners_locals])
        mine = [[self.doc_map[l] for r, l in req if r == self.rank] for req in wanted]
        answers = [None] * self.world
        dist.all_gather_object(answers, mine)
        cursor = [0] * self.world
        out = []
        for r, _ in owners_locals:
            out.append(answers[r][self.rank][cursor[r]])
            cursor[r] += 1
        return out

    def close(self):
        pass


def single_node(world):
    """True when every rank of the process group runs on this host.  Decided by exchanging host names (collective):
    LOCAL_WORLD_SIZE is only set by torchrun - the reference's own SLURM launcher (src/slurm.py) never sets it, and a
    multi-node job must not pick the /dev/shm store."""
    local = os.environ.get("LOCAL_WORLD_SIZE")
    if local is not None and int(local) != world:
        return False
    if not (dist.is_available() and dist.is_initialized()):
        return True
    import socket

    names = [None] * world
    dist.all_gather_object(names, socket.gethostname())
    return len(set(names)) == 1


def make_store(doc_map, rank, world):
    if world == 1:
        return LocalPassageStore(doc_map)
    if single_node(world) and os.path.isdir("/dev/shm") and os.environ.get("ATLAS_B200_PASSAGE_STORE", "shm") == "shm":
        return SharedPassageStore(doc_map, rank, world)
    return ExchangePassageStore(doc_map, rank, world)
