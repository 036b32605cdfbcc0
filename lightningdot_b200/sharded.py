"""Row-sharded exact search over the GPUs of one box (BASELINE config 4).

The reference has a single-process CPU index (dvl/indexer/faiss_indexers.py:63-87).  Here the candidate matrix is
split by ROWS over the ranks of a torch.distributed group (one process per GPU): rank r owns rows
[bounds[r], bounds[r+1]) as an ordinary FlatIPIndex with row_offset = bounds[r], so the ids it returns are already
global.  A search is

    every rank:  exact top-k of ALL queries against ITS rows      (fused score + top-k kernel, exact rescoring)
    exchange:    all-gather of the per-shard (score fp32, id int64) [nq, k] lists  (NCCL over NVLink; 12 B * nq * k per rank)
    every rank:  k best of the W * k candidates per query, ranked (score desc, id asc)   (ldot_topk_merge)

The union of the shard top-k lists contains the global top-k, and every shard list is exact, so the merged result
equals the single-index result bit for bit.  `index_id_to_db_id` is replicated on every rank (host memory).
Queries may be given replicated, or sharded (each rank encodes nq / W queries) and gathered here.
"""
from typing import List

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .indexer import DenseFlatIndexer, FlatIPIndex


def shard_bounds(n, world):
    """Row ranges of an n-row index over `world` ranks: sizes differ by at most one, rank order = row order."""
    base, rem = divmod(int(n), int(world))
    b = [0]
    for r in range(world):
        b.append(b[-1] + base + (1 if r < rem else 0))
    return b


class ShardedFlatIndexer(DenseFlatIndexer):
    """DenseFlatIndexer whose rows are spread over the ranks of `group` (default: the world group).

    index_data / index_matrix take the FULL id list and vector matrix on every rank (as the reference's callers
    build them) and keep only this rank's rows on the device; index_shard takes an already-partitioned shard.
    search_knn returns the full, merged result on every rank."""

    def __init__(self, vector_sz: int, buffer_size: int = 50000, group=None, **index_kw):
        super().__init__(vector_sz, buffer_size=buffer_size, **index_kw)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.vector_sz = vector_sz
        self._index_kw = index_kw
        self.n_global = 0
        self.bounds = [0] * (self.world + 1)

    # -- building -----------------------------------------------------------------------------------------------
    def index_matrix(self, db_ids: List[object], vectors):
        if self.n_global:
            raise NotImplementedError("a sharded index is built by one index_data / index_matrix / index_shard call")
        n = len(db_ids)
        assert n == vectors.shape[0]
        self.bounds = shard_bounds(n, self.world)
        lo, hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        self._set_shard(list(db_ids), vectors[lo:hi], lo, n)

    def index_data(self, data):
        ids = [t[0] for t in data]
        if len(data) and isinstance(data[0][1], torch.Tensor):
            mat = torch.stack([t[1].reshape(-1) for t in data], dim=0)
        else:
            mat = np.stack([np.reshape(t[1], (-1,)) for t in data], axis=0) if len(data) else \
                np.zeros((0, self.vector_sz), np.float32)
        self.index_matrix(ids, mat)

    def index_shard(self, all_db_ids: List[object], shard_vectors, bounds):
        """This rank's rows only: shard_vectors = rows [bounds[rank], bounds[rank + 1]) of the global matrix."""
        assert len(bounds) == self.world + 1 and bounds[-1] == len(all_db_ids)
        assert shard_vectors.shape[0] == bounds[self.rank + 1] - bounds[self.rank]
        self.bounds = list(bounds)
        self._set_shard(list(all_db_ids), shard_vectors, bounds[self.rank], bounds[-1])

    def _set_shard(self, all_ids, shard_vectors, row_offset, n_global):
        self.index_id_to_db_id = all_ids
        self.n_global = n_global
        self.index = self._make_local_index(row_offset)
        if shard_vectors.shape[0]:
            self.index.add(shard_vectors)

    def _make_local_index(self, row_offset):
        return FlatIPIndex(self.vector_sz, row_offset=row_offset, **self._index_kw)

    # -- searching ----------------------------------------------------------------------------------------------
    def gather_queries(self, local_queries, counts=None):
        """Each rank holds a contiguous slice of the query matrix (equal sizes except possibly the last ranks
        shorter): all-gather -> the full [nq, d] matrix in rank order.  `counts`: rows per rank when the caller knows
        them (shard_bounds of a known total): skips the count exchange and its host synchronisation, so the search
        kernels queue up behind the tower without a bubble."""
        if self.world == 1:
            return local_queries
        if counts is None:
            counts = self._all_gather_counts(local_queries.shape[0], local_queries.device)
        elif len(counts) != self.world or counts[self.rank] != local_queries.shape[0]:
            raise ValueError(f"counts {counts} do not match the local block of {local_queries.shape[0]} rows on rank {self.rank}")
        m = max(counts)
        if min(counts) == m:   # equal blocks: one collective straight into the result
            out = local_queries.new_empty((m * self.world, local_queries.shape[1]))
            dist.all_gather_into_tensor(out, local_queries.contiguous(), group=self.group)
            return out
        padded = local_queries.new_zeros((m, local_queries.shape[1]))
        padded[:local_queries.shape[0]] = local_queries
        parts = [torch.empty_like(padded) for _ in range(self.world)]
        dist.all_gather(parts, padded, group=self.group)
        return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)

    def _all_gather_counts(self, n, device):
        t = torch.tensor([n], dtype=torch.int64, device=device)
        parts = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(parts, t, group=self.group)
        return [int(p.item()) for p in parts]

    def search_device(self, queries, k):
        """queries: the full [nq, d] matrix (same on every rank), on this rank's device -> merged (scores [nq, k]
        fp32, global row ids [nq, k] int64), identical on every rank."""
        ls, li = self._local_search(queries, k)
        if self.world == 1:
            return ls, li
        nq = queries.shape[0]
        gs = torch.empty((self.world, nq, k), dtype=torch.float32, device=ls.device)
        gi = torch.empty((self.world, nq, k), dtype=torch.int64, device=ls.device)
        # one collective each, straight into the [W, nq, k] buffers (gloo in the CPU tests accepts the same calls)
        dist.all_gather_into_tensor(gs.view(self.world * nq, k), ls.contiguous(), group=self.group)
        dist.all_gather_into_tensor(gi.view(self.world * nq, k), li.contiguous(), group=self.group)
        return self._merge(gs, gi, k)

    def _local_search(self, queries, k):
        if self.index.ntotal == 0:  # more ranks than rows: an empty shard contributes nothing
            nq = queries.shape[0]
            return (torch.full((nq, k), -3.4028235e38, dtype=torch.float32, device=queries.device),
                    torch.full((nq, k), -1, dtype=torch.int64, device=queries.device))
        return self.index.search_device(queries, k)

    def _merge(self, gs, gi, k):
        lib = _lib.load()
        world, nq, _ = gs.shape
        out_s = torch.empty((nq, k), dtype=torch.float32, device=gs.device)
        out_i = torch.empty((nq, k), dtype=torch.int64, device=gs.device)
        _lib.check(lib.ldot_topk_merge(_lib.ptr(gs), _lib.ptr(gi), world, nq, k, _lib.ptr(out_s), _lib.ptr(out_i),
                                       _lib.stream_ptr()))
        return out_s, out_i

    def _queries_on_device(self, query_vectors):
        if isinstance(query_vectors, np.ndarray):
            query_vectors = torch.from_numpy(np.ascontiguousarray(query_vectors, dtype=np.float32))
        return query_vectors.detach().to(device=self._device(), dtype=torch.float32).contiguous()

    def _device(self):
        return self.index._device()

    def search(self, query_vectors, k: int):
        """faiss-level call (IndexFlatIP.search, faiss_indexers.py:83) over the sharded index:
        -> (scores float32 [nq, k], global row labels int64 [nq, k]) numpy arrays."""
        scores, idx = self.search_device(self._queries_on_device(query_vectors), k)
        return self._to_host(scores, idx)

    def _to_host(self, scores, idx):
        return self.index.to_host(scores, idx)

    def search_knn(self, query_vectors, top_docs: int):
        scores, idx = self.search_device(self._queries_on_device(query_vectors), top_docs)
        return self._format_result(*self._to_host(scores, idx))
