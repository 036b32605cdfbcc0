"""Host-side mirror of the reference's index classes (dvl/indexer/faiss_indexers.py) over libldot_sm100a.

  DenseIndexer / DenseFlatIndexer   same constructor, attributes (index, index_id_to_db_id, buffer_size) and methods
                                    (index_data, search_knn, serialize, deserialize_from) as faiss_indexers.py:22-87
  FlatIPIndex                       what `DenseFlatIndexer.index` holds instead of faiss.IndexFlatIP
                                    (faiss_indexers.py:67): d, ntotal, add(), search(), reset() - device resident

The index lives in HBM as an fp32 master copy [n, d] (exact rescoring reads it) plus a centred 16-bit copy for the
tensor-core pass.  search() is exact: ids are ranked by (correctly rounded fp32 inner product desc, row id asc);
queries whose exactness certificate fails are transparently re-run through the exhaustive fp64-accumulated scan.
There is no CPU path: without the CUDA extension and a B200 every call raises.
"""
import gc
import os
import logging
import pickle
import struct
from typing import List, Tuple

import numpy as np
import torch

from . import _lib

logger = logging.getLogger()

MAX_QUERY_BATCH = 32768  # queries per C-ABI call (bounds the candidate-list workspace)
WIDE_COARSE_K = 960      # candidate-list width of the second tier (queries whose first certificate failed)


def _as_device_f32(a, device):
    if isinstance(a, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        return t.to(device, non_blocking=False)
    if isinstance(a, torch.Tensor):
        return a.detach().to(device=device, dtype=torch.float32).contiguous()
    raise TypeError(f"expected numpy array or torch tensor, got {type(a)}")


class _Workspace:
    """Grow-only device scratch buffer handed to the C ABI (the library itself never allocates)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
        off = (-self.buf.data_ptr()) % 256
        return self.buf[off:off + nbytes]


class FlatIPIndex:
    """Exact inner-product index, device resident.  Mirrors the slice of faiss.IndexFlatIP the reference uses
    (add / search / ntotal / d; faiss_indexers.py:67,77,83)."""

    def __init__(self, d, device=None, coarse_dtype="fp16", center=True, coarse_k=0, row_offset=0):
        if d % 8 != 0 or not 8 <= d <= 4096:
            raise ValueError(f"vector size must be a multiple of 8 in [8, 4096], got {d}")
        self.d = int(d)
        self.device = torch.device(device) if device is not None else None
        self.coarse_dtype = {"fp16": _lib.COARSE_FP16, "bf16": _lib.COARSE_BF16}[coarse_dtype]
        self.center = bool(center)
        self.coarse_k = int(coarse_k)
        self.row_offset = int(row_offset)   # global id of local row 0 (row-sharded index)
        self.is_trained = True
        self._chunks = []
        self._n = 0
        self._x = None        # [n, d] fp32 master
        self._x16 = None      # [n, d] centred 16-bit copy
        self._mu = None       # [d]
        self._xstats = None   # [2]
        self._ws = _Workspace()
        self._pinned = None
        self._pin_counts = None
        self.last_flagged = 0      # queries of the last search() whose first-tier certificate failed
        self.last_exhaustive = 0   # ... of which the second tier could not certify either (exhaustive scan)

    # -- faiss-like surface -------------------------------------------------------------------------------------
    @property
    def ntotal(self):
        return self._n

    def reset(self):
        self._chunks, self._n = [], 0
        self._x = self._x16 = self._mu = self._xstats = None

    def add(self, x):
        dev = self._device()
        t = _as_device_f32(x, dev)
        if t.dim() != 2 or t.shape[1] != self.d:
            raise ValueError(f"add() expects [n, {self.d}], got {tuple(t.shape)}")
        if t.shape[0] == 0:
            return
        if self._x is not None:
            self._chunks = [self._x]
        self._chunks.append(t)
        self._n += t.shape[0]
        self._x = self._x16 = None

    def search(self, q, k):
        """-> (scores float32 [nq, k], labels int64 [nq, k]) as numpy arrays (the faiss return convention)."""
        dev = self._device()
        qd = _as_device_f32(q, dev)
        scores, idx = self.search_device(qd, k)
        return self.to_host(scores, idx)

    def to_host(self, scores, idx):
        """Device results -> numpy through cached pinned staging buffers (one synchronisation)."""
        nq, k = scores.shape
        pin = self._pinned
        if pin is None or pin[0].shape[0] < nq or pin[0].shape[1] != k:
            pin = (torch.empty((max(nq, 1), k), dtype=torch.float32, pin_memory=True),
                   torch.empty((max(nq, 1), k), dtype=torch.int64, pin_memory=True))
            self._pinned = pin
        pin[0][:nq].copy_(scores, non_blocking=True)
        pin[1][:nq].copy_(idx, non_blocking=True)
        torch.cuda.current_stream(scores.device).synchronize()
        return pin[0][:nq].numpy().copy(), pin[1][:nq].numpy().copy()

    # -- device-level API (used by the eval loop and the benchmark to avoid host round trips) -------------------
    def search_device(self, qd, k, resolve_flags=True, return_flags=False):
        """qd: cuda fp32 [nq, d].  -> (scores [nq, k] fp32, labels [nq, k] int64) cuda tensors.
        With resolve_flags (default) flagged queries are re-run exhaustively (one 4-byte D2H sync per call).
        resolve_flags=False, return_flags=True: no host synchronisation at all (CUDA-graph capturable); additionally
        returns (flags int32 [nq] on the device, per-call flagged counts in PINNED host memory) for the caller to check
        after it has synchronised (online.py)."""
        lib = _lib.load()
        self._finalize()
        if qd.dim() != 2 or qd.shape[1] != self.d:
            raise ValueError(f"search() expects [nq, {self.d}], got {tuple(qd.shape)}")
        if not 1 <= k <= 1024:
            raise ValueError(f"top_docs must be in [1, 1024], got {k}")
        nq = qd.shape[0]
        dev = qd.device
        scores = torch.empty((nq, k), dtype=torch.float32, device=dev)
        idx = torch.empty((nq, k), dtype=torch.int64, device=dev)
        if nq == 0:
            return scores, idx
        if self._n == 0:
            scores.fill_(-3.4028235e38)
            idx.fill_(-1)
            return scores, idx
        flags = torch.empty((nq,), dtype=torch.int32, device=dev)
        n_flag = self._search_pass(qd, k, self.coarse_k, scores, idx, flags)
        self.last_flagged = 0
        self.last_exhaustive = 0
        if resolve_flags:
            torch.cuda.current_stream(dev).synchronize()   # the flagged-query counts land in pinned host memory
            n_flag = int(n_flag.sum())
            self.last_flagged = n_flag
            if n_flag:
                # tier 2: the same tensor-core pipeline with a much wider candidate list (the certificate margin
                # grows with k' - k); tier 3: the exhaustive fp64-accumulated scan for whatever is still uncertified
                rows = torch.nonzero(flags, as_tuple=False).flatten()
                q2 = qd.index_select(0, rows).contiguous()
                s2 = torch.empty((n_flag, k), dtype=torch.float32, device=dev)
                i2 = torch.empty((n_flag, k), dtype=torch.int64, device=dev)
                wide = min(WIDE_COARSE_K, 1280)
                if wide >= 2 * k and wide > self._auto_coarse_k(k):
                    f2 = torch.empty((n_flag,), dtype=torch.int32, device=dev)
                    c2 = self._search_pass(q2, k, wide, s2, i2, f2)
                    torch.cuda.current_stream(dev).synchronize()
                    n2 = int(c2.sum())
                else:
                    f2, n2 = torch.ones((n_flag,), dtype=torch.int32, device=dev), n_flag
                if n2:
                    sub = torch.nonzero(f2, as_tuple=False).flatten()
                    es, ei = self.exact_search_device(q2.index_select(0, sub).contiguous(), k)
                    s2.index_copy_(0, sub, es)
                    i2.index_copy_(0, sub, ei)
                    self.last_exhaustive = n2
                scores.index_copy_(0, rows, s2)
                idx.index_copy_(0, rows, i2)
        if return_flags:
            return scores, idx, flags, n_flag
        return scores, idx

    def _auto_coarse_k(self, k):
        if self.coarse_k:
            return self.coarse_k
        kp = max(min(k + max(k // 2, 32), 1280), 64)
        return (kp + 31) // 32 * 32

    def _search_pass(self, qd, k, coarse_k, scores, idx, flags):
        """One ldot_flatip_search call per MAX_QUERY_BATCH queries -> per-call flagged counts (PINNED HOST int32
        tensor, valid once the stream has been synchronised)."""
        lib = _lib.load()
        nq, dev = qd.shape[0], qd.device
        nb_calls = (nq + MAX_QUERY_BATCH - 1) // MAX_QUERY_BATCH
        if self._pin_counts is None or self._pin_counts.numel() < nb_calls:
            self._pin_counts = torch.zeros((max(nb_calls, 16),), dtype=torch.int32, pin_memory=True)
        counts = self._pin_counts[:nb_calls]
        stream = _lib.stream_ptr()
        for bi, b in enumerate(range(0, nq, MAX_QUERY_BATCH)):
            e = min(nq, b + MAX_QUERY_BATCH)
            nb = e - b
            need = lib.ldot_flatip_search_workspace_bytes(nb, self._n, self.d, k, coarse_k)
            if need == 0:
                raise _lib.LdotError(f"invalid search shape: {lib.ldot_last_error().decode()}")
            ws = self._ws.get(need, dev)
            _lib.check(lib.ldot_flatip_search(
                _lib.ptr(qd[b:e]), nb, _lib.ptr(self._x), _lib.ptr(self._x16), _lib.ptr(self._mu),
                _lib.ptr(self._xstats), self._n, self.d, k, coarse_k, self.coarse_dtype, self.row_offset,
                _lib.ptr(scores[b:e]), _lib.ptr(idx[b:e]), _lib.ptr(flags[b:e]), _lib.ptr(counts[bi:bi + 1]),
                _lib.ptr(ws), need, stream))
        return counts

    # -- two-phase form for a row-sharded index (sharded.py): phase 1 up to the candidate lists + a per-query bound the
    # shards min-reduce, phase 2 rescoring pruned by that bound (include/ldot.h: ldot_flatip_search_phase)
    def search_phase1(self, qd, k, bound_m):
        """-> state for search_phase2; state['bound'] is the fp32 [nq] tensor the shards combine with a MIN all-reduce.
        One C-ABI call: nq <= MAX_QUERY_BATCH."""
        lib = _lib.load()
        self._finalize()
        nq, dev = qd.shape[0], qd.device
        if not (1 <= nq <= MAX_QUERY_BATCH):
            raise ValueError(f"search_phase1 takes 1 .. {MAX_QUERY_BATCH} queries, got {nq}")
        st = dict(q=qd, k=k, scores=torch.empty((nq, k), dtype=torch.float32, device=dev),
                  idx=torch.empty((nq, k), dtype=torch.int64, device=dev),
                  flags=torch.empty((nq,), dtype=torch.int32, device=dev),
                  count=torch.zeros((1,), dtype=torch.int32, device=dev),
                  bound=torch.empty((nq,), dtype=torch.float32, device=dev))
        need = lib.ldot_flatip_search_workspace_bytes(nq, self._n, self.d, k, self.coarse_k)
        if need == 0:
            raise _lib.LdotError(f"invalid search shape: {lib.ldot_last_error().decode()}")
        st["ws"], st["need"] = self._ws.get(need, dev), need
        self._phase_call(st, 1, bound_m, None)
        return st

    def search_phase2(self, st, tau):
        """tau: fp32 [nq] lower bounds of the global k-th best score (or None: no pruning) -> (scores, idx, flags,
        flagged-count tensor [1]) on the device; no host synchronisation."""
        self._phase_call(st, 2, 0, tau)
        return st["scores"], st["idx"], st["flags"], st["count"]

    def _phase_call(self, st, phase, bound_m, tau):
        lib = _lib.load()
        qd, k = st["q"], st["k"]
        _lib.check(lib.ldot_flatip_search_phase(
            _lib.ptr(qd), qd.shape[0], _lib.ptr(self._x), _lib.ptr(self._x16), _lib.ptr(self._mu), _lib.ptr(self._xstats),
            self._n, self.d, k, self.coarse_k, self.coarse_dtype, self.row_offset, _lib.ptr(st["scores"]),
            _lib.ptr(st["idx"]), _lib.ptr(st["flags"]), _lib.ptr(st["count"]), _lib.ptr(st["ws"]), st["need"], phase, bound_m,
            _lib.ptr(st["bound"]), _lib.ptr(tau), _lib.stream_ptr()))

    def exact_search_device(self, qd, k):
        """Exhaustive fp64-accumulated scan (no tensor cores) - the fallback path, also usable on its own."""
        lib = _lib.load()
        self._finalize()
        nq = qd.shape[0]
        scores = torch.empty((nq, k), dtype=torch.float32, device=qd.device)
        idx = torch.empty((nq, k), dtype=torch.int64, device=qd.device)
        if nq == 0:
            return scores, idx
        need = lib.ldot_flatip_exact_workspace_bytes(self._n)
        ws = self._ws.get(need, qd.device)
        _lib.check(lib.ldot_flatip_exact(_lib.ptr(qd), nq, _lib.ptr(self._x), self._n, self.d, k, self.row_offset,
                                         _lib.ptr(scores), _lib.ptr(idx), _lib.ptr(ws), need, _lib.stream_ptr()))
        return scores, idx

    # -- internals ----------------------------------------------------------------------------------------------
    def _device(self):
        if self.device is None:
            if not torch.cuda.is_available():
                raise _lib.LdotError("FlatIPIndex needs a CUDA device (B200); there is no CPU fallback")
            self.device = torch.device("cuda", torch.cuda.current_device())
        return self.device

    def _finalize(self):
        """Concatenate pending adds and (re)build the centred 16-bit copy + certificate statistics."""
        if self._x16 is not None or self._n == 0:
            return
        lib = _lib.load()
        _lib.check(lib.ldot_device_check())
        dev = self._device()
        if self._x is None:
            self._x = self._chunks[0] if len(self._chunks) == 1 else torch.cat(self._chunks, dim=0)
            self._chunks = []
        x16_dtype = torch.float16 if self.coarse_dtype == _lib.COARSE_FP16 else torch.bfloat16
        self._x16 = torch.empty((self._n, self.d), dtype=x16_dtype, device=dev)
        self._mu = torch.empty((self.d,), dtype=torch.float32, device=dev)
        self._xstats = torch.empty((2,), dtype=torch.float32, device=dev)
        need = lib.ldot_index_prepare_workspace_bytes(self._n, self.d)
        ws = self._ws.get(need, dev)
        _lib.check(lib.ldot_index_prepare(_lib.ptr(self._x), self._n, self.d, self.coarse_dtype, int(self.center),
                                          _lib.ptr(self._x16), _lib.ptr(self._mu), _lib.ptr(self._xstats),
                                          _lib.ptr(ws), need, _lib.stream_ptr()))

    def vectors(self):
        """fp32 master copy as a cuda tensor [ntotal, d]."""
        self._finalize()
        return self._x


# faiss on-disk layout of an IndexFlatIP (faiss/impl/index_write.cpp, faiss 1.6.x) [external, unverified here: faiss
# is not installable offline]: fourcc "IxFI", int32 d, int64 ntotal, int64 dummy, int64 dummy, uint8 is_trained,
# int32 metric_type (0 = inner product), uint64 len(xb), float32 xb[ntotal * d].
_FAISS_FLAT_IP_FOURCC = b"IxFI"
_FAISS_HEADER = struct.Struct("<iqqqBi")


def write_flat_ip_index(index: "FlatIPIndex", path: str):
    x = index.vectors().cpu().numpy() if index.ntotal else np.zeros((0, index.d), np.float32)
    with open(path, "wb") as f:
        f.write(_FAISS_FLAT_IP_FOURCC)
        f.write(_FAISS_HEADER.pack(index.d, index.ntotal, 1 << 20, 1 << 20, 1, 0))
        f.write(struct.pack("<Q", x.size))
        f.write(np.ascontiguousarray(x, np.float32).tobytes())


def read_flat_ip_index(path: str, **kw) -> "FlatIPIndex":
    with open(path, "rb") as f:
        if f.read(4) != _FAISS_FLAT_IP_FOURCC:
            raise ValueError(f"{path}: not a faiss IndexFlatIP file")
        d, ntotal, _, _, _, metric = _FAISS_HEADER.unpack(f.read(_FAISS_HEADER.size))
        if metric != 0:
            raise ValueError(f"{path}: metric_type {metric} is not inner product")
        (size,) = struct.unpack("<Q", f.read(8))
        if size != ntotal * d:
            raise ValueError(f"{path}: vector payload {size} != ntotal * d")
        x = np.frombuffer(f.read(size * 4), dtype=np.float32).reshape(ntotal, d)
    idx = FlatIPIndex(d, **kw)
    if ntotal:
        idx.add(x)
    return idx


_PYHOST = [False, None]   # [looked for, ldot_py_gather_lists or None]


def _pyhost_gather():
    """The CPython helper that builds search_knn's id lists (lightningdot_b200/_ldot_pyhost.so, built by build.py), or
    None when it is not there - the numpy formulation below is then used; both are host-side Python-object plumbing."""
    if not _PYHOST[0]:
        _PYHOST[0] = True
        import ctypes
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ldot_pyhost.so")
        if os.path.exists(path):
            try:
                fn = ctypes.PyDLL(path).ldot_py_gather_lists
                fn.restype = ctypes.py_object
                fn.argtypes = [ctypes.py_object, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong]
                _PYHOST[1] = fn
            except (OSError, AttributeError):
                _PYHOST[1] = None
    return _PYHOST[1]


class DenseIndexer(object):
    """dvl/indexer/faiss_indexers.py:22-60."""

    def __init__(self, buffer_size: int = 50000):
        self.buffer_size = buffer_size
        self.index_id_to_db_id = []
        self.index = None

    def index_data(self, data: List[Tuple[object, np.array]]):
        raise NotImplementedError

    def search_knn(self, query_vectors: np.array, top_docs: int) -> List[Tuple[List[object], List[float]]]:
        raise NotImplementedError

    def serialize(self, file: str):
        logger.info('Serializing index to %s', file)
        index_file = file + '.index.dpr'
        meta_file = file + '.index_meta.dpr'
        write_flat_ip_index(self.index, index_file)
        with open(meta_file, mode='wb') as f:
            pickle.dump(self.index_id_to_db_id, f)

    def deserialize_from(self, file: str):
        logger.info('Loading index from %s', file)
        index_file = file + '.index.dpr'
        meta_file = file + '.index_meta.dpr'
        self.index = read_flat_ip_index(index_file)
        logger.info('Loaded index of type %s and size %d', type(self.index), self.index.ntotal)
        with open(meta_file, "rb") as reader:
            self.index_id_to_db_id = pickle.load(reader)
        assert len(self.index_id_to_db_id) == self.index.ntotal, \
            'Deserialized index_id_to_db_id should match faiss index size'

    def _update_id_mapping(self, db_ids: List):
        self.index_id_to_db_id.extend(db_ids)


class DenseFlatIndexer(DenseIndexer):
    """dvl/indexer/faiss_indexers.py:63-87 with the flat index resident on the B200."""

    def __init__(self, vector_sz: int, buffer_size: int = 50000, **index_kw):
        super(DenseFlatIndexer, self).__init__(buffer_size=buffer_size)
        self.index = FlatIPIndex(vector_sz, **index_kw)

    def index_data(self, data: List[Tuple[object, np.array]]):
        n = len(data)
        for i in range(0, n, self.buffer_size):
            chunk = data[i:i + self.buffer_size]
            db_ids = [t[0] for t in chunk]
            if isinstance(chunk[0][1], torch.Tensor):
                vectors = torch.stack([t[1].reshape(-1) for t in chunk], dim=0)
            else:
                vectors = np.concatenate([np.reshape(t[1], (1, -1)) for t in chunk], axis=0)
            self._update_id_mapping(db_ids)
            self.index.add(vectors)
        indexed_cnt = len(self.index_id_to_db_id)
        logger.info('Total data indexed %d', indexed_cnt)

    def index_matrix(self, db_ids: List[object], vectors):
        """Bulk variant of index_data for embeddings that are already one [n, d] matrix (numpy or cuda tensor)."""
        assert len(db_ids) == vectors.shape[0]
        self._update_id_mapping(list(db_ids))
        self.index.add(vectors)

    def search_knn(self, query_vectors: np.array, top_docs: int) -> List[Tuple[List[object], List[float]]]:
        scores, indexes = self.index.search(query_vectors, top_docs)
        return self._format_result(scores, indexes)

    def _format_result(self, scores, indexes):
        # convert to external ids (a label of -1 - index shorter than top_docs - maps to the LAST id through
        # negative indexing, exactly like faiss_indexers.py:85).  One vectorised gather over an object array
        # instead of nq * k Python list look-ups.
        if not len(self.index_id_to_db_id):
            return [([], scores[i]) for i in range(len(indexes))]
        gather = _pyhost_gather()
        if gather is not None and isinstance(self.index_id_to_db_id, list):
            # one C loop over the label matrix (hostext/pylists.c): nq lists of k references, nothing in between
            idx = np.ascontiguousarray(indexes, dtype=np.int64)
            gc_was_on = gc.isenabled()
            gc.disable()
            try:
                rows = gather(self.index_id_to_db_id, idx.ctypes.data, idx.shape[0], idx.shape[1])
                return list(zip(rows, scores))
            finally:
                if gc_was_on:
                    gc.enable()
        id_map = self._id_array()
        # building nq lists of k references allocates ~nq container objects: keep the cyclic GC from re-scanning the
        # (possibly million-entry) id list while they are created
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            return list(zip(id_map[indexes].tolist(), scores))
        finally:
            if gc_was_on:
                gc.enable()

    def _id_array(self):
        cache = getattr(self, "_id_cache", None)
        if cache is None or cache[0] != len(self.index_id_to_db_id):
            arr = np.fromiter(self.index_id_to_db_id, dtype=object, count=len(self.index_id_to_db_id))
            cache = (len(self.index_id_to_db_id), arr)
            self._id_cache = cache
        return cache[1]


class DenseHNSWFlatIndexer(DenseIndexer):
    """Approximate HNSW search (faiss_indexers.py:90-154) is outside the exact-search hot path; the name stays
    importable because dvl/trainer.py:14 imports it."""

    def __init__(self, *a, **kw):
        raise NotImplementedError("DenseHNSWFlatIndexer (--hnsw_index) is not part of the B200 exact-search path; "
                                  "use DenseFlatIndexer")
