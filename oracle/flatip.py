"""TEST INFRASTRUCTURE (parity oracle) - CPU restatement of the reference's exact inner-product search.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

PARITY UNPINNED for the arithmetic: the reference delegates to faiss-cpu==1.6.3 (DVL.yml:80), which is not under
/root/reference and not installable offline, and the reference ships no golden vectors for this path.  The
restatement is anchored on the reference's call sites: dvl/indexer/faiss_indexers.py:67 (IndexFlatIP), :77 (add),
:83 (search), :85-87 (id remap and result format); dvl/trainer.py:160-171 (callers).  The Python wrapper logic
(id mapping, 50 000-row buffering, result format) IS pinned: oracle/make_golden.py runs the reference's own
DenseFlatIndexer over a numpy stand-in for faiss and the fixture is checked in tests/.

faiss IndexFlatIP.search, as published: fp32 inner products - BLAS sgemm over (4096 queries x 1024 rows) blocks
when nq >= 20, a SIMD dot-product loop otherwise - pushed through a per-query binary min-heap of size k and
returned in descending score order, labels int64, missing results (k > ntotal) as label -1 / score -FLT_MAX.
The order of equal scores inside the result is heap dependent (unspecified); this oracle - and the product -
define it as (score desc, row id asc).

Two scorers:
  scores_f64   fp32 inputs, fp64 accumulation, rounded to fp32: the correctly-rounded inner product, independent
               of summation order.  This is the parity bar (ids bit-exact, scores bit-exact / within 1e-3 rel).
  scores_f32   plain fp32 sgemm, i.e. what faiss itself computes (order-dependent rounding ~1e-7 relative).
               Used for the CPU baseline timing and to show how far a real fp32 BLAS run sits from the f64 bar.
"""
import numpy as np

NEG = np.float32(-3.4028235e38)


def scores_f64(q, x):
    return (q.astype(np.float64) @ x.astype(np.float64).T).astype(np.float32)


def scores_f32(q, x):
    return np.ascontiguousarray(q, np.float32) @ np.ascontiguousarray(x, np.float32).T


def rank_topk(scores, k):
    """(score desc, row id asc) top-k of a [nq, n] fp32 score matrix -> (scores [nq,k] f32, ids [nq,k] i64)."""
    nq, n = scores.shape
    kk = min(k, n)
    if kk < n:
        # keep everything >= the kk-th largest value (ties included), then order exactly
        part = np.partition(scores, n - kk, axis=1)[:, n - kk]
        out_s = np.full((nq, k), NEG, np.float32)
        out_i = np.full((nq, k), -1, np.int64)
        for r in range(nq):
            cand = np.nonzero(scores[r] >= part[r])[0]
            order = np.lexsort((cand, -scores[r, cand].astype(np.float64)))[:kk]
            out_i[r, :kk] = cand[order]
            out_s[r, :kk] = scores[r, cand[order]]
        return out_s, out_i
    ids = np.broadcast_to(np.arange(n, dtype=np.int64), scores.shape)
    order = np.lexsort((ids, -scores.astype(np.float64)), axis=1)
    out_s = np.full((nq, k), NEG, np.float32)
    out_i = np.full((nq, k), -1, np.int64)
    out_s[:, :n] = np.take_along_axis(scores, order, axis=1)
    out_i[:, :n] = order
    return out_s, out_i


def search(q, x, k, scorer=scores_f64, q_block=1024):
    """IndexFlatIP.search restated: -> (scores [nq, k] float32, labels [nq, k] int64)."""
    q = np.ascontiguousarray(q, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out_s = np.empty((len(q), k), np.float32)
    out_i = np.empty((len(q), k), np.int64)
    for b in range(0, len(q), q_block):
        s, i = rank_topk(scorer(q[b:b + q_block], x), k)
        out_s[b:b + q_block] = s
        out_i[b:b + q_block] = i
    return out_s, out_i


def search_blocked(q, x, k, scorer=scores_f64, row_block=125000):
    """search() for an index too large to hold a full [nq, n] score matrix: exact top-k per block of rows, then the
    top-k of the per-block winners.  Candidates are concatenated in row order, so the (score desc, position asc) rule
    of rank_topk is still (score desc, row id asc).  Same result as search()."""
    q = np.ascontiguousarray(q, np.float32)
    cs, ci = [], []
    for r in range(0, len(x), row_block):
        s, i = rank_topk(scorer(q, np.ascontiguousarray(x[r:r + row_block], np.float32)), k)
        cs.append(s)
        ci.append(np.where(i >= 0, i + r, -1))
    cs, ci = np.concatenate(cs, axis=1), np.concatenate(ci, axis=1)
    keep = ci >= 0
    s2, pos = rank_topk(np.where(keep, cs, NEG), k)
    ids = np.take_along_axis(ci, np.maximum(pos, 0), axis=1)
    valid = (pos >= 0) & np.take_along_axis(keep, np.maximum(pos, 0), axis=1)
    return np.where(valid, s2, NEG).astype(np.float32), np.where(valid, ids, -1)


class FlatIndexer:
    """DenseFlatIndexer restated (dvl/indexer/faiss_indexers.py:63-87): id list + flat index + search_knn."""

    def __init__(self, vector_sz, buffer_size=50000, scorer=scores_f64):
        self.buffer_size = buffer_size
        self.index_id_to_db_id = []
        self.xb = np.zeros((0, vector_sz), np.float32)
        self.scorer = scorer

    def index_data(self, data):
        for i in range(0, len(data), self.buffer_size):
            chunk = data[i:i + self.buffer_size]
            self.index_id_to_db_id.extend(t[0] for t in chunk)
            self.xb = np.concatenate([self.xb] + [np.reshape(t[1], (1, -1)).astype(np.float32) for t in chunk], axis=0)

    def search_knn(self, query_vectors, top_docs):
        scores, idx = search(query_vectors, self.xb, top_docs, self.scorer)
        # faiss_indexers.py:85 - label -1 (short index) silently maps to the LAST id through Python indexing
        db_ids = [[self.index_id_to_db_id[i] for i in row] for row in idx]
        return [(db_ids[i], scores[i]) for i in range(len(db_ids))]
