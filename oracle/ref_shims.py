"""TEST INFRASTRUCTURE - import the UNMODIFIED reference (/root/reference) in this container.

Only used by oracle/make_golden.py (run here, where /root/reference exists) to pin the oracle restatement and to
mint the fixtures under tests/golden/.  Nothing on the product path, and nothing that runs on the GPU box, imports
this module.

The reference depends on wheels that are not installable offline (faiss, apex, horovod, lmdb, lz4, msgpack_numpy,
toolz/cytoolz, tensorboardX) and on transformers==2.3.0 (the container has 5.x).  The stubs below are the minimum
that lets `dvl.models.bi_encoder`, `dvl.indexer.faiss_indexers`, `dvl.trainer`, `dvl.utils`, `dvl.data.itm` and
`uniter_model.model.{model,layer}` import and run on CPU.
"""
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = "/root/reference"


class NumpyIndexFlatIP:
    """Stand-in for faiss.IndexFlatIP (faiss-cpu==1.6.3, DVL.yml:80; not installable offline).

    faiss computes fp32 inner products with BLAS sgemm in (4096 query x 1024 db) blocks and keeps a per-query
    binary heap; its tie order inside the result is heap dependent.  Restated as fp32 matmul + a stable ordering
    (score desc, row id asc)."""

    def __init__(self, d):
        self.d = d
        self.xb = np.zeros((0, d), dtype=np.float32)

    @property
    def ntotal(self):
        return self.xb.shape[0]

    def add(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        self.xb = np.concatenate([self.xb, x], axis=0)

    def search(self, q, k):
        q = np.ascontiguousarray(q, dtype=np.float32)
        scores = q @ self.xb.T
        n = self.ntotal
        ids = np.broadcast_to(np.arange(n, dtype=np.int64), scores.shape)
        order = np.lexsort((ids, -scores), axis=1)[:, :k]
        out_s = np.take_along_axis(scores, order, axis=1)
        out_i = order.astype(np.int64)
        if k > n:
            pad = k - n
            out_s = np.concatenate([out_s, np.full((len(q), pad), -3.4028235e38, np.float32)], axis=1)
            out_i = np.concatenate([out_i, np.full((len(q), pad), -1, np.int64)], axis=1)
        return out_s, out_i


class _FlatLmdbEnv:
    """Stand-in for lmdb.Environment over the flat record file lightningdot_b200/data.py writes when the `lmdb` wheel is
    absent (records.ldkv): lets the reference's OWN DetectFeatLmdb / TxtLmdb classes read a database directory so that
    the mirror's datasets can be pinned against the reference's (make_golden.dataset_case)."""

    def __init__(self, path, **kw):
        from lightningdot_b200.data import FlatKV
        self._kv = FlatKV(path)

    def begin(self, **kw):
        return self

    def get(self, key=None, default=None):
        v = self._kv.get(bytes(key))
        return default if v is None else v

    def close(self):
        self._kv.close()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Install the stub modules and compatibility patches, then put the reference on sys.path."""
    if getattr(install, "_done", False):
        return
    hvd = _stub("horovod.torch", init=lambda: None, size=lambda: 1, rank=lambda: 0, local_rank=lambda: 0,
                local_size=lambda: 1)
    _stub("horovod", torch=hvd)
    fln = _stub("apex.normalization.fused_layer_norm", FusedLayerNorm=torch.nn.LayerNorm)
    norm = _stub("apex.normalization", fused_layer_norm=fln)
    _stub("apex", normalization=norm)
    _stub("faiss", IndexFlatIP=NumpyIndexFlatIP)
    _stub("lmdb", open=lambda path, **kw: _FlatLmdbEnv(path, **kw))
    frame = _stub("lz4.frame", compress=lambda b: b, decompress=lambda b: b)
    _stub("lz4", frame=frame)
    _stub("msgpack_numpy", patch=lambda: None)
    sandbox = _stub("toolz.sandbox", unzip=lambda s: zip(*s))
    _stub("toolz", sandbox=sandbox)

    def concat(seqs):
        for s in seqs:
            for x in s:
                yield x

    def partition_all(n, seq):
        seq = list(seq)
        for i in range(0, len(seq), n):
            yield tuple(seq[i:i + n])

    _stub("cytoolz", concat=concat, partition_all=partition_all, curry=lambda f: f)
    _stub("tensorboardX", SummaryWriter=object)

    import transformers
    import transformers.optimization as topt

    if not hasattr(topt, "AdamW"):
        topt.AdamW = torch.optim.AdamW
    try:
        import transformers.models.bert.tokenization_bert as tb
        sys.modules.setdefault("transformers.tokenization_bert", tb)
    except Exception:  # tokenizer only needed by the scripts
        pass

    from transformers import BertModel, BertPreTrainedModel

    # bi_encoder.py:91 calls self.init_weights() (transformers 2.x API); in transformers 5.x the entry point is
    # post_init(), which sets up bookkeeping attributes and then calls init_weights() itself.
    _orig_init_weights = BertPreTrainedModel.init_weights

    def init_weights(self):
        if not hasattr(self, "all_tied_weights_keys"):
            self.post_init()
        else:
            _orig_init_weights(self)

    BertPreTrainedModel.init_weights = init_weights
    # bi_encoder.py:110-119 tuple-unpacks BertModel's output (transformers 2.x returned tuples)
    _orig_forward = BertModel.forward

    def forward(self, *a, **kw):
        kw.setdefault("return_dict", False)
        return _orig_forward(self, *a, **kw)

    BertModel.forward = forward

    # The reference's `dvl` and `uniter_model` are NAMESPACE packages (no top-level __init__.py); the repository ships
    # regular packages of the same names (the drop-in boundary), and a regular package wins over a namespace portion
    # wherever it sits on sys.path.  Bind the two names to the reference's directories explicitly.
    import os
    for name in ("dvl", "uniter_model"):
        have = sys.modules.get(name)
        if have is not None and not any(str(p).startswith(REFERENCE_ROOT) for p in getattr(have, "__path__", [])):
            raise RuntimeError(f"'{name}' is already imported from {getattr(have, '__path__', '?')}: the reference must be "
                               "imported in a process that has not imported the repository's namesake packages")
        if have is None:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(REFERENCE_ROOT, name)]
            sys.modules[name] = pkg
    while REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)
    install._done = True
