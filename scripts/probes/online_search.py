"""Probe: per-kernel device time of one small-batch search (nq queries against n rows) from the library's own event
accounting - the online / sharded regime (BASELINE configs[3] shard: 125 000 rows, 128 queries)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lightningdot_b200 import _lib  # noqa: E402
from lightningdot_b200.indexer import FlatIPIndex  # noqa: E402

for n, nq in ((125000, 128), (1000000, 128), (125000, 1250), (1000000, 10000)):
    d, k = 768, 100
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(n, d, device="cuda", generator=g) / d ** 0.5
    q = torch.randn(nq, d, device="cuda", generator=g) / d ** 0.5
    idx = FlatIPIndex(d)
    idx.add(x)
    for _ in range(3):
        idx.search_device(q, k, resolve_flags=False)
    torch.cuda.synchronize()
    _lib.prof_reset()
    _lib.prof_enable(True)
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        idx.search_device(q, k, resolve_flags=False)
    e1.record()
    torch.cuda.synchronize()
    _lib.prof_enable(False)
    p = _lib.prof_read()
    per = {name: round(1e3 * v["ms"] / reps, 1) for name, v in p.items() if v["launches"]}
    print(f"n={n} nq={nq}: {1e3 * e0.elapsed_time(e1) / reps:.1f} us per search (events on), kernels us: {per}, "
          f"flagged {idx.last_flagged}", flush=True)
    del idx, x
