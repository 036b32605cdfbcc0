"""Workload for ncu: one warm-up + `reps` searches.  python scripts/profile_search.py <n> <nq> [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightningdot_b200.indexer import FlatIPIndex  # noqa: E402

n, nq = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
d, k = 768, 100
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(n, d, device="cuda", generator=g) / d ** 0.5
q = torch.randn(nq, d, device="cuda", generator=g) / d ** 0.5
idx = FlatIPIndex(d)
idx.add(x)
for _ in range(1 + reps):
    idx.search_device(q, k, resolve_flags=False)
torch.cuda.synchronize()
print("done", idx.last_flagged)
