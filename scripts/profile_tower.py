"""Workload for ncu: text tower over `b` captions of length 32 (one warm-up + `reps` passes).
python scripts/profile_tower.py <b> [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightningdot_b200 import synth  # noqa: E402
from lightningdot_b200.towers import TowerEngine  # noqa: E402

b = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sd = synth.random_tower_state("txt", seed=42, layers=12)
eng = TowerEngine("txt", 768, 12, 3072, 12, dtype=torch.bfloat16)
eng.load(sd, "cuda")
t = synth.text_batch(b, 32, seed=3)
ids, mask, pos = t["input_ids"].cuda(), t["attention_mask"].cuda(), t["position_ids"].cuda()
for _ in range(1 + reps):
    eng.encode_text(ids, mask, pos)
torch.cuda.synchronize()
print("done")
