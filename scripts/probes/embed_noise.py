"""Probe: embeddings of the configs[0]-sized synthetic sets (5000 captions, 1000 images, 12-layer seeded towers) from the
CUDA towers in bf16 and fp16, saved for comparison with the reference's fp32 CPU embeddings (rank-margin planning of
the planted Recall fixture, oracle/make_golden.py --configs0)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lightningdot_b200 import synth  # noqa: E402
from lightningdot_b200.bi_encoder import BertEncoder, TowerConfig, UniterEncoder  # noqa: E402

n_img, cpi = 1000, 5
tb = synth.text_batch(n_img * cpi, 32, seed=0, ragged=True)
ib = synth.image_batch(n_img, 36, seed=0, ragged=True)
ib["img_feat"] = ib["img_feat"].half().float()
ib["img_pos_feat"][..., :6] = ib["img_pos_feat"][..., :6].half().float()
ib["img_pos_feat"][..., 6] = ib["img_pos_feat"][..., 4] * ib["img_pos_feat"][..., 5]
os.makedirs("gpurun_out", exist_ok=True)
for dt, tag in ((torch.bfloat16, "bf16"), (torch.float16, "fp16")):
    mt = BertEncoder(TowerConfig(vocab_size=synth.VOCAB), project_dim=768)
    mt.load_state_dict(synth.random_tower_state("txt", seed=42, perturb=True, layers=12), strict=True)
    mi = UniterEncoder(TowerConfig(vocab_size=synth.VOCAB), project_dim=768)
    mi.load_state_dict(synth.random_tower_state("img", seed=43, perturb=True, layers=12), strict=True)
    mt.compute_dtype = mi.compute_dtype = dt
    mt.cuda().eval()
    mi.cuda().eval()
    with torch.no_grad():
        T = mt(tb["input_ids"].cuda(), tb["attention_mask"].cuda(), tb["position_ids"].cuda(), need_sequence=False)[1]
        I = mi(ib["input_ids"].cuda(), ib["attention_mask"].cuda(), ib["position_ids"].cuda(), ib["img_feat"].cuda(),
               ib["img_pos_feat"].cuda(), None, ib["gather_index"].cuda(), need_sequence=False)[1]
    np.save(f"gpurun_out/T_{tag}.npy", T.float().cpu().numpy())
    np.save(f"gpurun_out/I_{tag}.npy", I.float().cpu().numpy())
    print(tag, T.shape, I.shape, float(T.norm(dim=1).mean()))
