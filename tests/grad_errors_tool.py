"""Per-parameter gradient error of the GPU training step against the CPU oracle (diagnostic).
python tests/grad_errors_tool.py [bf16|fp16] [layers] [batch] [head_scale] [loss_scale]
(lives under tests/: it runs the oracle as the checker)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from lightningdot_b200 import synth  # noqa: E402
from oracle import train as otrain  # noqa: E402
import test_gpu_training as T  # noqa: E402

dtype = torch.float16 if len(sys.argv) > 1 and sys.argv[1] == "fp16" else torch.bfloat16
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 6
head_scale = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
loss_scale = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
seed = 201 if not head_scale else 401
if head_scale:
    sd_t = synth.conditioned_tower_state("txt", seed=seed, head_scale=head_scale, layers=layers)
    sd_i = synth.conditioned_tower_state("img", seed=seed + 1, head_scale=head_scale, layers=layers)
else:
    sd_t = synth.random_tower_state("txt", seed=seed, perturb=True, layers=layers)
    sd_i = synth.random_tower_state("img", seed=seed + 1, perturb=True, layers=layers)
tb = synth.text_batch(batch, 32, seed=seed, ragged=True)
ib = synth.image_batch(batch, 36, seed=seed, ragged=True)
mt, mi = T.towers(layers, sd_t, sd_i)
mt.compute_dtype = mi.compute_dtype = dtype
loss, correct = T.gpu_step(mt, mi, tb, ib, batch)
(loss * loss_scale).backward()
for m_ in (mt, mi):
    for p_ in m_.parameters():
        if p_.grad is not None:
            p_.grad.div_(loss_scale)
oloss, _, gt, gi = otrain.train_step(sd_t, sd_i, tb, ib)
print("loss", loss.item(), "oracle", oloss.item())
for tag, m, want in (("txt", mt, gt), ("img", mi, gi)):
    for n, p in m.named_parameters():
        if n in want:
            g, w = p.grad.cpu(), want[n]
            cos = torch.nn.functional.cosine_similarity(g.reshape(1, -1), w.reshape(1, -1)).item()
            print(f"{tag} {n:60s} |g| {w.norm().item():10.3e} rel {((g - w).norm() / w.norm().clamp_min(1e-30)).item():8.4f} cos {cos:.5f}")
num = den = 0.0
for tag, m, want in (("txt", mt, gt), ("img", mi, gi)):
    for n, p in m.named_parameters():
        if n in want:
            num += float((p.grad.cpu() - want[n]).norm()) ** 2
            den += float(want[n].norm()) ** 2
print(f"GLOBAL relative L2 over all parameters: {(num / den) ** 0.5:.4f}")
