"""Debug probe: with lr = 0 (weights fixed, dropout off) the graphed step and the eager step must report the same loss for
every batch; prints both sequences (twice for eager, to see its own run-to-run spread)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from lightningdot_b200 import synth
from lightningdot_b200.bi_encoder import BiEncoder, BiEncoderNllLoss, TowerConfig, get_optimizer
from lightningdot_b200.training import GraphedTrainStep
from lightningdot_b200.utils import _calc_loss

B, steps = 8, 7
lr = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
largs = types.SimpleNamespace(caption_score_weight=0.0)
batches = [{"txts": synth.text_batch(B, 24, seed=10 + s, ragged=True), "imgs": synth.image_batch(B, 20, seed=30 + s, ragged=True),
            "caps": {"input_ids": None}, "sample_size": B, "pos_ctx_indices": list(range(B)), "neg_ctx_indices": []}
           for s in range(steps)]

def make():
    args = types.SimpleNamespace(img_model_type='uniter-base', img_model_config=TowerConfig(num_hidden_layers=2),
                                 img_checkpoint=None, txt_model_type='bert-base',
                                 txt_model_config=TowerConfig(num_hidden_layers=2), txt_checkpoint=None)
    torch.manual_seed(3)
    model = BiEncoder(args, project_dim=768)
    opt = get_optimizer(model, learning_rate=lr, adam_eps=1e-4, weight_decay=0.01)
    opt.max_grad_norm = 2.0
    return model.cuda().eval(), opt

def fb(model):
    def run(bt):
        t, i, _ = model(bt)
        l1, _, _ = _calc_loss(largs, BiEncoderNllLoss(), i, t, None, bt["pos_ctx_indices"], None)
        l2, _, _ = _calc_loss(largs, BiEncoderNllLoss(), t, i, None, bt["pos_ctx_indices"], None)
        loss = 0.5 * l1 + 0.5 * l2
        loss.backward()
        return loss
    return run

for rep in range(2):
    m, o = make()
    r = fb(m)
    out = []
    for bt in batches:
        out.append(r(bt).item()); o.step(); o.zero_grad()
    print("eager  ", " ".join(f"{v:.6f}" for v in out))
m, o = make()
g = GraphedTrainStep(fb(m), o, batches[0], warmup=1, layout_step=False)
out = [v.item() for v in g.warmup_losses]
for bt in batches[1:]:
    out.append(g(bt).item())
print("graphed", " ".join(f"{v:.6f}" for v in out))
out = [g(bt).item() for bt in batches]
print("graphed", " ".join(f"{v:.6f}" for v in out), "(second pass over the batches)")
