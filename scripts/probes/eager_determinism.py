"""Debug probe: two identically initialised models stepped in lockstep on the same batches (eager, dropout off): per step,
the worst gradient and parameter differences between the two runs (run-to-run noise of the training kernels)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from lightningdot_b200 import synth
from lightningdot_b200.bi_encoder import BiEncoder, BiEncoderNllLoss, TowerConfig, get_optimizer
from lightningdot_b200.utils import _calc_loss

B, steps = 8, 6
lr = float(sys.argv[1]) if len(sys.argv) > 1 else 2e-6
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

def fb(model, bt):
    t, i, _ = model(bt)
    l1, _, _ = _calc_loss(largs, BiEncoderNllLoss(), i, t, None, bt["pos_ctx_indices"], None)
    l2, _, _ = _calc_loss(largs, BiEncoderNllLoss(), t, i, None, bt["pos_ctx_indices"], None)
    loss = 0.5 * l1 + 0.5 * l2
    loss.backward()
    return loss.item(), t.detach().clone(), i.detach().clone()

ma, oa = make()
mb, ob = make()
for s, bt in enumerate(batches):
    la, ta, ia = fb(ma, bt)
    lb, tb, ib = fb(mb, bt)
    print(f"step {s}: loss {la:.6f} {lb:.6f}  |dt| {(ta - tb).abs().max().item():.3e} |di| {(ia - ib).abs().max().item():.3e}")
    worst = []
    pb = dict(mb.named_parameters())
    for n, p in ma.named_parameters():
        if p.grad is None:
            continue
        d = (p.grad - pb[n].grad).norm().item()
        w = p.grad.norm().item()
        worst.append((d / (w + 1e-12), d, w, n))
    worst.sort(reverse=True)
    for r, d, w, n in worst[:3]:
        print(f"    grad {n}: |d| {d:.3e} |g| {w:.3e} rel {r:.3e}")
    oa.step(); oa.zero_grad(); ob.step(); ob.zero_grad()
    wp = max(((p.detach() - pb[n].detach()).abs().max().item(), n) for n, p in ma.named_parameters())
    print(f"    after step: worst param diff {wp[0]:.3e} ({wp[1]})")
