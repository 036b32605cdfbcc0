"""Debug probe: eager model A vs graphed model B (same init, same batches, same schedule): parameter / moment differences
after every update."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from lightningdot_b200 import synth
from lightningdot_b200.bi_encoder import BiEncoder, BiEncoderNllLoss, TowerConfig, get_optimizer, get_schedule_linear
from lightningdot_b200.training import GraphedTrainStep
from lightningdot_b200.utils import _calc_loss

B, steps = 8, 5
lr = 2e-6
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
    return model.cuda().eval(), opt, get_schedule_linear(opt, 3, 50)

def fb(model):
    def run(bt):
        t, i, _ = model(bt)
        l1, _, _ = _calc_loss(largs, BiEncoderNllLoss(), i, t, None, bt["pos_ctx_indices"], None)
        l2, _, _ = _calc_loss(largs, BiEncoderNllLoss(), t, i, None, bt["pos_ctx_indices"], None)
        loss = 0.5 * l1 + 0.5 * l2
        loss.backward()
        return loss
    return run

ma, oa, sa = make()
mb, ob, sb = make()
ra = fb(ma)
g = None
for s, bt in enumerate(batches):
    la = ra(bt).item()
    ga = {n: p.grad.clone() for n, p in ma.named_parameters() if p.grad is not None}
    oa.step(); sa.step(); oa.zero_grad()
    if g is None:
        g = GraphedTrainStep(fb(mb), ob, bt, scheduler=sb, warmup=1, layout_step=False)
        lb = g.warmup_losses[0].item()
    else:
        lb = g(bt).item()
    pb = dict(mb.named_parameters())
    diffs = sorted((((p.detach() - pb[n].detach()).abs().max().item(), n) for n, p in ma.named_parameters()), reverse=True)
    print(f"step {s}: loss eager {la:.6f} graph {lb:.6f}; lr {oa.param_groups[0]['lr']:.3e} {ob.param_groups[0]['lr']:.3e}; "
          f"steps {oa._steps} {ob._steps}; hyper {[round(v, 6) for v in ob._hyper[0].cpu().tolist()]}")
    for d, n in diffs[:4]:
        print(f"    param {n}: max |dp| {d:.3e}")
    fa, fbf = oa._flat, ob._flat
    for gi in range(len(fa)):
        if fa[gi] is None:
            continue
        print(f"    group {gi}: |dm| {(fa[gi]['m'] - fbf[gi]['m']).abs().max().item():.3e} |dv| {(fa[gi]['v'] - fbf[gi]['v']).abs().max().item():.3e} "
              f"|dp| {(fa[gi]['p'] - fbf[gi]['p']).abs().max().item():.3e} |dg| {fbf[gi]['g'].abs().max().item():.3e}")
