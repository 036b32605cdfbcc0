"""Debug probe: eager plain optimiser (A) vs eager device-hyper optimiser (B), lockstep: gradient / state differences."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from lightningdot_b200 import synth, _lib
from lightningdot_b200.bi_encoder import BiEncoder, BiEncoderNllLoss, TowerConfig, get_optimizer, get_schedule_linear
from lightningdot_b200.utils import _calc_loss

B, steps = 8, 4
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
    opt.max_grad_norm = 0.0
    return model.cuda().eval(), opt, get_schedule_linear(opt, 3, 50)

def fb(model, bt):
    t, i, _ = model(bt)
    l1, _, _ = _calc_loss(largs, BiEncoderNllLoss(), i, t, None, bt["pos_ctx_indices"], None)
    l2, _, _ = _calc_loss(largs, BiEncoderNllLoss(), t, i, None, bt["pos_ctx_indices"], None)
    loss = 0.5 * l1 + 0.5 * l2
    loss.backward()
    return loss.item()

ma, oa, sa = make()
mb, ob, sb = make()
ob.device_hyper(True)
def engine_diff():
    out = []
    for tag in ("txt_model", "img_model"):
        ea, eb = getattr(ma, tag).engine(), getattr(mb, tag).engine()
        for k in ea.w:
            ta, tb = ea.w[k], eb.w[k]
            if torch.is_tensor(ta):
                d = (ta.float() - tb.float()).abs().max().item()
                if d > 1e-6:
                    out.append((d, tag + "." + k))
    sda, sdb = ma.state_dict(), mb.state_dict()
    sd = max(((sda[k].float() - sdb[k].float()).abs().max().item(), k) for k in sda)
    return sorted(out, reverse=True)[:6], sd

def fresh_loss(src, bt):
    args = types.SimpleNamespace(img_model_type='uniter-base', img_model_config=TowerConfig(num_hidden_layers=2),
                                 img_checkpoint=None, txt_model_type='bert-base',
                                 txt_model_config=TowerConfig(num_hidden_layers=2), txt_checkpoint=None)
    f = BiEncoder(args, project_dim=768)
    f.load_state_dict(src.state_dict())
    f.cuda().eval()
    with torch.no_grad():
        t, i, _ = f(bt)
        l1, _, _ = _calc_loss(largs, BiEncoderNllLoss(), i, t, None, bt["pos_ctx_indices"], None)
        l2, _, _ = _calc_loss(largs, BiEncoderNllLoss(), t, i, None, bt["pos_ctx_indices"], None)
    return (0.5 * l1 + 0.5 * l2).item()

for s, bt in enumerate(batches):
    ed, sd = engine_diff()
    print(f"step {s} before forward: engine tensor diffs {ed}; worst state_dict diff {sd}; fresh-model loss from A {fresh_loss(ma, bt):.6f} from B {fresh_loss(mb, bt):.6f}")
    la, lb = fb(ma, bt), fb(mb, bt)
    pb = dict(mb.named_parameters())
    gd = max(((p.grad - pb[n].grad).abs().max().item(), n) for n, p in ma.named_parameters() if p.grad is not None)
    print(f"step {s}: loss {la:.6f} {lb:.6f}; worst grad diff {gd[0]:.3e} ({gd[1]}); lr {oa.param_groups[0]['lr']:.4e} {ob.param_groups[0]['lr']:.4e}")
    oa.step(); ob.step()
    torch.cuda.synchronize()
    print("    hyper B:", [ob._hyper[g].cpu().tolist() for g in sorted(ob._hyper)], "steps", oa._steps, ob._steps)
    for gi in range(len(oa._flat)):
        fa, fb_ = oa._flat[gi], ob._flat[gi]
        if fa is None:
            continue
        dp = (fa['p'] - fb_['p']).abs()
        i = int(dp.argmax())
        print(f"    group {gi}: |dm| {(fa['m'] - fb_['m']).abs().max().item():.3e} |dv| {(fa['v'] - fb_['v']).abs().max().item():.3e} "
              f"|dp| {dp.max().item():.3e} at {i}: pA {fa['p'][i].item():.9e} pB {fb_['p'][i].item():.9e} g {fa['g'][i].item():.3e} "
              f"m {fa['m'][i].item():.3e} v {fa['v'][i].item():.3e}; |dp16| {(fa['p16'].float() - fb_['p16'].float()).abs().max().item() if fa['p16'] is not None else 0:.3e}")
    sa.step(); sb.step(); oa.zero_grad(); ob.zero_grad()
