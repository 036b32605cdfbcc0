"""Debug probe: loss sequences of the same 6 steps under: eager; eager with device-side hyper-parameters; graphed; and the
eager / graphed pair without gradient clipping."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from lightningdot_b200 import synth
from lightningdot_b200.bi_encoder import BiEncoder, BiEncoderNllLoss, TowerConfig, get_optimizer, get_schedule_linear
from lightningdot_b200.training import GraphedTrainStep
from lightningdot_b200.utils import _calc_loss

B, steps = 8, 6
lr = 2e-6
largs = types.SimpleNamespace(caption_score_weight=0.0)
batches = [{"txts": synth.text_batch(B, 24, seed=10 + s, ragged=True), "imgs": synth.image_batch(B, 20, seed=30 + s, ragged=True),
            "caps": {"input_ids": None}, "sample_size": B, "pos_ctx_indices": list(range(B)), "neg_ctx_indices": []}
           for s in range(steps)]

def make(clip):
    args = types.SimpleNamespace(img_model_type='uniter-base', img_model_config=TowerConfig(num_hidden_layers=2),
                                 img_checkpoint=None, txt_model_type='bert-base',
                                 txt_model_config=TowerConfig(num_hidden_layers=2), txt_checkpoint=None)
    torch.manual_seed(3)
    model = BiEncoder(args, project_dim=768)
    opt = get_optimizer(model, learning_rate=lr, adam_eps=1e-4, weight_decay=0.01)
    opt.max_grad_norm = clip
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

def eager(clip, dev_hyper=False, reload=False):
    m, o, s = make(clip)
    if dev_hyper:
        o.device_hyper(True)
    r = fb(m)
    out = []
    for bt in batches:
        if reload:   # force the towers to re-convert every weight from the fp32 parameters
            m.txt_model._engine = None
            m.img_model._engine = None
        out.append(r(bt).item()); o.step(); s.step(); o.zero_grad()
    return out, m

def graphed(clip):
    m, o, s = make(clip)
    g = GraphedTrainStep(fb(m), o, batches[0], scheduler=s, warmup=1, layout_step=False)
    out = [g.warmup_losses[0].item()] + [g(bt).item() for bt in batches[1:]]
    return out, m

def show(tag, seq):
    print(f"{tag:28s}", " ".join(f"{v:.6f}" for v in seq))

e, me = eager(2.0); show("eager clip 2", e)
print("   eager engines aliased:", me.txt_model.engine().aliased, me.img_model.engine().aliased)
show("eager clip 2, reload/step", eager(2.0, reload=True)[0])
show("eager clip 2, device hyper", eager(2.0, dev_hyper=True)[0])
show("graph clip 2", graphed(2.0)[0])
show("eager no clip", eager(0.0)[0])
show("eager no clip, reload/step", eager(0.0, reload=True)[0])
show("graph no clip", graphed(0.0)[0])
