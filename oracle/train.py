"""TEST INFRASTRUCTURE (parity oracle) - one train_itm.py step on the CPU in fp32: both towers, the symmetric in-batch
NLL and the gradients of every parameter, by torch autograd over the functional restatement in oracle/towers.py.

Pinned by oracle/make_golden.py: the UNMODIFIED reference modules (BertEncoder / UniterEncoder in eval() mode so that
dropout is off, BiEncoderNllLoss and _calc_loss imported from /root/reference) run the same step with
loss.backward(); this restatement must reproduce the reference's loss and parameter gradients, and the reference's
values are stored in tests/golden/train_step_*.npz.

Follows train_itm.py:191-222,252-258 (forward, the two _calc_loss calls, 0.5 / 0.5 mix, backward).
"""
import torch

from . import loss as oloss
from . import towers

SAMPLES = 8


def sample_index(numel):
    """The fixed entries of a flattened gradient that the golden fixture stores."""
    return (torch.arange(SAMPLES, dtype=torch.int64) * 7919 + 13) % numel


def train_step(sd_txt, sd_img, txt, img):
    """-> (loss, correct, {name: grad} for the text tower, {name: grad} for the image tower); parameters the forward
    does not read (pooler, mask_embedding) get no entry, as torch leaves their .grad None."""
    pt = {k: v.clone().requires_grad_(True) for k, v in sd_txt.items()}
    pi = {k: v.clone().requires_grad_(True) for k, v in sd_img.items()}
    _, t = towers.text_tower(pt, txt["input_ids"], txt["attention_mask"], txt["position_ids"])
    _, i = towers.image_tower(pi, img["input_ids"], img["attention_mask"], img["position_ids"], img["img_feat"],
                              img["img_pos_feat"], img["gather_index"])
    loss, correct = oloss.symmetric_nll(t, i)
    loss.backward()
    # nn.Embedding(padding_idx=0) never accumulates a gradient for row 0 (uniter_model/model/model.py:221-222)
    for p in (pt, pi):
        g = p["bert.embeddings.word_embeddings.weight"].grad
        if g is not None:
            g[0].zero_()
    gt = {k: v.grad for k, v in pt.items() if v.grad is not None}
    gi = {k: v.grad for k, v in pi.items() if v.grad is not None}
    return loss.detach(), correct, gt, gi


def tower_vjp(kind, sd, batch, upstream, drop=None):
    """-> (pooled, {name: grad}) of one tower for a given upstream gradient d(pooled): the vector-Jacobian product the
    tower's backward computes, isolated from the loss (whose softmax amplifies forward rounding).  drop: training-mode
    dropout hook (oracle/dropout.py: Dropper), None = eval mode."""
    p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    kw = {} if drop is None else {"drop": drop}
    if kind == "txt":
        _, pooled = towers.text_tower(p, batch["input_ids"], batch["attention_mask"], batch["position_ids"], **kw)
    else:
        _, pooled = towers.image_tower(p, batch["input_ids"], batch["attention_mask"], batch["position_ids"],
                                       batch["img_feat"], batch["img_pos_feat"], batch["gather_index"], **kw)
    pooled.backward(upstream)
    g = p["bert.embeddings.word_embeddings.weight"].grad
    if g is not None:
        g[0].zero_()   # padding_idx=0, as in train_step
    return pooled.detach(), {k: v.grad for k, v in p.items() if v.grad is not None}
