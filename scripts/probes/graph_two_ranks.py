"""Debug probe (torchrun --nproc-per-node 2): GraphedTrainStep under a 2-rank NCCL group on a small model, with a
faulthandler watchdog that prints where every thread is if the run has not finished after 60 s."""
import faulthandler
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
faulthandler.dump_traceback_later(int(os.environ.get("PROBE_WATCHDOG_S", "60")), exit=True)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from lightningdot_b200 import synth  # noqa: E402
from lightningdot_b200.bi_encoder import (BiEncoder, BiEncoderNllLoss, TowerConfig, get_optimizer,  # noqa: E402
                                          get_schedule_linear, setup_for_distributed_mode)
from lightningdot_b200.training import GraphedTrainStep  # noqa: E402
from lightningdot_b200.utils import _calc_loss  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
b, layers, steps = 16, 2, 5
B = b * world
torch.manual_seed(9)
cfg = dict(img_model_type='uniter-base', img_model_config=TowerConfig(vocab_size=synth.VOCAB, num_hidden_layers=layers),
           img_checkpoint=None, txt_model_type='bert-base',
           txt_model_config=TowerConfig(vocab_size=synth.VOCAB, num_hidden_layers=layers), txt_checkpoint=None)
model = BiEncoder(types.SimpleNamespace(**cfg), project_dim=768)
opt = get_optimizer(model, learning_rate=2e-6, adam_eps=1e-4, weight_decay=0.01)
opt.max_grad_norm = 2.0
model, opt = setup_for_distributed_mode(model, opt, dev, 1, rank, False)
model.train()
opt.overlap_grad_sync = os.environ.get("PROBE_OVERLAP", "1") == "1"
opt.early_sync_bytes = int(os.environ.get("PROBE_EARLY_MB", "4")) << 20
sched = get_schedule_linear(opt, 2, 50)
la = types.SimpleNamespace(caption_score_weight=0.0, distributed_world_size=world)


def batch(seed):
    tb, ib = synth.text_batch(B, 32, seed=seed, ragged=True), synth.image_batch(B, 36, seed=seed + 50, ragged=True)
    lo, hi = rank * b, rank * b + b

    def sl(d):
        return {k: (v[lo:hi].contiguous() if (torch.is_tensor(v) and v.shape[0] == B) else v) for k, v in d.items()}
    return {"txts": sl(tb), "imgs": sl(ib), "caps": {"input_ids": None}, "pos_ctx_indices": list(range(b))}


def fwd_bwd(bt):
    t, i, _ = model(bt)
    l1, _, _ = _calc_loss(la, BiEncoderNllLoss(), i, t, None, bt["pos_ctx_indices"], None)
    l2, _, _ = _calc_loss(la, BiEncoderNllLoss(), t, i, None, bt["pos_ctx_indices"], None)
    loss = 0.5 * l1 + 0.5 * l2
    loss.backward()
    return loss


print(f"[{rank}] building", flush=True)
gstep = GraphedTrainStep(fwd_bwd, opt, batch(0), scheduler=sched, warmup=1)
print(f"[{rank}] captured; warm-up losses {[round(v.item(), 4) for v in gstep.warmup_losses]}", flush=True)
losses = []
for s in range(2, steps):
    losses.append(gstep(batch(s)).item())
    print(f"[{rank}] replay {s}: {losses[-1]:.4f}", flush=True)
worst = 0.0
for f in opt._flat:
    if f is not None:
        ref = f["p"].clone()
        dist.broadcast(ref, src=0)
        worst = max(worst, (ref - f["p"]).abs().max().item())
print(f"[{rank}] done: losses {json.dumps(losses)} finite {all(np.isfinite(losses))} worst param diff vs rank 0 {worst:.3e}", flush=True)
gstep.release()   # (without this destroy_process_group() blocks: the graph still references the communicator)
faulthandler.cancel_dump_traceback_later()
dist.destroy_process_group()
