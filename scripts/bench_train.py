#!/usr/bin/env python
"""Training-step measurement for BASELINE configs[4]: in-batch-negative contrastive step of train_itm.py (both towers
forward + backward, symmetric NLL over the GLOBAL batch via the differentiable embedding all-gather, gradient average,
clip, AdamW) at 512 caption/image pairs per GPU (global batch 4096 on 8 GPUs), L = 32, R = 36, bf16.

    python scripts/bench_train.py [--per-gpu-batch 512] [--steps 5] [--warmup 3] [--check]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_train.py ...

Not the headline bench (bench.py measures queries/s); this reports the training side of the hot path: ms per step,
pairs/s, and the tcgen05 GEMM rate of the step measured live with CUDA events around every launch (ldot_prof_*).
Algorithmic FLOPs: 3 x (5.48 GFLOP per caption + 6.45 GFLOP per image) per pair (forward + dgrad + wgrad).

--check (N > 1): also runs the same global batch on every rank alone and compares loss and gradients of the
distributed step with it (the exactness claim of SURVEY.md 8e).
"""
import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from lightningdot_b200 import _lib, synth  # noqa: E402
from lightningdot_b200.bi_encoder import (BiEncoder, BiEncoderNllLoss, TowerConfig, get_optimizer,  # noqa: E402
                                          get_schedule_linear, setup_for_distributed_mode)
from lightningdot_b200.utils import _calc_loss  # noqa: E402


def default_args(**over):
    a = types.SimpleNamespace(per_gpu_batch=512, seq_len=32, regions=36, layers=12, steps=5, warmup=3, fp16=False, check=False, graph=True)
    a.__dict__.update(over)
    return a


def measure(a, rank, world, local_rank, dev):
    """One measurement of the training step inside an ALREADY initialised process group (bench.py's `train_step` key calls
    this at every N).  -> the result dict on rank 0, None elsewhere."""
    b = a.per_gpu_batch
    B = b * world

    def make_model(seed):
        torch.manual_seed(seed)
        cfg = dict(img_model_type='uniter-base', img_model_config=TowerConfig(vocab_size=synth.VOCAB, num_hidden_layers=a.layers),
                   img_checkpoint=None, txt_model_type='bert-base',
                   txt_model_config=TowerConfig(vocab_size=synth.VOCAB, num_hidden_layers=a.layers), txt_checkpoint=None)
        model = BiEncoder(types.SimpleNamespace(**cfg), project_dim=768)
        opt = get_optimizer(model, learning_rate=1e-5, weight_decay=0.01)
        opt.max_grad_norm = 2.0   # train_itm.py:262-267, fused into the optimiser kernel after the gradient average
        return model, opt

    model, opt = make_model(42)
    model, opt = setup_for_distributed_mode(model, opt, dev, 1, local_rank if world > 1 else -1, a.fp16)
    model.train()
    opt.overlap_grad_sync = world > 1   # one backward per step here: the gradient average may start under the backward
    sched = get_schedule_linear(opt, 10, 1000)
    largs = types.SimpleNamespace(caption_score_weight=0.0, distributed_world_size=world)

    # the global batch, seeded; every rank takes its slice (pinned host memory: the H2D copy is inside the step)
    tb = synth.text_batch(B, a.seq_len, seed=7)
    ib = synth.image_batch(B, a.regions, seed=8)

    def slice_batch(lo, hi, pin=True):
        def sl(d):
            out = {}
            for k, v in d.items():
                if torch.is_tensor(v) and v.shape[0] == B:
                    v = v[lo:hi].contiguous()
                out[k] = v.pin_memory() if (pin and torch.is_tensor(v)) else v
            return out
        n = hi - lo
        return {"txts": sl(tb), "imgs": sl(ib), "caps": {"input_ids": None}, "sample_size": n,
                "pos_ctx_indices": list(range(n)), "neg_ctx_indices": []}

    batch = slice_batch(rank * b, rank * b + b)

    def to_dev(bt):
        out = dict(bt)
        for key in ("txts", "imgs"):
            out[key] = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in bt[key].items()}
        return out

    def fwd_bwd(m, bt, la):
        """train_itm.py:191-258 for one batch on the device: both towers, the symmetric in-batch NLL, loss.backward()"""
        t, i, _ = m(bt)
        l1, c1, _ = _calc_loss(la, BiEncoderNllLoss(), i, t, None, bt["pos_ctx_indices"], None)
        l2, c2, _ = _calc_loss(la, BiEncoderNllLoss(), t, i, None, bt["pos_ctx_indices"], None)
        loss = 0.5 * l1 + 0.5 * l2
        loss.backward()
        return loss

    def step(m, o, s, bt, la, on_device=False):
        loss = fwd_bwd(m, bt if on_device else to_dev(bt), la)
        if o is not None:
            o.step()
            s.step()
            o.zero_grad()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    check = None
    if a.check and world > 1:
        # distributed gradients (averaged over ranks) vs the same global batch on one rank alone; eval() mode = dropout
        # off (the two runs would otherwise draw different masks), gradients are recorded all the same
        model.eval()
        overlap, opt.overlap_grad_sync = opt.overlap_grad_sync, False   # (the single-process leg below must not reduce)
        loss_d = step(model, None, None, batch, largs)
        opt.sync_gradients()
        g_dist = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        loss_mean = loss_d.detach().clone()
        dist.all_reduce(loss_mean, op=dist.ReduceOp.AVG)
        model.zero_grad()
        opt.zero_grad()
        solo = types.SimpleNamespace(caption_score_weight=0.0, distributed_world_size=1)
        loss_s = step(model, None, None, slice_batch(0, B), solo)
        num = den = 0.0
        worst = ("", 0.0)
        for n, p in model.named_parameters():
            if p.grad is None:
                continue
            d, w = (g_dist[n] - p.grad).norm().item(), p.grad.norm().item()
            num, den = num + d * d, den + w * w
            if w > 0 and d / w > worst[1] and w > 1e-6:
                worst = (n, d / w)
        model.train()
        check = {"dropout": "off for the check (eval mode)", "loss_distributed_mean": loss_mean.item(), "loss_single_process": loss_s.item(),
                 "grad_rel_l2_all_params": (num / den) ** 0.5, "worst_param": worst[0], "worst_param_rel": worst[1]}
        model.zero_grad()
        opt.zero_grad()
        opt.overlap_grad_sync = overlap

    # batches arrive through PrefetchLoader (uniter_model/data/loader.py mirror): the H2D copy of step i + 1 (151 MB of
    # region features) runs on a side stream under step i, exactly how eval_itm.py / train_itm.py feed the model
    from lightningdot_b200.loader import PrefetchLoader
    for bt in PrefetchLoader([batch] * a.warmup, dev):
        step(model, opt, sched, bt, largs, on_device=True)
    barrier()
    _lib.prof_reset()
    _lib.prof_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loader = PrefetchLoader([batch] * a.steps, dev)
    barrier()
    e0.record()
    losses = []
    for bt in loader:
        losses.append(step(model, opt, sched, bt, largs, on_device=True).detach())   # (no autograd graph kept alive)
    e1.record()
    barrier()
    _lib.prof_enable(False)
    prof = _lib.prof_read()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / a.steps
    eager_ms_step = ms_step
    # the same step captured as ONE CUDA graph (training.GraphedTrainStep) and replayed: the launch-bound fifth of the
    # eager step goes away.  Batches still arrive through PrefetchLoader (H2D on the side stream) and are copied into the
    # graph's static inputs; lr schedule, Adam bias corrections and dropout masks advance per replay.
    graph_info = None
    # (under a process group the captured step contains NCCL collectives - embedding all-gather / reduce-scatter, the
    # overlapped gradient all-reduce; the graph is released below, before anything can destroy the process group)
    want_graph = getattr(a, "graph", True)
    if want_graph:
        from lightningdot_b200.training import GraphedTrainStep
        ok = 1
        gstep = None
        try:
            dev_batch = next(iter(PrefetchLoader([batch], dev)))
            gstep = GraphedTrainStep(lambda bt: fwd_bwd(model, bt, largs), opt, dev_batch, scheduler=sched, warmup=1)
            for bt in PrefetchLoader([batch] * a.warmup, dev):
                gstep(bt)
        except Exception as exc:   # noqa: BLE001 - reported in the line, the eager figure stands
            ok = 0
            import traceback
            traceback.print_exc(file=sys.stderr)
            graph_info = {"captured": False, "error": f"{type(exc).__name__}: {exc}"[:300]}
        if world > 1:
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag.item())
        if ok:
            loader = PrefetchLoader([batch] * a.steps, dev)
            barrier()
            e0.record()
            glosses = []
            for bt in loader:
                glosses.append(gstep(bt).clone())
            e1.record()
            barrier()
            gms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([gms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                gms = float(t.item())
            ms_step = gms / a.steps
            graph_info = {"captured": True, "ms_per_step": ms_step, "eager_ms_per_step": eager_ms_step,
                          "loss_first_last": [glosses[0].item(), glosses[-1].item()]}
        if gstep is not None:
            gstep.release()
    peak_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(peak_file)).get("bf16_tflops_sustained", 1400.0) if os.path.exists(peak_file) else 1400.0
    if rank == 0:
        lin = prof["linear_tcgen05"]
        tot = sum(v["ms"] for v in prof.values())
        # per layer count scaled to --layers; head + img_linear terms are small and included at their 12-layer share
        flops_step = 3.0 * b * (5.48e9 + 6.45e9) * a.layers / 12.0
        line = {
            "workload": f"train_itm.py step: {b} pairs per GPU x {world} GPU(s) = global batch {B}, L={a.seq_len}, R={a.regions}, "
                        f"{a.layers} layers, {'fp16' if a.fp16 else 'bf16'}; symmetric in-batch NLL over the global batch, "
                        "gradient average, clip 2.0, AdamW",
            "n_gpus": world, "ms_per_step": ms_step, "pairs_per_s": B / (ms_step * 1e-3),
            "model_tflops_per_gpu": flops_step / (ms_step * 1e-3) / 1e12,
            "model_frac_of_bf16_sustained_peak": flops_step / (ms_step * 1e-3) / 1e12 / peak,
            "linear_tcgen05": {"ms_per_step": lin["ms"] / a.steps, "tflops": lin["flops"] / (lin["ms"] * 1e-3) / 1e12 if lin["ms"] else 0.0,
                               "frac_of_peak": (lin["flops"] / (lin["ms"] * 1e-3) / 1e12 / peak) if lin["ms"] else 0.0,
                               "launches_per_step": lin["launches"] / a.steps},
            "cuda_graph": graph_info,
            "kernel_ms_per_step": {k: round(v["ms"] / a.steps, 3) for k, v in prof.items() if v["launches"]},
            "kernel_time_share_of_step": tot / a.steps / ms_step,
            "kernel_times_from": "the eager run of the same step (per-launch CUDA events cannot bracket graph nodes)",
            "gpu_launches_per_step": sum(v["launches"] for v in prof.values()) / a.steps,
            "loss_first_last": [losses[0].item(), losses[-1].item()],
            "peak_tflops": peak, "distributed_check": check,
        }
        return line
    return None


def main():
    # stdout carries exactly one JSON line (NCCL prints its version banner there): everything else goes to stderr
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-gpu-batch", type=int, default=512)
    ap.add_argument("--seq-len", type=int, default=32)
    ap.add_argument("--regions", type=int, default=36)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--fp16", action="store_true")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="time the eager step only")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    line = measure(a, rank, world, local_rank, dev)
    if line is not None:
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
