"""Stand-in for the slice of `apex.amp` that train_itm.py touches when `args.fp16` is set (train_itm.py:252-258;
dvl/models/bi_encoder.py:589-601 calls amp.initialize, which setup_for_distributed_mode replaces):

    with amp.scale_loss(loss, optimizer) as scaled_loss:
        scaled_loss.backward()
    torch.nn.utils.clip_grad_norm_(amp.master_params(optimizer), args.max_grad_norm)

apex keeps fp32 master weights and scales the loss so that fp16 gradients do not underflow.  Here the master weights are
the module's own fp32 parameters (FusedAdamW's flat buffers) and parameter gradients are always accumulated in fp32, but
the ACTIVATION gradients of an fp16 tower are fp16 - the loss scale protects those.  Dynamic scaling follows apex: start
at 2^16, halve and skip the step when a gradient is non-finite, double after 2000 clean steps.  With bf16 towers (the
default) the scale stays 1 and nothing is skipped.  Alias it for unmodified scripts:

    import sys, types, lightningdot_b200.amp as amp
    apex = types.ModuleType("apex"); apex.amp = amp; sys.modules["apex"] = apex; sys.modules["apex.amp"] = amp
"""
import contextlib

import torch

INIT_SCALE = 2.0 ** 16
GROWTH_INTERVAL = 2000


class LossScaler(object):
    def __init__(self, enabled=True):
        self.enabled = enabled
        self.scale = INIT_SCALE if enabled else 1.0
        self.clean_steps = 0
        self.skipped = 0

    def update(self, overflow):
        if not self.enabled:
            return
        if overflow:
            self.scale = max(self.scale / 2.0, 1.0)
            self.clean_steps = 0
            self.skipped += 1
        else:
            self.clean_steps += 1
            if self.clean_steps % GROWTH_INTERVAL == 0:
                self.scale *= 2.0


def _fp16_in_use(optimizer):
    return getattr(optimizer, "shadow_dtype", None) == torch.float16


def _scaler(optimizer):
    sc = getattr(optimizer, "_amp_scaler", None)
    if sc is None:
        sc = LossScaler(enabled=_fp16_in_use(optimizer))
        optimizer._amp_scaler = sc
    return sc


def initialize(models, optimizers=None, opt_level="O1", **kwargs):
    """apex.amp.initialize: nothing to patch here - the towers' compute dtype is set by setup_for_distributed_mode."""
    if optimizers is None:
        return models
    return models, optimizers


def master_params(optimizer):
    """The fp32 parameters the optimiser steps (what apex calls master params)."""
    for group in optimizer.param_groups:
        for p in group["params"]:
            yield p


def _grads(optimizer):
    flat = getattr(optimizer, "_flat", None)
    if flat:
        views = [f["g"] for f in flat if f is not None]
        inside = set()
        for f in flat:
            if f is not None:
                inside.update(id(p) for p in f["params"] if p.grad is not None and
                              f["g"].data_ptr() <= p.grad.data_ptr() < f["g"].data_ptr() + 4 * f["g"].numel())
        rest = [p.grad for p in master_params(optimizer) if p.grad is not None and id(p) not in inside]
        return views + rest
    return [p.grad for p in master_params(optimizer) if p.grad is not None]


@contextlib.contextmanager
def scale_loss(loss, optimizer, **kwargs):
    """Yields loss * scale; on exit the gradients THIS backward produced are unscaled (so clip_grad_norm_ and the
    optimiser see true gradients) and checked: a non-finite gradient makes the next optimizer.step() a no-op and halves
    the scale.  Gradients accumulated by earlier micro-batches (gradient_accumulation_steps > 1: train_itm.py calls this
    every micro-step and zero_grad only after step()) were unscaled when they were produced: they are set aside on entry
    and added back after the unscale, as apex does with its stashed gradients."""
    sc = _scaler(optimizer)
    if not sc.enabled:
        yield loss
        return
    stash = [(g, g.clone()) for g in _grads(optimizer)]
    for g, _ in stash:
        g.zero_()
    yield loss * sc.scale
    grads = _grads(optimizer)
    if not grads:
        return
    torch._foreach_mul_(grads, 1.0 / sc.scale)
    total = torch.stack([g.abs().max() for g in grads]).max()
    overflow = ~torch.isfinite(total)
    if getattr(optimizer, "distributed", False):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            flag = overflow.to(torch.int32)        # every rank must skip (or take) the step together: step() is collective
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            overflow = flag > 0
    overflow = bool(overflow.item())               # (one host sync per step, as apex's dynamic scaler does)
    sc.update(overflow)
    if overflow:
        for g in grads:
            g.zero_()
        _skip_next_step(optimizer)
        return
    live = {g.data_ptr(): g for g in grads}
    for g, saved in stash:
        tgt = live.get(g.data_ptr())
        if tgt is not None and tgt.shape == saved.shape:
            tgt.add_(saved)


def _skip_next_step(optimizer):
    if getattr(optimizer, "_amp_real_step", None) is not None:
        return
    real = optimizer.step

    def skipped_step(*a, **k):
        optimizer.step = real
        optimizer._amp_real_step = None
        return None
    optimizer._amp_real_step = real
    optimizer.step = skipped_step
