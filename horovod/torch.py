"""horovod.torch surface over torch.distributed.

  init()            joins the torchrun rendezvous (NCCL when CUDA is there, gloo otherwise); a plain `python script.py`
                    run is a world of one and touches nothing
  size / rank / local_rank / local_size / is_initialized / shutdown
  allreduce(_) / allgather / broadcast(_) / broadcast_parameters    thin torch.distributed wrappers for code that calls them
"""
import os

import torch as _torch
import torch.distributed as _dist


def _env(name, default):
    return int(os.environ.get(name, default))


def init(*_args, **_kwargs):
    if _env("WORLD_SIZE", 1) > 1 and not _dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if _torch.cuda.is_available():
            _torch.cuda.set_device(local_rank())
            _dist.init_process_group("nccl", device_id=_torch.device("cuda", local_rank()))
        else:
            _dist.init_process_group("gloo")


def shutdown():
    if _dist.is_initialized():
        _dist.destroy_process_group()


def is_initialized():
    return True


def size():
    return _dist.get_world_size() if _dist.is_initialized() else _env("WORLD_SIZE", 1)


def rank():
    return _dist.get_rank() if _dist.is_initialized() else _env("RANK", 0)


def local_rank():
    return _env("LOCAL_RANK", 0)


def local_size():
    return _env("LOCAL_WORLD_SIZE", size())


def _alone():
    return not _dist.is_initialized() or _dist.get_world_size() == 1


def allreduce_(tensor, average=True, name=None, op=None):
    if not _alone():
        _dist.all_reduce(tensor)
        if average and op is None:
            tensor.div_(size())
    return tensor


def allreduce(tensor, average=True, name=None, op=None):
    return allreduce_(tensor.clone(), average, name, op)


def allgather(tensor, name=None):
    """Concatenation along dim 0 of every rank's tensor (equal trailing shapes, like hvd.allgather on equal rows)."""
    if _alone():
        return tensor.clone()
    out = [_torch.empty_like(tensor) for _ in range(size())]
    _dist.all_gather(out, tensor.contiguous())
    return _torch.cat(out, 0)


def broadcast_(tensor, root_rank=0, name=None):
    if not _alone():
        _dist.broadcast(tensor, src=root_rank)
    return tensor


def broadcast(tensor, root_rank=0, name=None):
    return broadcast_(tensor.clone(), root_rank, name)


def broadcast_parameters(params, root_rank=0):
    items = params.items() if isinstance(params, dict) else params
    for _, p in items:
        if _torch.is_tensor(p):
            broadcast_(p.data if hasattr(p, "data") else p, root_rank)
