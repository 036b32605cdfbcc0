"""Host-side mirror of uniter_model/data/loader.py's PrefetchLoader (eval_itm.py:16, train_itm.py): wraps a dataloader
and moves batch i + 1 to the GPU on a side stream while batch i is being computed (uniter_model/data/loader.py:80-160).

The nested itm_fast_collate batch (dict of dicts / lists / tensors / None, dvl/data/itm.py:203-288) is walked recursively;
tensors are copied with non_blocking=True (pinned sources overlap with compute) and `record_stream`-ed on the consumer's
stream when handed out, so the caching allocator does not recycle them while kernels of the main stream still read them -
the kernels of this package launch on torch's current stream, which is what makes this composition valid.
"""
import torch


def move_to_cuda(batch, device=None):
    if isinstance(batch, torch.Tensor):
        return batch.cuda(device, non_blocking=True)
    if isinstance(batch, list):
        return [move_to_cuda(t, device) for t in batch]
    if isinstance(batch, tuple):
        return tuple(move_to_cuda(t, device) for t in batch)
    if isinstance(batch, dict):
        return {n: move_to_cuda(t, device) for n, t in batch.items()}
    return batch


def record_cuda_stream(batch):
    if isinstance(batch, torch.Tensor):
        batch.record_stream(torch.cuda.current_stream())
    elif isinstance(batch, (list, tuple)):
        for t in batch:
            record_cuda_stream(t)
    elif isinstance(batch, dict):
        for t in batch.values():
            record_cuda_stream(t)


class PrefetchLoader(object):
    """Same interface and hand-out order as uniter_model/data/loader.py:80-160: iterate, len(), attribute pass-through."""

    def __init__(self, loader, device=None):
        self.loader = loader
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.batch = None

    def __iter__(self):
        it = iter(self.loader)
        self.preload(it)
        batch = self.next(it)
        while batch is not None:
            yield batch
            batch = self.next(it)

    def __len__(self):
        return len(self.loader)

    def preload(self, it):
        try:
            self.batch = next(it)
        except StopIteration:
            self.batch = None
            return
        with torch.cuda.stream(self.stream):
            self.batch = move_to_cuda(self.batch, self.device)

    def next(self, it):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)
        batch = self.batch
        if batch is not None:
            record_cuda_stream(batch)
        self.preload(it)
        return batch

    def __getattr__(self, name):
        return getattr(self.loader, name)
