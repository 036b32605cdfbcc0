"""`horovod` - namesake shim: the reference scripts use Horovod for rank bookkeeping (eval_itm.py:76-83,
train_itm.py:72-76, uniter_model/data/data.py:36-41,186-187).  Here one process per GPU is launched by torchrun and the
collectives run on torch.distributed (NCCL over NVLink); horovod.torch answers from that."""
from . import torch  # noqa: F401

# (the scripts import `transformers.tokenization_bert` right after this package - train_itm.py:15-21: register the alias)
from lightningdot_b200 import compat as _compat  # noqa: E402

_compat.install()
