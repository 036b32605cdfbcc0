"""`apex` - namesake shim for the two apex.amp calls of train_itm.py's fp16 branch (train_itm.py:252-258:
amp.scale_loss, amp.master_params).  NVIDIA apex is not required: fp16 / bf16 selection is done by
setup_for_distributed_mode, dynamic loss scaling by lightningdot_b200.amp."""
from lightningdot_b200 import amp  # noqa: F401
