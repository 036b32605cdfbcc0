"""lightningdot_b200 - B200-native (sm_100a) implementation of LightningDOT's bi-encoder retrieval hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); every hot operation is a hand-written
CUDA kernel reached through the C ABI of libldot_sm100a.so (include/ldot.h).  See DESIGN.md.
"""
__version__ = "0.1.0"
