"""TEST INFRASTRUCTURE (parity oracle) - CPU restatement of the product's counter-based dropout masks
(lightningdot_b200/csrc/dropout.cuh).  The reference draws dropout masks from torch's Philox stream
(nn.Dropout in uniter_model/model/layer.py:93,113,154 and model.py:245,272), which cannot be reproduced element for
element outside torch; parity under dropout is therefore checked with the product's own mask function restated here:
same keep probability, same 1 / (1 - p) scaling, and exact agreement of values and gradients for the same masks.

Sites (shared with lightningdot_b200/training.py): EMB after the embedding LayerNorm of the whole [B, S, H] input,
4 l + 0 attention probabilities of layer l, 4 l + 1 BertSelfOutput dense output, 4 l + 2 BertOutput dense output.
"""
import numpy as np
import torch

SITE_EMB = 0x7E0


def _mix32(x):
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7feb352d)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846ca68b)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x


def site_key(seed, site):
    lo, hi = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    inner = _mix32(np.array([(int(hi) + 0x9E3779B9 * site) & 0xFFFFFFFF], dtype=np.uint64))
    return _mix32(np.array([int(lo)], dtype=np.uint64) ^ inner)[0]


def keep_mask(shape, p, seed, site):
    """bool tensor of `shape`: element with flat (row-major) index i is kept iff mix32(i_lo ^ mix32(i_hi ^ key)) >= p 2^32."""
    n = int(np.prod(shape))
    idx = np.arange(n, dtype=np.uint64)
    key = site_key(seed, site)
    inner = _mix32((idx >> np.uint64(32)) ^ key)
    h = _mix32((idx & np.uint64(0xFFFFFFFF)) ^ inner)
    thr = min(int(float(np.float32(p)) * 4294967296.0), 4294967295)
    return torch.from_numpy((h >= np.uint64(thr)).reshape(shape))


class Dropper(object):
    """drop(site, x) hook for oracle.towers: x * mask / (1 - p) with the product's mask of that site."""

    def __init__(self, p_hidden, p_attn, seed):
        self.p_hidden, self.p_attn, self.seed = float(np.float32(p_hidden)), float(np.float32(p_attn)), int(seed)

    def __call__(self, site, x, attention=False):
        p = self.p_attn if attention else self.p_hidden
        if p <= 0:
            return x
        m = keep_mask(tuple(x.shape), p, self.seed, site)
        return x * m.to(x.dtype) / (1.0 - p)
