"""ctypes binding of libldot_sm100a.so (C ABI in include/ldot.h).

There is no CPU or PyTorch fallback: if the shared library is missing, or the device is not sm_100, every entry
point raises.  Build the library with `python -c "import __graft_entry__ as g; g.build()"` (or
`python -m lightningdot_b200.build`).
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libldot_sm100a.so")

ABI_VERSION = 6   # LDOT_ABI_VERSION of include/ldot.h
COARSE_FP16 = 0
COARSE_BF16 = 1

_lib = None
# bumped by in-place parameter updates that bypass torch's version counters (FusedAdamW): part of the towers' cache key
param_generation = [0]
# flat buffers of the live FusedAdamW optimisers: dicts with "p" (fp32 master), "p16" (16-bit mirror the AdamW kernel
# keeps up to date, or None) and "g" (fp32 gradients).  Towers alias their weights to them (towers.py: load).
flat_buffers = []


def shadow_view(t, dtype):
    """The 16-bit mirror of the fp32 parameter tensor `t` inside a FusedAdamW flat buffer (same shape), or None."""
    if t.device.type != "cuda" or not t.is_contiguous() or t.dtype.itemsize != 4:
        return None
    ptr, n = t.data_ptr(), t.numel()
    for f in flat_buffers:
        p16 = f.get("p16")
        if p16 is None or p16.dtype != dtype or p16.device != t.device:
            continue
        base = f["p"].data_ptr()
        if base <= ptr and ptr + 4 * n <= base + 4 * f["p"].numel():
            off = (ptr - base) // 4
            return p16[off:off + n].view(t.shape)
    return None

# name -> (restype, argtypes): every symbol include/ldot.h declares
SIGNATURES = {
    "ldot_abi_version": (c_int32, []),
    "ldot_last_error": (c_char_p, []),
    "ldot_device_check": (c_int32, []),
    "ldot_prof_num_classes": (c_int32, []),
    "ldot_prof_class_name": (c_char_p, [c_int32]),
    "ldot_prof_enable": (c_int32, [c_int32]),
    "ldot_prof_reset": (c_int32, []),
    "ldot_prof_read": (c_int32, [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ldot_index_prepare_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "ldot_index_prepare": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_size_t, c_void_p]),
    "ldot_flatip_search_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32, c_int32, c_int32]),
    "ldot_flatip_search": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                     c_int32, c_int32, c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_size_t, c_void_p]),
    "ldot_flatip_search_phase": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                           c_int32, c_int32, c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_size_t, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "ldot_flatip_exact_workspace_bytes": (c_size_t, [c_int64]),
    "ldot_flatip_exact": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_int64, c_void_p,
                                    c_void_p, c_void_p, c_size_t, c_void_p]),
    "ldot_topk_merge": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_int64, c_int64, c_void_p, c_void_p,
                                  c_void_p]),
    "ldot_linear": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                              c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "ldot_linear_ln": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                 c_void_p, c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p]),
    "ldot_layernorm": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int32,
                                 c_int32, c_void_p]),
    "ldot_embed_text": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "ldot_embed_image": (c_int32, [c_void_p] * 12 + [c_int32] * 6 + [c_void_p]),
    "ldot_attention": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "ldot_cast_f32": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "ldot_qkv_attention": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                     c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "ldot_split16": (c_int32, [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]),
    "ldot_inbatch_nll": (c_int32, [c_void_p, c_void_p, c_float, c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    # training step
    "ldot_gemm": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int64, c_void_p,
                            c_int64, c_int64, c_int32, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "ldot_layernorm_bwd": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int64,
                                     c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p]),
    "ldot_attention_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                     c_float, ctypes.c_uint64, c_int32, c_int32, c_void_p]),
    "ldot_attention_train": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_float,
                                       ctypes.c_uint64, c_int32, c_int32, c_void_p]),
    "ldot_dropout": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_float, ctypes.c_uint64, c_int32,
                               c_int32, c_void_p]),
    "ldot_gelu": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "ldot_gelu_bwd": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "ldot_colsum16": (c_int32, [c_void_p, c_int64, c_int64, c_int32, c_void_p, c_int32, c_void_p]),
    "ldot_embed_text_sum": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                      c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "ldot_embed_scatter": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                     c_int32, c_int32, c_void_p]),
    "ldot_embed_image_pre": (c_int32, [c_void_p] * 11 + [c_int64, c_int32, c_void_p]),
    "ldot_pos_wgrad": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "ldot_inbatch_nll_bwd": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int32, c_void_p, c_int64, c_int32,
                                       c_void_p]),
    "ldot_sumsq": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p]),
    "ldot_adamw": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                             c_float, c_int32, c_void_p, c_float, c_int32, c_void_p]),
    # fused training forms + CUDA-graph support (ABI 6)
    "ldot_linear_dropout": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                      c_int64, c_int32, c_int32, c_int32, c_float, ctypes.c_uint64, c_int32, c_void_p]),
    "ldot_linear_gelu_grad": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                       c_int64, c_int32, c_int32, c_int32, c_void_p]),
    "ldot_gelu_grad": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "ldot_layernorm_bwd_dropout": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64,
                                             c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_float, ctypes.c_uint64,
                                             c_int32, c_int32, c_void_p]),
    "ldot_dropout_epoch": (c_int32, [c_void_p]),
    "ldot_adamw_dev": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float, c_float,
                                 c_float, c_float, c_void_p, c_float, c_int32, c_void_p]),
}


class LdotError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises LdotError when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LdotError(
            f"{LIB_PATH} not found: the CUDA extension is not built and there is no fallback path. "
            "Run `python -m lightningdot_b200.build`.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.ldot_abi_version() != ABI_VERSION:
        raise LdotError(f"ABI version mismatch: library reports {lib.ldot_abi_version()}, binding expects {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().ldot_last_error()
        raise LdotError(f"libldot_sm100a error {rc}: {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def prof_enable(on=True):
    """Start / stop bracketing every kernel launch of the library with CUDA events (measurement support)."""
    check(load().ldot_prof_enable(1 if on else 0))


def prof_reset():
    check(load().ldot_prof_reset())


def prof_read():
    """-> {kernel class: dict(launches, timed, ms, flops, bytes)} accumulated since the last prof_reset()."""
    lib = load()
    n = lib.ldot_prof_num_classes()
    i64, f64 = ctypes.c_int64 * n, ctypes.c_double * n
    launches, timed, ms, flops, nbytes = i64(), i64(), f64(), f64(), f64()
    check(lib.ldot_prof_read(n, launches, timed, ms, flops, nbytes))
    return {lib.ldot_prof_class_name(c).decode(): dict(launches=launches[c], timed=timed[c], ms=ms[c], flops=flops[c],
                                                       bytes=nbytes[c]) for c in range(n)}
