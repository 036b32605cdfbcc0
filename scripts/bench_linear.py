"""Per-shape timing of the encoder GEMM kernels next to cuBLAS (torch.matmul) on the same box, back to back for
long enough that both run at the sustained (power-capped) clock.  Not a bench line: a development probe that tells
how far each of our tcgen05 kernels is from what the library reaches on the same shape under the same conditions.

    python scripts/bench_linear.py [tokens] [seconds_per_case] [case,case...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lightningdot_b200 import _lib  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 320000
SECS = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
ONLY = set(sys.argv[3].split(",")) if len(sys.argv) > 3 else None   # case filter (ncu captures of one shape)
lib = _lib.load()
dt = torch.bfloat16
fmt = 1


def timed(fn, flops):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    one = e0.elapsed_time(e1)
    reps = max(5, int(SECS * 1e3 / max(one, 1e-3)))
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, flops / ms / 1e9


def case(name, N, K, act=0, ln=False, res=False):
    if ONLY is not None and name not in ONLY:
        return
    g = torch.Generator(device="cuda").manual_seed(N + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(dt)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.02).to(dt)
    b = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g).to(dt) if (res or ln) else None
    gamma, beta = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=dt)
    s = _lib.stream_ptr()
    flops = 2.0 * M * N * K
    if ln:
        def ours():
            _lib.check(lib.ldot_linear_ln(_lib.ptr(a), K, _lib.ptr(w), K, _lib.ptr(b), _lib.ptr(r), N, _lib.ptr(gamma),
                                          _lib.ptr(beta), _lib.ptr(out), N, M, N, K, fmt, s))
    else:
        def ours():
            _lib.check(lib.ldot_linear(_lib.ptr(a), K, _lib.ptr(w), K, _lib.ptr(b), _lib.ptr(r), N, _lib.ptr(out), N,
                                       M, N, K, fmt, act, 0, s))
    wt = w.t()

    def cublas():
        torch.matmul(a, wt, out=out)

    ms_c, tf_c = timed(cublas, flops)
    ms_o, tf_o = timed(ours, flops)
    print(f"{name:10s} M={M} N={N:5d} K={K:5d}: ours {ms_o:8.3f} ms {tf_o:7.1f} TF/s | cuBLAS (no epilogue) {ms_c:8.3f} ms "
          f"{tf_c:7.1f} TF/s | ours/cuBLAS {tf_o / tf_c:.3f}", flush=True)


def case_qkv_attn(name, S=32, H=768, heads=12):
    """BertSelfAttention as one kernel (ldot_qkv_attention) next to the two kernels it replaces and to cuBLAS's bare
    projection GEMM; FLOPs counted: projection + the two attention contractions."""
    if ONLY is not None and name not in ONLY:
        return
    B = M // S
    T = B * S
    g = torch.Generator(device="cuda").manual_seed(S)
    x = (torch.randn(T, H, device="cuda", generator=g) * 0.5).to(dt)
    w = (torch.randn(3 * H, H, device="cuda", generator=g) * 0.02).to(dt)
    b = torch.randn(3 * H, device="cuda", generator=g) * 0.1
    mask = torch.ones(B, S, dtype=torch.int64, device="cuda")
    qkv = torch.empty(T, 3 * H, device="cuda", dtype=dt)
    ctx = torch.empty(T, H, device="cuda", dtype=dt)
    s = _lib.stream_ptr()
    flops = 2.0 * T * 3 * H * H + 4.0 * T * S * H

    def fused():
        _lib.check(lib.ldot_qkv_attention(_lib.ptr(x), H, _lib.ptr(w), H, _lib.ptr(b), _lib.ptr(mask), _lib.ptr(ctx), B, S, H,
                                          heads, H, fmt, s))

    def two():
        _lib.check(lib.ldot_linear(_lib.ptr(x), H, _lib.ptr(w), H, _lib.ptr(b), None, 0, _lib.ptr(qkv), 3 * H, T, 3 * H, H,
                                   fmt, 0, 0, s))
        _lib.check(lib.ldot_attention(_lib.ptr(qkv), _lib.ptr(mask), _lib.ptr(ctx), B, S, H, heads, S, fmt, s))
    wt = w.t()

    def cublas():
        torch.matmul(x, wt, out=qkv)
    ms_c, tf_c = timed(cublas, 2.0 * T * 3 * H * H)
    ms_t, tf_t = timed(two, flops)
    ms_f, tf_f = timed(fused, flops)
    print(f"{name:10s} T={T} S={S}: fused {ms_f:8.3f} ms {tf_f:7.1f} TF/s | projection + attention kernels {ms_t:8.3f} ms "
          f"{tf_t:7.1f} TF/s | cuBLAS projection only {ms_c:8.3f} ms {tf_c:7.1f} TF/s", flush=True)


print(torch.cuda.get_device_name(0), flush=True)
case_qkv_attn("qkv+attn", 32)
case_qkv_attn("qkv+attn37", 37)
case("qkv", 2304, 768)
case("o+ln", 768, 768, ln=True)
case("o+res", 768, 768, res=True)
case("ffn1+gelu", 3072, 768, act=1)
case("ffn1", 3072, 768)
case("ffn2+ln", 768, 3072, ln=True)
case("ffn2+res", 768, 3072, res=True)
