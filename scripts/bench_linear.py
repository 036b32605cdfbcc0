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


print(torch.cuda.get_device_name(0), flush=True)
case("qkv", 2304, 768)
case("o+ln", 768, 768, ln=True)
case("o+res", 768, 768, res=True)
case("ffn1+gelu", 3072, 768, act=1)
case("ffn1", 3072, 768)
case("ffn2+ln", 768, 3072, ln=True)
case("ffn2+res", 768, 3072, res=True)
