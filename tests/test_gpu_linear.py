"""tcgen05 Linear (ldot_linear) against a plain PyTorch fp32 reference of the same op."""
import pytest
import torch

from lightningdot_b200 import _lib

pytestmark = pytest.mark.gpu


def run_linear(M, N, K, fmt, act=0, bias=True, res=False, out_f32=True, seed=0):
    lib = _lib.load()
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(dt)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(dt)
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    r = torch.randn(M, N, device="cuda", generator=g).to(dt) if res else None
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if out_f32 else dt)
    _lib.check(lib.ldot_linear(_lib.ptr(a), K, _lib.ptr(w), K, _lib.ptr(b), _lib.ptr(r), N, _lib.ptr(out), N, M, N, K,
                               fmt, act, int(out_f32), _lib.stream_ptr()))
    ref = a.float() @ w.float().t()
    if bias:
        ref = ref + b
    if act:
        ref = torch.nn.functional.gelu(ref)  # erf form
    if res:
        ref = ref + r.float()
    return out.float(), ref


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 768, 768), (1000, 3072, 768), (4096, 768, 3072),
                                   (77, 1536, 768), (130, 200, 72), (1, 768, 1536)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_linear_fp32_out(cuda_lib, M, N, K, fmt):
    out, ref = run_linear(M, N, K, fmt)
    # inputs are exactly representable 16-bit values, accumulation is fp32: only summation-order noise remains
    torch.testing.assert_close(out, ref, atol=1e-4, rtol=1e-5)


def test_linear_gelu_16bit_out(cuda_lib):
    out, ref = run_linear(512, 3072, 768, 1, act=1, out_f32=False)
    torch.testing.assert_close(out, ref, atol=2e-2, rtol=8e-3)   # bf16 output rounding: 2^-8 relative
    out, ref = run_linear(512, 3072, 768, 0, act=1, out_f32=False)
    torch.testing.assert_close(out, ref, atol=2e-3, rtol=1e-3)   # fp16 output: 2^-11 relative


def test_linear_residual(cuda_lib):
    out, ref = run_linear(512, 768, 3072, 1, res=True, out_f32=True)
    torch.testing.assert_close(out, ref, atol=2e-4, rtol=1e-5)
    out, ref = run_linear(640, 768, 768, 1, res=True, out_f32=False, bias=False)
    torch.testing.assert_close(out, ref, atol=4e-2, rtol=8e-3)


def test_linear_tails_and_pitches(cuda_lib):
    """N not a multiple of the 32-column store box, M not a multiple of 32, output / residual pitches wider than N
    (the TMA store clips at the tensor edges; nothing outside [M, N] may be touched)."""
    lib = _lib.load()
    M, N, K, ld = 77, 200, 72, 264
    g = torch.Generator(device="cuda").manual_seed(3)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, ld, device="cuda", generator=g).to(torch.bfloat16)
    for out_f32, dt in ((1, torch.float32), (0, torch.bfloat16)):
        out = torch.full((M + 3, ld), 7.0, device="cuda", dtype=dt)
        _lib.check(lib.ldot_linear(_lib.ptr(a), K, _lib.ptr(w), K, _lib.ptr(b), _lib.ptr(r), ld, _lib.ptr(out), ld, M, N, K,
                                   1, 0, out_f32, _lib.stream_ptr()))
        ref = a.float() @ w.float().t() + b + r[:, :N].float()
        tol = dict(atol=1e-4, rtol=1e-5) if out_f32 else dict(atol=4e-2, rtol=8e-3)
        torch.testing.assert_close(out[:M, :N].float(), ref, **tol)
        assert (out[M:] == 7.0).all() and (out[:, N:] == 7.0).all()


def test_gelu_epilogue_is_the_erf_form(cuda_lib):
    """The fused activation against the exact erf-GELU of uniter_model/model/layer.py:31-37 in fp64: identity weights
    route a dense sweep of inputs straight to the epilogue.  Tolerance 5e-6 absolute (the sigmoid-of-polynomial
    evaluation is within 3.4e-6; tanh-GELU would miss this bar by two orders of magnitude)."""
    import math
    lib = _lib.load()
    K = 64
    x = torch.linspace(-12.0, 12.0, 4096 * K, device="cuda").to(torch.float16).view(-1, K)
    w = torch.eye(K, device="cuda", dtype=torch.float16)
    out = torch.empty((x.shape[0], K), device="cuda", dtype=torch.float32)
    _lib.check(lib.ldot_linear(_lib.ptr(x), K, _lib.ptr(w), K, None, None, 0, _lib.ptr(out), K, x.shape[0], K, K,
                               0, 1, 1, _lib.stream_ptr()))
    xd = x.double()
    ref = 0.5 * xd * (1.0 + torch.erf(xd / math.sqrt(2.0)))
    assert (out.double() - ref).abs().max().item() <= 5e-6
    tanh_form = 0.5 * xd * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (xd + 0.044715 * xd ** 3)))
    assert (tanh_form - ref).abs().max().item() > 1e-4   # (what the bar excludes)


# (M >= 16 384 with N = 768 and a residual runs the CTA-pair form - 2 x 3 clusters, linear_ln2.cuh; 16 384 + 131 rows leave a
# partial 256-row block and a partial 128-row half; the no-residual second call of the test stays on the single-CTA form)
@pytest.mark.parametrize("M,N,K", [(128, 768, 768), (300, 768, 3072), (4097, 768, 768), (77, 256, 64), (1000, 512, 128),
                                   (130, 96, 72), (16384 + 131, 768, 768), (16384 + 257, 768, 3072), (45 * 256, 768, 768)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_linear_layernorm_fused(cuda_lib, M, N, K, fmt):
    """ldot_linear_ln = LayerNorm(A W^T + bias + residual) * gamma + beta against fp32 torch.  The statistics are exact
    fp32 sums of the fp32 pre-LayerNorm values, so the only error is the final 16-bit rounding (2^-9 / 2^-12 relative)
    plus fp32 summation-order noise."""
    lib = _lib.load()
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(dt)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(dt)
    b = torch.randn(N, device="cuda", generator=g)
    r = (torch.randn(M, N, device="cuda", generator=g) + 0.3).to(dt)
    gamma = 1.0 + 0.1 * torch.randn(N, device="cuda", generator=g)
    beta = 0.1 * torch.randn(N, device="cuda", generator=g)
    out = torch.full((M + 2, N), 7.0, device="cuda", dtype=dt)
    _lib.check(lib.ldot_linear_ln(_lib.ptr(a), K, _lib.ptr(w), K, _lib.ptr(b), _lib.ptr(r), N, _lib.ptr(gamma), _lib.ptr(beta),
                                  _lib.ptr(out), N, M, N, K, fmt, _lib.stream_ptr()))
    x = a.float() @ w.float().t() + b + r.float()
    ref = torch.nn.functional.layer_norm(x, (N,), gamma, beta, 1e-12)
    tol = dict(atol=4e-3, rtol=2e-3) if fmt == 0 else dict(atol=3e-2, rtol=8e-3)
    torch.testing.assert_close(out[:M].float(), ref, **tol)
    assert (out[M:] == 7.0).all()
    # no bias / no residual
    _lib.check(lib.ldot_linear_ln(_lib.ptr(a), K, _lib.ptr(w), K, None, None, 0, _lib.ptr(gamma), _lib.ptr(beta),
                                  _lib.ptr(out), N, M, N, K, fmt, _lib.stream_ptr()))
    ref2 = torch.nn.functional.layer_norm(a.float() @ w.float().t(), (N,), gamma, beta, 1e-12)
    torch.testing.assert_close(out[:M].float(), ref2, **tol)


@pytest.mark.parametrize("B,S", [(8, 32), (9, 32), (1, 32), (7, 37), (64, 37), (5, 17), (130, 1), (3, 64), (4, 65), (2, 128),
                                 (11, 50), (6, 100)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_fused_qkv_attention_equals_projection_then_attention(cuda_lib, B, S, fmt):
    """ldot_qkv_attention (BertSelfAttention as one kernel, Q | K | V kept on chip) against the two-kernel path it replaces -
    ldot_linear (N = 3H) + ldot_attention - on the same inputs: both round Q, K, V to 16 bit after the bias and run the same
    attention arithmetic, so the context tensors must agree to the last bit; and against a torch fp32 reference of
    uniter_model/model/layer.py:60-101.  Shapes cover whole / partial last tiles, sequences that do not divide 128, key
    blocks that run past row 127 of the tile, S = 1 and S = 128; masks are ragged."""
    lib = _lib.load()
    H, heads = 768, 12
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(100 * B + S)
    x = (torch.randn(B * S, H, device="cuda", generator=g) * 0.8).to(dt)
    w = (torch.randn(3 * H, H, device="cuda", generator=g) * 0.04).to(dt)
    b = torch.randn(3 * H, device="cuda", generator=g) * 0.1
    lens = torch.randint(1, S + 1, (B,), device="cuda", generator=g)
    mask = (torch.arange(S, device="cuda")[None, :] < lens[:, None]).long()
    st = _lib.stream_ptr()
    qkv = torch.empty((B * S, 3 * H), dtype=dt, device="cuda")
    want = torch.empty((B * S, H), dtype=dt, device="cuda")
    _lib.check(lib.ldot_linear(_lib.ptr(x), H, _lib.ptr(w), H, _lib.ptr(b), None, 0, _lib.ptr(qkv), 3 * H, B * S, 3 * H, H,
                               fmt, 0, 0, st))
    _lib.check(lib.ldot_attention(_lib.ptr(qkv), _lib.ptr(mask), _lib.ptr(want), B, S, H, heads, S, fmt, st))
    got = torch.full((B * S, H), float("nan"), dtype=dt, device="cuda")
    _lib.check(lib.ldot_qkv_attention(_lib.ptr(x), H, _lib.ptr(w), H, _lib.ptr(b), _lib.ptr(mask), _lib.ptr(got), B, S, H,
                                      heads, H, fmt, st))
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all()
    assert torch.equal(got, want), float((got.float() - want.float()).abs().max())
    q, k, v = (qkv.float().view(B, S, 3, heads, 64)[:, :, j].permute(0, 2, 1, 3) for j in range(3))
    sc = q @ k.transpose(-1, -2) / 8.0 + (1.0 - mask.float())[:, None, None, :] * -10000.0
    ref = (torch.softmax(sc, dim=-1) @ v).permute(0, 2, 1, 3).reshape(B * S, H)
    torch.testing.assert_close(got.float(), ref, atol=3e-2 if fmt else 4e-3, rtol=2e-2 if fmt else 3e-3)
