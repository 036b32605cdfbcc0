"""Training-step kernels one by one against plain PyTorch fp32 references of the same op:
the tcgen05 GEMM in its backward forms (MN-major operands, split-K accumulation, GELU-gradient epilogue), LayerNorm /
attention / GELU backward, bias column sums, embedding backward pieces, NLL backward, AdamW."""
import math

import numpy as np
import pytest
import torch

from lightningdot_b200 import _lib

pytestmark = pytest.mark.gpu

DT = {0: torch.float16, 1: torch.bfloat16}


def gen(seed=0):
    return torch.Generator(device="cuda").manual_seed(seed)


# ------------------------------------------------------------------------------------------------------- GEMM forms
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 768, 768), (1000, 768, 3072), (77, 1536, 768), (130, 200, 72),
                                   (40000, 768, 2304), (33000, 3072, 768)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_dgrad_form(cuda_lib, M, N, K, fmt):
    """dX[M, N = in] = dY[M, K = out] . W[K = out, N = in]: B is the nn.Linear weight as stored (MN-major)."""
    g = gen(1)
    dy = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(DT[fmt])
    w = (torch.randn(K, N, device="cuda", generator=g) * 0.05).to(DT[fmt])
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    _lib.check(cuda_lib.ldot_gemm(_lib.ptr(dy), K, 0, _lib.ptr(w), N, 1, None, None, 0, _lib.ptr(out), N, M, N, K, fmt, 0, 1,
                                  0, _lib.stream_ptr()))
    ref = dy.float() @ w.float()
    torch.testing.assert_close(out, ref, atol=2e-4, rtol=1e-5)


def test_dgrad_epilogues(cuda_lib):
    """epi 3 (+ aux: the residual branch of the gradient) and epi 2 (* GELU'(aux)), 16-bit outputs."""
    M, N, K, fmt = 700, 3072, 768, 1
    g = gen(2)
    dy = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(DT[fmt])
    w = (torch.randn(K, N, device="cuda", generator=g) * 0.05).to(DT[fmt])
    aux = torch.randn(M, N, device="cuda", generator=g).to(DT[fmt])
    out = torch.empty((M, N), device="cuda", dtype=DT[fmt])
    _lib.check(cuda_lib.ldot_gemm(_lib.ptr(dy), K, 0, _lib.ptr(w), N, 1, None, _lib.ptr(aux), N, _lib.ptr(out), N, M, N, K,
                                  fmt, 3, 0, 0, _lib.stream_ptr()))
    ref = dy.float() @ w.float() + aux.float()
    torch.testing.assert_close(out.float(), ref, atol=4e-2, rtol=8e-3)
    _lib.check(cuda_lib.ldot_gemm(_lib.ptr(dy), K, 0, _lib.ptr(w), N, 1, None, _lib.ptr(aux), N, _lib.ptr(out), N, M, N, K,
                                  fmt, 2, 0, 0, _lib.stream_ptr()))
    x = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    ref = (dy.float() @ w.float()) * x.grad
    torch.testing.assert_close(out.float(), ref, atol=4e-2, rtol=8e-3)
    # epi 4: * aux (the GELU'(z) stored by the training forward)
    _lib.check(cuda_lib.ldot_gemm(_lib.ptr(dy), K, 0, _lib.ptr(w), N, 1, None, _lib.ptr(aux), N, _lib.ptr(out), N, M, N, K,
                                  fmt, 4, 0, 0, _lib.stream_ptr()))
    torch.testing.assert_close(out.float(), (dy.float() @ w.float()) * aux.float(), atol=4e-2, rtol=8e-3)


@pytest.mark.parametrize("T,O,I", [(64, 256, 128), (1000, 768, 768), (5000, 3072, 768), (4100, 768, 3072), (37, 1536, 768),
                                   (4, 768, 1536), (20000, 2304, 768), (900, 768, 2048)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_wgrad_form_accumulates(cuda_lib, T, O, I, fmt):
    """dW[O, I] += dY[T, O]^T . X[T, I]: both operands MN-major, fp32 accumulation into the gradient buffer with split-K."""
    g = gen(3)
    dy = (torch.randn(T, O, device="cuda", generator=g) * 0.1).to(DT[fmt])
    x = (torch.randn(T, I, device="cuda", generator=g) * 0.5).to(DT[fmt])
    init = torch.randn(O, I, device="cuda", generator=g)
    out = init.clone()
    for _ in range(2):   # two accumulation passes (gradient accumulation across micro-batches)
        _lib.check(cuda_lib.ldot_gemm(_lib.ptr(dy), O, 1, _lib.ptr(x), I, 1, None, None, 0, _lib.ptr(out), I, O, I, T, fmt, 0,
                                      1, 1, _lib.stream_ptr()))
    ref = init + 2.0 * (dy.float().t() @ x.float())
    torch.testing.assert_close(out, ref, atol=2e-3 * math.sqrt(T / 1000 + 1), rtol=1e-4)


def test_wgrad_strided_operands(cuda_lib):
    """The [CLS] rows of a [B, S, H] activation as the X operand (row pitch S * H), as the projection-head wgrad uses."""
    B, S, H, O = 12, 5, 768, 1536
    g = gen(4)
    h = (torch.randn(B * S, H, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
    dy = (torch.randn(B, O, device="cuda", generator=g) * 0.1).to(torch.bfloat16)
    out = torch.zeros(O, H, device="cuda")
    _lib.check(cuda_lib.ldot_gemm(_lib.ptr(dy), O, 1, _lib.ptr(h), S * H, 1, None, None, 0, _lib.ptr(out), H, O, H, B, 1, 0, 1,
                                  1, _lib.stream_ptr()))
    ref = dy.float().t() @ h.view(B, S, H)[:, 0].float()
    torch.testing.assert_close(out, ref, atol=1e-3, rtol=1e-4)


def test_gemm_rejects_unsupported_forms(cuda_lib):
    a = torch.zeros(64, 64, device="cuda", dtype=torch.bfloat16)
    o = torch.zeros(64, 64, device="cuda")
    rc = cuda_lib.ldot_gemm(_lib.ptr(a), 64, 1, _lib.ptr(a), 64, 0, None, None, 0, _lib.ptr(o), 64, 64, 64, 64, 1, 0, 1, 0,
                            _lib.stream_ptr())
    assert rc == -1 and b"unsupported" in cuda_lib.ldot_last_error()


# --------------------------------------------------------------------------------------------------- LayerNorm bwd
@pytest.mark.parametrize("rows,H", [(1, 768), (37, 768), (5000, 768), (9, 1536), (300, 256)])
@pytest.mark.parametrize("fmt", [0, 1])
@pytest.mark.parametrize("f32", [False, True])
def test_layernorm_bwd(cuda_lib, rows, H, fmt, f32):
    g = gen(5)
    dt = torch.float32 if f32 else DT[fmt]
    x = (torch.randn(rows, H, device="cuda", generator=g) * 1.5 + 0.3).to(dt)
    dy = (torch.randn(rows, H, device="cuda", generator=g) * 0.2).to(dt)
    gamma = torch.rand(H, device="cuda", generator=g) + 0.5
    beta = torch.randn(H, device="cuda", generator=g)
    dx = torch.empty((rows, H), device="cuda", dtype=dt)
    dgamma = torch.ones(H, device="cuda")
    dbeta = torch.ones(H, device="cuda")
    dxsum = torch.ones(H, device="cuda")
    _lib.check(cuda_lib.ldot_layernorm_bwd(_lib.ptr(dy), H, int(f32), _lib.ptr(x), H, int(f32), _lib.ptr(gamma), _lib.ptr(dx),
                                           H, int(f32), _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.ptr(dxsum), rows, H, fmt,
                                           _lib.stream_ptr()))
    xr = x.double().requires_grad_(True)
    gr = gamma.double().requires_grad_(True)
    br = beta.double().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xr, (H,), gr, br, eps=1e-12)
    y.backward(dy.double())
    tol = dict(atol=1e-4, rtol=1e-4) if f32 else (dict(atol=3e-2, rtol=1.6e-2) if fmt == 1 else dict(atol=4e-3, rtol=2e-3))
    torch.testing.assert_close(dx.double(), xr.grad, **tol)
    red = dict(atol=1e-3 * math.sqrt(rows), rtol=1e-4)
    torch.testing.assert_close(dgamma.double() - 1, gr.grad, **red)
    torch.testing.assert_close(dbeta.double() - 1, br.grad, **red)
    torch.testing.assert_close(dxsum.double() - 1, xr.grad.sum(0), **red)


# ----------------------------------------------------------------------------------------------------- attention bwd
def attention_ref(qkv, mask, B, S, H, heads):
    q, k, v = qkv.view(B, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0 + (1.0 - mask.to(qkv.dtype))[:, None, None, :] * -10000.0
    p = torch.softmax(s, dim=-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B * S, H)


@pytest.mark.parametrize("B,S", [(3, 32), (2, 37), (1, 128), (5, 7), (2, 62)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_attention_bwd(cuda_lib, B, S, fmt):
    H, heads = 768, 12
    g = gen(6)
    qkv = (torch.randn(B * S, 3 * H, device="cuda", generator=g) * 1.2).to(DT[fmt])
    dctx = (torch.randn(B * S, H, device="cuda", generator=g) * 0.3).to(DT[fmt])
    lens = torch.randint(max(1, S // 2), S + 1, (B,), device="cuda", generator=g)
    mask = (torch.arange(S, device="cuda")[None, :] < lens[:, None]).to(torch.int64)
    ctx = torch.empty((B * S, H), device="cuda", dtype=DT[fmt])
    _lib.check(cuda_lib.ldot_attention(_lib.ptr(qkv), _lib.ptr(mask), _lib.ptr(ctx), B, S, H, heads, S, fmt, _lib.stream_ptr()))
    dqkv = torch.full((B * S, 3 * H), float("nan"), device="cuda", dtype=DT[fmt])
    _lib.check(cuda_lib.ldot_attention_bwd(_lib.ptr(qkv), _lib.ptr(mask), _lib.ptr(ctx), _lib.ptr(dctx), _lib.ptr(dqkv), B, S,
                                           H, heads, 0.0, 0, 0, fmt, _lib.stream_ptr()))
    x = qkv.double().requires_grad_(True)
    attention_ref(x, mask, B, S, H, heads).backward(dctx.double())
    tol = dict(atol=2e-2, rtol=3e-2) if fmt == 1 else dict(atol=3e-3, rtol=4e-3)
    torch.testing.assert_close(dqkv.double(), x.grad, **tol)


# ----------------------------------------------------------------------------------------------------------- dropout
@pytest.mark.parametrize("rows,cols,p", [(37, 768, 0.1), (1000, 3072, 0.25), (3, 8, 0.5)])
def test_dropout_kernel_matches_oracle_mask(cuda_lib, rows, cols, p):
    """ldot_dropout against the CPU restatement of the mask function (oracle/dropout.py): the SAME elements are dropped,
    survivors are scaled by 1 / (1 - p), the residual is added after; the keep rate is 1 - p."""
    from oracle import dropout as odrop
    g = gen(21)
    seed, site = 0x1234_5678_9ABC, 9
    x = torch.randn(rows, cols, device="cuda", generator=g).to(torch.bfloat16)
    res = torch.randn(rows, cols, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.empty_like(x)
    _lib.check(cuda_lib.ldot_dropout(_lib.ptr(x), _lib.ptr(res), _lib.ptr(out), rows, cols, cols, p, seed, site, 1,
                                     _lib.stream_ptr()))
    keep = odrop.keep_mask((rows, cols), p, seed, site)
    want = (x.float().cpu() * keep / (1.0 - float(np.float32(p))) + res.float().cpu()).to(torch.bfloat16)
    # (atol: x / (1 - p) + res can cancel to ~0, where one bf16 rounding of the product is the whole result)
    torch.testing.assert_close(out.cpu().float(), want.float(), atol=2 ** -8, rtol=2 ** -7)
    if rows * cols > 10000:
        assert abs(keep.float().mean().item() - (1 - p)) < 4 * (p * (1 - p) / (rows * cols)) ** 0.5 + 1e-3
    # masks of different sites / seeds are different functions
    assert not torch.equal(keep, odrop.keep_mask((rows, cols), p, seed, site + 1)) or rows * cols < 64
    # in place, no residual (the gradient-masking call of the backward)
    y = x.clone()
    _lib.check(cuda_lib.ldot_dropout(_lib.ptr(y), None, _lib.ptr(y), rows, cols, cols, p, seed, site, 1, _lib.stream_ptr()))
    assert (y.cpu() == 0).eq(~keep | (x.cpu() == 0)).all()


def attention_ref_drop(qkv, mask, B, S, H, heads, keep, p):
    q, k, v = qkv.view(B, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0 + (1.0 - mask.to(qkv.dtype))[:, None, None, :] * -10000.0
    pr = torch.softmax(s, dim=-1) * keep.to(qkv.dtype) / (1.0 - p)
    return (pr @ v).permute(0, 2, 1, 3).reshape(B * S, H)


@pytest.mark.parametrize("B,S", [(3, 32), (2, 37), (1, 128)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_attention_dropout_fwd_bwd(cuda_lib, B, S, fmt):
    """ldot_attention_train / ldot_attention_bwd with attention-probability dropout against torch autograd (fp64) over the
    same mask (uniter_model/model/layer.py:93)."""
    from oracle import dropout as odrop
    H, heads, p, seed, site = 768, 12, 0.1, 77, 4
    g = gen(6)
    qkv = (torch.randn(B * S, 3 * H, device="cuda", generator=g) * 1.2).to(DT[fmt])
    dctx = (torch.randn(B * S, H, device="cuda", generator=g) * 0.3).to(DT[fmt])
    lens = torch.randint(max(1, S // 2), S + 1, (B,), device="cuda", generator=g)
    mask = (torch.arange(S, device="cuda")[None, :] < lens[:, None]).to(torch.int64)
    ctx = torch.empty((B * S, H), device="cuda", dtype=DT[fmt])
    _lib.check(cuda_lib.ldot_attention_train(_lib.ptr(qkv), _lib.ptr(mask), _lib.ptr(ctx), B, S, H, heads, p, seed, site, fmt,
                                             _lib.stream_ptr()))
    dqkv = torch.full((B * S, 3 * H), float("nan"), device="cuda", dtype=DT[fmt])
    _lib.check(cuda_lib.ldot_attention_bwd(_lib.ptr(qkv), _lib.ptr(mask), _lib.ptr(ctx), _lib.ptr(dctx), _lib.ptr(dqkv), B, S,
                                           H, heads, p, seed, site, fmt, _lib.stream_ptr()))
    keep = odrop.keep_mask((B, heads, S, S), p, seed, site).cuda()
    x = qkv.double().requires_grad_(True)
    ref = attention_ref_drop(x, mask, B, S, H, heads, keep, float(np.float32(p)))
    ref.backward(dctx.double())
    tol = dict(atol=2e-2, rtol=3e-2) if fmt == 1 else dict(atol=3e-3, rtol=4e-3)
    torch.testing.assert_close(ctx.double(), ref.detach(), **tol)
    torch.testing.assert_close(dqkv.double(), x.grad, **tol)


# ------------------------------------------------------------------------------------------- fused training forms
@pytest.mark.parametrize("M,N,K", [(37, 768, 768), (1000, 768, 3072), (20000, 768, 768), (300, 264, 72)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_linear_dropout_equals_linear_then_dropout(cuda_lib, M, N, K, fmt):
    """ldot_linear_dropout = dropout(A W^T + b) + residual with the mask of ldot_dropout at the same (seed, site): the
    SAME elements are dropped as the CPU restatement says, kept values agree with the fp32 reference."""
    from oracle import dropout as odrop
    g = gen(21)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(DT[fmt])
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(DT[fmt])
    b = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g).to(DT[fmt])
    out = torch.empty((M, N), device="cuda", dtype=DT[fmt])
    p, seed, site = 0.1, 0x1234ABCD5678, 9
    _lib.check(cuda_lib.ldot_linear_dropout(_lib.ptr(a), K, _lib.ptr(w), K, _lib.ptr(b), _lib.ptr(res), N, _lib.ptr(out), N,
                                            M, N, K, fmt, p, seed, site, _lib.stream_ptr()))
    keep = odrop.keep_mask((M, N), p, seed, site).cuda()
    dense = a.float() @ w.float().T + b
    ref = dense * keep / (1.0 - float(np.float32(p))) + res.float()
    tol = dict(atol=4e-2, rtol=1.6e-2) if fmt == 1 else dict(atol=6e-3, rtol=2e-3)
    torch.testing.assert_close(out.float(), ref, **tol)
    # dropped positions carry the residual alone, exactly
    assert torch.equal(out[~keep], res[~keep])
    assert 0.85 < keep.float().mean().item() < 0.95 or M * N < 20000


@pytest.mark.parametrize("M,N,K", [(37, 3072, 768), (20000, 3072, 768), (130, 200, 72)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_linear_gelu_grad_writes_both(cuda_lib, M, N, K, fmt):
    """ldot_linear_gelu_grad: GELU(z) and GELU'(z) of the 16-bit-rounded pre-activation z from one kernel.  GELU(z) is
    bit-identical to ldot_linear (act 0) followed by ldot_gelu on its 16-bit output (what amp does in the reference);
    GELU'(z) against torch autograd on that same rounded z."""
    g = gen(22)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(DT[fmt])
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.1).to(DT[fmt])
    b = torch.randn(N, device="cuda", generator=g)
    gp = torch.zeros((M, N), device="cuda", dtype=DT[fmt])
    out = torch.zeros((M, N), device="cuda", dtype=DT[fmt])
    _lib.check(cuda_lib.ldot_linear_gelu_grad(_lib.ptr(a), K, _lib.ptr(w), K, _lib.ptr(b), _lib.ptr(gp), N, _lib.ptr(out), N,
                                              M, N, K, fmt, _lib.stream_ptr()))
    z = torch.zeros_like(out)
    out2 = torch.zeros_like(out)
    _lib.check(cuda_lib.ldot_linear(_lib.ptr(a), K, _lib.ptr(w), K, _lib.ptr(b), None, 0, _lib.ptr(z), N, M, N, K, fmt,
                                    0, 0, _lib.stream_ptr()))
    _lib.check(cuda_lib.ldot_gelu(_lib.ptr(z), _lib.ptr(out2), z.numel(), fmt, _lib.stream_ptr()))
    assert torch.equal(out, out2)
    zr = z.double().requires_grad_(True)
    torch.nn.functional.gelu(zr).sum().backward()
    tol = dict(atol=8e-3, rtol=8e-3) if fmt == 1 else dict(atol=1e-3, rtol=1e-3)
    torch.testing.assert_close(gp.double(), zr.grad, **tol)
    ref = torch.nn.functional.gelu(a.float() @ w.float().T + b)
    tol = dict(atol=4e-2, rtol=1.6e-2) if fmt == 1 else dict(atol=6e-3, rtol=2e-3)
    torch.testing.assert_close(out.float(), ref, **tol)


@pytest.mark.parametrize("fmt", [0, 1])
def test_gelu_grad_in_place(cuda_lib, fmt):
    """ldot_gelu_grad: out = GELU(z) (bit-identical to ldot_gelu) and the z buffer holds GELU'(z) afterwards."""
    g = gen(25)
    z = (torch.randn(777, 3072, device="cuda", generator=g) * 2).to(DT[fmt])
    want_g = torch.empty_like(z)
    _lib.check(cuda_lib.ldot_gelu(_lib.ptr(z), _lib.ptr(want_g), z.numel(), fmt, _lib.stream_ptr()))
    zr = z.double().requires_grad_(True)
    torch.nn.functional.gelu(zr).sum().backward()
    buf, out = z.clone(), torch.empty_like(z)
    _lib.check(cuda_lib.ldot_gelu_grad(_lib.ptr(buf), _lib.ptr(out), z.numel(), fmt, _lib.stream_ptr()))
    assert torch.equal(out, want_g)
    tol = dict(atol=8e-3, rtol=8e-3) if fmt == 1 else dict(atol=1e-3, rtol=1e-3)
    torch.testing.assert_close(buf.double(), zr.grad, **tol)


def test_gelu_grad_saturates_cleanly(cuda_lib):
    """GELU' at extreme pre-activations (the polynomial argument is clamped): exactly 0 / 1, no NaN."""
    x = torch.tensor([-3e38, -65504.0, -100.0, -21.0, 21.0, 100.0, 65504.0, 3e38], device="cuda").to(torch.bfloat16)
    dy = torch.ones_like(x)
    dx = torch.empty_like(x)
    _lib.check(cuda_lib.ldot_gelu_bwd(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(dx), x.numel(), 1, _lib.stream_ptr()))
    assert dx.float().tolist() == [0.0, 0.0, 0.0, 0.0, 1.0, 1.0, 1.0, 1.0]


@pytest.mark.parametrize("rows,H", [(1, 768), (37, 768), (20000, 768), (300, 256)])
@pytest.mark.parametrize("fmt", [0, 1])
def test_layernorm_bwd_dropout(cuda_lib, rows, H, fmt):
    """The fused form against LayerNorm backward (fp64 autograd) followed by the oracle's mask: dx, the masked dx, and the
    column sums of the masked dx."""
    from oracle import dropout as odrop
    g = gen(23)
    dt = DT[fmt]
    x = (torch.randn(rows, H, device="cuda", generator=g) * 1.5 + 0.3).to(dt)
    dy = (torch.randn(rows, H, device="cuda", generator=g) * 0.2).to(dt)
    gamma = torch.rand(H, device="cuda", generator=g) + 0.5
    dx = torch.empty((rows, H), device="cuda", dtype=dt)
    dxm = torch.empty((rows, H), device="cuda", dtype=dt)
    dgamma = torch.ones(H, device="cuda")
    dbeta = torch.ones(H, device="cuda")
    dxsum = torch.ones(H, device="cuda")
    p, seed, site = 0.1, 0xBEEF0000CAFE, 6
    _lib.check(cuda_lib.ldot_layernorm_bwd_dropout(_lib.ptr(dy), H, _lib.ptr(x), H, _lib.ptr(gamma), _lib.ptr(dx), _lib.ptr(dxm),
                                                   H, _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.ptr(dxsum), rows, H, p, seed,
                                                   site, fmt, _lib.stream_ptr()))
    xr = x.double().requires_grad_(True)
    gr = gamma.double().requires_grad_(True)
    y = torch.nn.functional.layer_norm(xr, (H,), gr, torch.zeros(H, device="cuda", dtype=torch.float64), eps=1e-12)
    y.backward(dy.double())
    keep = odrop.keep_mask((rows, H), p, seed, site).cuda()
    ref_m = xr.grad * keep / (1.0 - float(np.float32(p)))
    tol = dict(atol=3e-2, rtol=1.6e-2) if fmt == 1 else dict(atol=4e-3, rtol=2e-3)
    torch.testing.assert_close(dx.double(), xr.grad, **tol)
    torch.testing.assert_close(dxm.double(), ref_m, **tol)
    assert torch.all(dxm[~keep] == 0)
    red = dict(atol=1e-3 * math.sqrt(rows), rtol=1e-4)
    torch.testing.assert_close(dgamma.double() - 1, gr.grad, **red)
    torch.testing.assert_close(dbeta.double() - 1, dy.double().sum(0), **red)
    torch.testing.assert_close(dxsum.double() - 1, ref_m.sum(0), **red)


def test_dropout_epoch_changes_masks_on_the_device(cuda_lib):
    """ldot_dropout_epoch: the mask depends on a device word read at run time (what a replayed CUDA graph needs): same
    word -> same mask, bumped word -> another mask, cleared -> the host-keyed mask again."""
    rows, cols, p, seed, site = 64, 768, 0.3, 77, 3
    x = torch.ones((rows, cols), device="cuda", dtype=torch.bfloat16)

    def run():
        out = torch.empty_like(x)
        _lib.check(cuda_lib.ldot_dropout(_lib.ptr(x), None, _lib.ptr(out), rows, cols, cols, p, seed, site, 1, _lib.stream_ptr()))
        return out

    base = run()
    epoch = torch.zeros(1, device="cuda", dtype=torch.int32)
    try:
        _lib.check(cuda_lib.ldot_dropout_epoch(_lib.ptr(epoch)))
        e0, e0b = run(), run()
        epoch.add_(1)
        e1 = run()
    finally:
        _lib.check(cuda_lib.ldot_dropout_epoch(None))
    again = run()
    assert torch.equal(e0, e0b) and torch.equal(base, again)
    assert not torch.equal(e0, e1) and not torch.equal(e0, base)
    for t in (e0, e1):
        assert abs((t == 0).float().mean().item() - p) < 0.02


def test_adamw_dev_matches_adamw(cuda_lib):
    """ldot_adamw_dev (step scalars from device memory) == ldot_adamw with the same scalars by value."""
    n, lr, b1, b2, eps, wd, step = 10000, 3e-4, 0.9, 0.999, 1e-8, 0.01, 7
    g = gen(24)
    p0 = torch.randn(n, device="cuda", generator=g)
    gr = torch.randn(n, device="cuda", generator=g)
    m0 = torch.randn(n, device="cuda", generator=g) * 0.1
    v0 = torch.rand(n, device="cuda", generator=g) * 0.01
    outs = []
    for dev in (False, True):
        p_, m_, v_ = p0.clone(), m0.clone(), v0.clone()
        p16 = torch.empty(n, device="cuda", dtype=torch.bfloat16)
        if dev:
            hyper = torch.tensor([lr, 1.0 - b1 ** step, math.sqrt(1.0 - b2 ** step)], dtype=torch.float32, device="cuda")
            _lib.check(cuda_lib.ldot_adamw_dev(_lib.ptr(p_), _lib.ptr(gr), _lib.ptr(m_), _lib.ptr(v_), _lib.ptr(p16), n,
                                               _lib.ptr(hyper), b1, b2, eps, wd, None, 0.0, 1, _lib.stream_ptr()))
        else:
            _lib.check(cuda_lib.ldot_adamw(_lib.ptr(p_), _lib.ptr(gr), _lib.ptr(m_), _lib.ptr(v_), _lib.ptr(p16), n, lr, b1, b2,
                                           eps, wd, step, None, 0.0, 1, _lib.stream_ptr()))
        outs.append((p_, m_, v_, p16))
    for a, b in zip(*outs):
        torch.testing.assert_close(a.float(), b.float(), atol=1e-7, rtol=1e-6)


# ------------------------------------------------------------------------------------------------ elementwise / sums
@pytest.mark.parametrize("fmt", [0, 1])
def test_gelu_and_gelu_bwd(cuda_lib, fmt):
    g = gen(7)
    x = (torch.randn(1000, 3072, device="cuda", generator=g) * 2).to(DT[fmt])
    dy = torch.randn(1000, 3072, device="cuda", generator=g).to(DT[fmt])
    y = torch.empty_like(x)
    dx = torch.empty_like(x)
    _lib.check(cuda_lib.ldot_gelu(_lib.ptr(x), _lib.ptr(y), x.numel(), fmt, _lib.stream_ptr()))
    _lib.check(cuda_lib.ldot_gelu_bwd(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(dx), x.numel(), fmt, _lib.stream_ptr()))
    xr = x.double().requires_grad_(True)
    yr = torch.nn.functional.gelu(xr)
    yr.backward(dy.double())
    tol = dict(atol=1e-2, rtol=8e-3) if fmt == 1 else dict(atol=1e-3, rtol=1e-3)
    torch.testing.assert_close(y.double(), yr.detach(), **tol)
    torch.testing.assert_close(dx.double(), xr.grad, **tol)


@pytest.mark.parametrize("rows,N", [(1, 768), (4097, 2304), (300, 3072), (10, 8)])
def test_colsum16(cuda_lib, rows, N):
    g = gen(8)
    x = torch.randn(rows, N, device="cuda", generator=g).to(torch.bfloat16)
    out = torch.full((N,), 2.0, device="cuda")
    _lib.check(cuda_lib.ldot_colsum16(_lib.ptr(x), N, rows, N, _lib.ptr(out), 1, _lib.stream_ptr()))
    torch.testing.assert_close(out, 2.0 + x.float().sum(0), atol=1e-3 * math.sqrt(rows), rtol=1e-5)


def test_text_embedding_backward_pieces(cuda_lib):
    B, L, H, V, P = 5, 9, 768, 1000, 64
    g = gen(9)
    ids = torch.randint(0, V, (B, L), device="cuda", generator=g)
    ids[:, -2:] = 0   # padding
    pos = torch.arange(L, device="cuda")[None, :].contiguous()
    word = (torch.randn(V, H, device="cuda", generator=g) * 0.02).to(torch.bfloat16)
    post = (torch.randn(P, H, device="cuda", generator=g) * 0.02).to(torch.bfloat16)
    type0 = (torch.randn(H, device="cuda", generator=g) * 0.02).to(torch.bfloat16)
    s = torch.empty((B * L, H), device="cuda")
    _lib.check(cuda_lib.ldot_embed_text_sum(_lib.ptr(ids), _lib.ptr(pos), 0, _lib.ptr(word), _lib.ptr(post), _lib.ptr(type0),
                                            _lib.ptr(s), B, L, H, V, P, 1, _lib.stream_ptr()))
    ref = word.float()[ids.view(-1)] + post.float()[pos.expand(B, L).reshape(-1)] + type0.float()
    torch.testing.assert_close(s, ref, atol=1e-6, rtol=1e-6)
    dx = torch.randn(B * L, H, device="cuda", generator=g)
    dword = torch.zeros(V, H, device="cuda")
    dpos = torch.zeros(P, H, device="cuda")
    _lib.check(cuda_lib.ldot_embed_scatter(_lib.ptr(dx), _lib.ptr(ids), _lib.ptr(pos), 0, _lib.ptr(dword), _lib.ptr(dpos), B, L,
                                           H, V, P, _lib.stream_ptr()))
    rw = torch.zeros(V, H, device="cuda").index_add_(0, ids.view(-1), dx)
    rw[0] = 0   # padding_idx
    rp = torch.zeros(P, H, device="cuda").index_add_(0, pos.expand(B, L).reshape(-1), dx)
    torch.testing.assert_close(dword, rw, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(dpos, rp, atol=1e-5, rtol=1e-5)


def test_image_embedding_backward_pieces(cuda_lib):
    rows, H = 77, 768
    g = gen(10)
    lin = torch.randn(rows, H, device="cuda", generator=g)
    box = torch.rand(rows, 7, device="cuda", generator=g)
    pw = torch.randn(H, 7, device="cuda", generator=g) * 0.2
    pb = torch.randn(H, device="cuda", generator=g) * 0.1
    gs = [torch.rand(H, device="cuda", generator=g) + 0.5 for _ in range(2)]
    bs = [torch.randn(H, device="cuda", generator=g) * 0.1 for _ in range(2)]
    type1 = torch.randn(H, device="cuda", generator=g) * 0.02
    q = torch.empty((rows, H), device="cuda")
    spre = torch.empty((rows, H), device="cuda")
    _lib.check(cuda_lib.ldot_embed_image_pre(_lib.ptr(lin), _lib.ptr(box), _lib.ptr(gs[0]), _lib.ptr(bs[0]), _lib.ptr(pw),
                                             _lib.ptr(pb), _lib.ptr(gs[1]), _lib.ptr(bs[1]), _lib.ptr(type1), _lib.ptr(q),
                                             _lib.ptr(spre), rows, H, _lib.stream_ptr()))
    ln = torch.nn.functional.layer_norm
    qr = box @ pw.t() + pb
    sr = ln(lin, (H,), gs[0], bs[0], 1e-12) + ln(qr, (H,), gs[1], bs[1], 1e-12) + type1
    torch.testing.assert_close(q, qr, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(spre, sr, atol=1e-4, rtol=1e-5)
    dq = torch.randn(rows, H, device="cuda", generator=g)
    dw = torch.ones(H, 7, device="cuda")
    _lib.check(cuda_lib.ldot_pos_wgrad(_lib.ptr(dq), _lib.ptr(box), rows, H, _lib.ptr(dw), _lib.stream_ptr()))
    torch.testing.assert_close(dw, 1.0 + dq.t() @ box, atol=1e-4, rtol=1e-5)


@pytest.mark.parametrize("bq,bc,reduction", [(7, 13, 0), (96, 96, 0), (300, 1000, 1)])
def test_nll_bwd(cuda_lib, bq, bc, reduction):
    g = gen(11)
    s = torch.randn(bq, bc, device="cuda", generator=g) * 3
    pos = torch.randint(0, bc, (bq,), device="cuda", generator=g)
    up = torch.tensor([0.7], device="cuda")
    ld = (bc + 7) // 8 * 8
    ds = torch.full((bq, ld), float("nan"), device="cuda", dtype=torch.bfloat16)
    _lib.check(cuda_lib.ldot_inbatch_nll_bwd(_lib.ptr(s), _lib.ptr(pos), bq, bc, _lib.ptr(up), reduction, _lib.ptr(ds), ld, 1,
                                             _lib.stream_ptr()))
    sr = s.double().requires_grad_(True)
    loss = torch.nn.functional.nll_loss(torch.log_softmax(sr, 1), pos, reduction="mean" if reduction == 0 else "sum")
    (loss * 0.7).backward()
    torch.testing.assert_close(ds[:, :bc].double(), sr.grad, atol=1e-5, rtol=8e-3)
    assert (ds[:, bc:] == 0).all()


def test_adamw_matches_torch(cuda_lib):
    g = gen(12)
    n = 100003
    p0 = torch.randn(n, device="cuda", generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=3e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    p16 = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    for step in range(1, 4):
        grad = torch.randn(n, device="cuda", generator=g) * 3
        ref.grad = grad.clone()
        total = torch.nn.utils.clip_grad_norm_([ref], 2.0)
        opt.step()
        ss = torch.zeros(1, device="cuda")
        _lib.check(cuda_lib.ldot_sumsq(_lib.ptr(grad), n, _lib.ptr(ss), _lib.stream_ptr()))
        torch.testing.assert_close(ss.sqrt()[0], total, rtol=1e-5, atol=0)
        _lib.check(cuda_lib.ldot_adamw(_lib.ptr(p), _lib.ptr(grad), _lib.ptr(m), _lib.ptr(v), _lib.ptr(p16), n, 3e-3, 0.9, 0.999,
                                       1e-8, 0.01, step, _lib.ptr(ss), 2.0, 1, _lib.stream_ptr()))
        torch.testing.assert_close(p, ref.data, atol=1e-6, rtol=1e-5)
    torch.testing.assert_close(p16.float(), p, atol=0, rtol=8e-3)
