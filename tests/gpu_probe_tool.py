"""Development probe run under gpurun: checks the tcgen05 GEMM and the search pipeline step by step and prints
diagnostics (not a test, not a benchmark).  Usage: python tests/gpu_probe_tool.py [gemm] [search] [time]
(Lives under tests/ because it runs the oracle as its checker; pytest does not collect it.)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root
from lightningdot_b200 import _lib, synth  # noqa: E402
from lightningdot_b200.indexer import FlatIPIndex  # noqa: E402
from oracle import flatip  # noqa: E402  (probe only: the oracle is the checker)


def gemm_case(M, N, K, fmt, act=0, bias=True, res=False, out_f32=True):
    lib = _lib.load()
    dt = torch.float16 if fmt == 0 else torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(dt)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).to(dt)
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    r = (torch.randn(M, N, device="cuda", generator=g)).to(dt) if res else None
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if out_f32 else dt)
    _lib.check(lib.ldot_linear(_lib.ptr(a), K, _lib.ptr(w), K, _lib.ptr(b), _lib.ptr(r), N, _lib.ptr(out), N, M, N, K,
                               fmt, act, int(out_f32), _lib.stream_ptr()))
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    if bias:
        ref = ref + b
    if act:
        ref = torch.nn.functional.gelu(ref)
    if res:
        ref = ref + r.float()
    err = (out.float() - ref).abs()
    nan = int(torch.isnan(out.float()).sum())
    print(f"gemm M={M} N={N} K={K} fmt={fmt} act={act} bias={bias} res={res} f32={out_f32}: max_abs_err={err.nan_to_num(9e9).max().item():.3e} "
          f"ref_absmax={ref.abs().max().item():.3f} nan={nan}", flush=True)
    if nan or err.max().item() > (2e-3 if out_f32 else 5e-2) * max(1.0, ref.abs().max().item()):
        bad = torch.nonzero(torch.isnan(out.float()) | (err > 1e-2))
        print("   first bad entries:", bad[:8].tolist(), " bad rows:", torch.unique(bad[:, 0])[:16].tolist(),
              " bad cols:", torch.unique(bad[:, 1])[:16].tolist(), flush=True)
        print("   out[0,:8]", out[0, :8].tolist(), "\n   ref[0,:8]", ref[0, :8].tolist(), flush=True)
        return False
    return True


def probe_gemm():
    ok = True
    for (M, N, K) in [(128, 256, 64), (128, 256, 768), (256, 512, 768), (300, 768, 768), (1000, 3072, 768),
                      (4096, 768, 3072), (77, 1536, 768), (130, 200, 72)]:
        for fmt in (1, 0):
            ok &= gemm_case(M, N, K, fmt)
    ok &= gemm_case(512, 3072, 768, 1, act=1, out_f32=False)
    ok &= gemm_case(512, 768, 3072, 1, res=True, out_f32=False)
    ok &= gemm_case(512, 768, 768, 1, res=True, out_f32=True, bias=False)
    print("GEMM PROBE", "OK" if ok else "FAILED", flush=True)
    return ok


def search_case(n, nq, k, kind="gauss", dtype="fp16", center=True, d=768, sigma=3.0, coarse_k=0):
    x = synth.gaussian_index(n, d, seed=1) if kind == "gauss" else synth.collinear_index(n, d, seed=1)
    q, gt = synth.planted_queries(x, nq, sigma=sigma, seed=2)
    idx = FlatIPIndex(d, coarse_dtype=dtype, center=center, coarse_k=coarse_k)
    idx.add(x)
    qd = torch.from_numpy(q).cuda()
    t0 = time.time()
    s_raw, i_raw = idx.search_device(qd, k, resolve_flags=False)
    torch.cuda.synchronize()
    t1 = time.time()
    s, i = idx.search_device(qd, k)
    torch.cuda.synchronize()
    es, ei = idx.exact_search_device(qd[:min(nq, 64)].contiguous(), k)
    torch.cuda.synchronize()
    os_, oi = flatip.search(q[:min(nq, 512)], x, k)
    m = min(nq, 512)
    i_np, s_np = i.cpu().numpy()[:m], s.cpu().numpy()[:m]
    raw_np = i_raw.cpu().numpy()[:m]
    id_ok = (i_np == oi).all(axis=1)
    raw_ok = (raw_np == oi).all(axis=1)
    ex_ok = (ei.cpu().numpy() == oi[:min(nq, 64)]).all()
    ex_s_ok = np.array_equal(es.cpu().numpy(), os_[:min(nq, 64)])
    rel = np.abs(s_np - os_) / np.maximum(np.abs(os_), 1e-30)
    print(f"search n={n} nq={nq} k={k} {kind} {dtype} center={center}: flagged={idx.last_flagged}/{nq} "
          f"ids_exact={id_ok.mean():.4f} (pre-fallback {raw_ok.mean():.4f}) score_bits_equal={np.array_equal(s_np, os_)} "
          f"max_rel={rel.max():.2e} exact_path ids={bool(ex_ok)} scores_bits={ex_s_ok} first_call_s={t1 - t0:.3f}", flush=True)
    if not id_ok.all():
        r = int(np.nonzero(~id_ok)[0][0])
        print("   row", r, "got", i_np[r][:12], "want", oi[r][:12], "\n   got_s", s_np[r][:6], "want_s", os_[r][:6], flush=True)
    return bool(id_ok.all()) and bool(ex_ok)


def probe_search():
    ok = True
    ok &= search_case(1000, 64, 10)
    ok &= search_case(5000, 200, 100)
    ok &= search_case(5000, 200, 100, dtype="bf16")
    ok &= search_case(70000, 300, 100)
    ok &= search_case(70000, 300, 100, kind="collinear")
    ok &= search_case(70000, 300, 100, kind="collinear", dtype="bf16")
    ok &= search_case(70000, 300, 100, kind="collinear", center=False)
    ok &= search_case(50, 10, 100)
    ok &= search_case(300000, 1000, 100)
    ok &= search_case(20000, 100, 1000)
    print("SEARCH PROBE", "OK" if ok else "FAILED", flush=True)
    return ok


def probe_time():
    d, k = 768, 100
    for n, nq in [(1000000, 128), (1000000, 10000), (125000, 10000)]:
        x = torch.randn(n, d, device="cuda") / d ** 0.5
        q = torch.randn(nq, d, device="cuda") / d ** 0.5
        idx = FlatIPIndex(d)
        idx.add(x)
        for _ in range(2):
            idx.search_device(q, k)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(5):
            idx.search_device(q, k, resolve_flags=False)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 5
        tf = 2.0 * nq * n * d / ms / 1e9
        gb = n * d * 2 / ms / 1e6
        print(f"time n={n} nq={nq}: {ms:.3f} ms/search  {nq / ms * 1e3:.0f} q/s  {tf:.1f} TFLOP/s  index-read {gb:.0f} GB/s "
              f"flagged={idx.last_flagged}", flush=True)
        del idx, x, q
        torch.cuda.empty_cache()


if __name__ == "__main__":
    what = sys.argv[1:] or ["gemm", "search", "time"]
    print(torch.cuda.get_device_name(0), flush=True)
    if "gemm" in what:
        probe_gemm()
    if "search" in what:
        probe_search()
    if "time" in what:
        probe_time()


def probe_towers():
    from lightningdot_b200.towers import TowerEngine
    from oracle import towers as otowers
    for dtype in (torch.bfloat16, torch.float16):
        for kind, layers in (("txt", 2), ("img", 2), ("txt", 12), ("img", 12)):
            sd = synth.random_tower_state(kind, seed=42, perturb=True, layers=layers)
            eng = TowerEngine(kind, 768, 12, 3072, layers, dtype=dtype)
            eng.load(sd, "cuda")
            if kind == "txt":
                b = synth.text_batch(6, 32, seed=1, ragged=True)
                seq, pooled = eng.encode_text(b["input_ids"], b["attention_mask"], b["position_ids"], want_seq=True)
                with torch.no_grad():
                    oseq, opooled = otowers.text_tower(sd, b["input_ids"], b["attention_mask"], b["position_ids"])
            else:
                b = synth.image_batch(5, 36, seed=1, ragged=True)
                seq, pooled = eng.encode_image(b["input_ids"], b["attention_mask"], b["position_ids"], b["img_feat"],
                                               b["img_pos_feat"], b["gather_index"], want_seq=True)
                with torch.no_grad():
                    oseq, opooled = otowers.image_tower(sd, b["input_ids"], b["attention_mask"], b["position_ids"],
                                                        b["img_feat"], b["img_pos_feat"], b["gather_index"])
            torch.cuda.synchronize()
            got = pooled.float().cpu()
            cos = torch.nn.functional.cosine_similarity(got, opooled, dim=-1).min().item()
            err = (got - opooled).abs().max().item() / opooled.abs().max().item()
            herr = (seq[:, 0].float().cpu() - oseq[:, 0]).abs().max().item()
            print(f"tower {kind} L{layers} {dtype}: pooled cos_min={cos:.6f} rel_max_err={err:.3e} cls_hidden_abs_err={herr:.3e} "
                  f"nan={int(torch.isnan(got).sum())}", flush=True)
    # throughput: 10k captions, L=32
    sd = synth.random_tower_state("txt", seed=42, layers=12)
    eng = TowerEngine("txt", 768, 12, 3072, 12, dtype=torch.bfloat16)
    eng.load(sd, "cuda")
    b = synth.text_batch(10000, 32, seed=3)
    ids, mask, pos = b["input_ids"].cuda(), b["attention_mask"].cuda(), b["position_ids"].cuda()
    for _ in range(2):
        eng.encode_text(ids, mask, pos)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(3):
        eng.encode_text(ids, mask, pos)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 3
    print(f"text tower 10000 x 32: {ms:.2f} ms  {10000 / ms * 1e3:.0f} captions/s  {10000 * 5.48e9 / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__" and "towers" in sys.argv[1:]:
    probe_towers()
