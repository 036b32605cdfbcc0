#!/usr/bin/env python
"""Headline benchmark: text queries/sec over a 1M-candidate index (BASELINE.json metric, configs[3]).

    python bench.py --gpus N --steps K --warmup W            # this framework (B200, N = 1, 2, 4, 8 via torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port) on host cores

One step = the whole hot path over one batch of synthetic input: 10 000 token-id queries (seq_len 32) ->
BERT-base text tower (random init, bf16 tcgen05 GEMMs) -> 768-d embeddings -> exact inner-product top-100 over a
1 000 000 x 768 index (row-sharded over the N GPUs, per-shard fused score+top-k, NCCL all-gather, merge).

  value   queries/s with token ids already resident in HBM, CUDA-event timed, max over ranks
  e2e     the same through the reference-facing calls (BertEncoder.forward + DenseFlatIndexer.search_knn) from
          PINNED HOST token ids to host (db_id list, score array) results, H2D/D2H inside the timed region
  roofline / roofline_search / roofline_online   per-kernel achieved rates measured live with CUDA events
          around every launch of the timed region (libldot's ldot_prof_* accounting), against MEASURED_PEAKS.json
  cpu_baseline   the oracle port (fp32 torch-CPU tower + fp32 sgemm flat-IP search) on a bounded sample, rank 0, N = 1

Only the cpu_baseline / --impl reference legs import oracle/ (as the thing timed beside the product, and as the
parity checker); the product legs fail loudly if libldot_sm100a.so is missing.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "text queries/sec over 1M-candidate index (encode + exact top-100)"
UNIT = "queries/s"
D = 768
GEN_BLOCK = 62500   # index rows per generator block (seeded by block number: identical data for every N)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--index-rows", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--seq-len", type=int, default=32)
    ap.add_argument("--cpu-sample", type=int, default=2048, help="queries of the cpu_baseline sample (~10 s of CPU work)")
    ap.add_argument("--ref-sample", type=int, default=256, help="queries per step of --impl reference (~1.5 s of CPU work per step)")
    ap.add_argument("--parity-sample", type=int, default=1024,
                    help="queries whose GPU top-k is compared with the oracle's exact CPU search over the full index")
    ap.add_argument("--train-steps", type=int, default=4, help="timed steps of the train_step leg (BASELINE configs[4]); 0 = skip")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            m = json.load(f)
        return dict(hbm=float(m["hbm_gbs"]), tc_burst=float(m["bf16_tflops"]),
                    tc_sustained=float(m.get("bf16_tflops_sustained", m["bf16_tflops"])), source="measured")
    # /opt/skills/guides/B200_PROFILING.md fallback
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback")


# --------------------------------------------------------------------------------------------------- clocks sampling
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], 0.0, set()
        for t, line in self.lines:
            if not (t0 <= t <= t1 + 0.1):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = max(smax, float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": smax or None, "reasons": sorted(reasons), "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- synthetic workload
def make_queries(nq, seq_len):
    from lightningdot_b200 import synth
    b = synth.text_batch(nq, seq_len, seed=3)
    return b["input_ids"], b["attention_mask"], b["position_ids"]


def index_block(block, rows, device):
    """Rows [block * GEN_BLOCK, ...) of the synthetic index: X ~ N(0, 1) / sqrt(d), seeded by the block number."""
    import torch
    g = torch.Generator(device=device).manual_seed(1000 + block)
    return torch.randn(rows, D, device=device, generator=g) / D ** 0.5


def build_shard(lo, hi, device):
    import torch
    x = torch.empty((hi - lo, D), dtype=torch.float32, device=device)
    b = lo // GEN_BLOCK
    while b * GEN_BLOCK < hi:
        r0, r1 = b * GEN_BLOCK, (b + 1) * GEN_BLOCK
        blk = index_block(b, GEN_BLOCK, device)
        s0, s1 = max(lo, r0), min(hi, r1)
        x[s0 - lo:s1 - lo] = blk[s0 - r0:s1 - r0]
        b += 1
    return x


def planted_rows(emb, n):
    """One planted index row per query: x = c * unit(e - mean e) so that the query's score against it sits around the
    maximum of its n gaussian competitors (Recall@1 well inside (0, 1): ranking errors would show)."""
    import torch
    nq = emb.shape[0]
    g = torch.Generator(device=emb.device).manual_seed(77)
    dirs = emb - emb.mean(0, keepdim=True)
    dirs = dirs / dirs.norm(dim=1, keepdim=True).clamp_min(1e-20)
    proj = (emb * dirs).sum(1).clamp_min(1e-6)
    target = emb.norm(dim=1) * (4.9 + 0.5 * torch.randn(nq, device=emb.device, generator=g)) / D ** 0.5
    # (rows of queries that sit almost on the mean would need a huge norm to reach the target score: capped at the
    # norm scale of the gaussian rows, so those queries simply score lower)
    rows = dirs * (target / proj).clamp_max(1.25)[:, None]
    gt = (torch.arange(nq, device=emb.device, dtype=torch.int64) * n) // nq + 7
    return rows, gt.clamp_max(n - 1)


def recall_at(ids, gt, ks=(1, 5, 10)):
    hit = ids == gt[:, None]
    return {str(k): float(hit[:, :k].any(dim=1).float().mean().item()) for k in ks}


# ----------------------------------------------------------------------------------------------------- reference arm
def cpu_text_tower(sd, ids, mask, pos, batch=64):
    import torch
    from oracle import towers as otowers
    out = []
    with torch.no_grad():
        for b in range(0, ids.shape[0], batch):
            _, pooled = otowers.text_tower(sd, ids[b:b + batch], mask[b:b + batch], pos)
            out.append(pooled)
    return torch.cat(out, 0)


def cpu_search_f32(q, x, k, q_block=512):
    """faiss IndexFlatIP.search restated for timing: fp32 sgemm over 512-query blocks (2 GB of scores at 1M rows) + top-k (oracle/flatip.py docstring)."""
    import torch
    s_out, i_out = [], []
    with torch.no_grad():
        for b in range(0, q.shape[0], q_block):
            s = q[b:b + q_block] @ x.t()
            v, i = torch.topk(s, min(k, x.shape[0]), dim=1)
            s_out.append(v)
            i_out.append(i)
    return torch.cat(s_out, 0), torch.cat(i_out, 0)


def run_reference(args):
    """The reference's CPU eval path for the same metric: fp32 BERT tower (oracle restatement of BertEncoder, pinned
    to the reference classes by tests/golden) + fp32 flat inner-product search, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from lightningdot_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n, k, L, s = args.index_rows, args.k, args.seq_len, args.ref_sample
    sd = synth.random_tower_state("txt", seed=42)
    g = torch.Generator().manual_seed(1000)
    x = torch.randn(n, D, generator=g) / D ** 0.5
    ids, mask, pos = make_queries(s * (args.steps + args.warmup), L)
    times = []
    for step in range(args.warmup + args.steps):
        sl = slice(step * s, (step + 1) * s)
        t0 = time.perf_counter()
        emb = cpu_text_tower(sd, ids[sl], mask[sl], pos)
        cpu_search_f32(emb, x, k)
        t1 = time.perf_counter()
        if step >= args.warmup:
            times.append(t1 - t0)
    total = sum(times)
    qps = s * len(times) / total
    sample = (f"{s} queries per step (seq_len {L}) through the fp32 CPU text tower + fp32 sgemm/top-{k} search over the "
              f"full {n} x {D} index; {len(times)} timed steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(json.dumps(line))


def workload_config(args, world):
    return {
        "workload": (f"BASELINE configs[3]: {args.queries} text queries (token ids, seq_len {args.seq_len}) -> BERT-base "
                     f"text tower -> exact inner-product top-{args.k} over a {args.index_rows} x {D} index"),
        "index_rows": args.index_rows, "queries": args.queries, "k": args.k, "seq_len": args.seq_len, "d": D,
        "parallelism": f"index row-sharded over {world} GPU(s), queries sharded for encoding, all-gather + merge",
        "l2": "inputs larger than L2: every step streams the 1.5 GB 16-bit index copy and 170 MB of tower weights",
    }


# ----------------------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from lightningdot_b200 import _lib, synth
    from lightningdot_b200.bi_encoder import BertEncoder, TowerConfig
    from lightningdot_b200.sharded import ShardedFlatIndexer, shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    _lib.check(lib.ldot_device_check())
    peaks = measured_peaks()
    n, nq, k, L = args.index_rows, args.queries, args.k, args.seq_len

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- model + inputs (reference initialiser, seed 42; generated on the CPU generator, shared with the oracle)
    sd = synth.random_tower_state("txt", seed=42)
    txt_model = BertEncoder(TowerConfig(vocab_size=synth.VOCAB), project_dim=D)
    txt_model.load_state_dict(sd, strict=True)
    txt_model.to(dev).eval()
    ids_h, mask_h, pos_h = make_queries(nq, L)
    ids_pin, mask_pin = ids_h.pin_memory(), mask_h.pin_memory()
    ids_d, mask_d, pos_d = ids_h.to(dev), mask_h.to(dev), pos_h.to(dev)
    qb = shard_bounds(nq, world)
    q_lo, q_hi = qb[rank], qb[rank + 1]
    q_counts = [qb[r + 1] - qb[r] for r in range(world)]

    # ---- index: synthetic gaussian rows + one planted row per query (needs the query embeddings once)
    with torch.no_grad():
        _, emb_all, _ = txt_model(ids_d, mask_d, pos_d)
    bounds = shard_bounds(n, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    x = build_shard(lo, hi, dev)
    prow, gt = planted_rows(emb_all, n)
    mine = (gt >= lo) & (gt < hi)
    x[gt[mine] - lo] = prow[mine]
    indexer = ShardedFlatIndexer(D)
    all_ids = [f"img_{i:07d}.npz" for i in range(n)]
    indexer.index_shard(all_ids, x, bounds)
    indexer.index._finalize()
    del prow
    barrier()

    def step_device():
        with torch.no_grad():
            _, emb, _ = txt_model(ids_d[q_lo:q_hi], mask_d[q_lo:q_hi], pos_d, need_sequence=False)
            q_all = indexer.gather_queries(emb, q_counts)
            # (no host synchronisation inside the step: the uncertified-query count is looked at after the timed loop)
            return indexer.search_device(q_all, k, lazy_flags=True) if world > 1 else indexer.search_device(q_all, k)

    def step_e2e(api="search"):
        """Pinned host token ids -> host results.  api="search": the faiss-level call (scores, labels as numpy
        arrays - what faiss_indexers.py:83 receives); api="search_knn": DenseFlatIndexer.search_knn, which
        additionally materialises the [(db_id list, scores)] Python structure of faiss_indexers.py:85-87."""
        with torch.no_grad():
            ids = ids_pin[q_lo:q_hi].to(dev, non_blocking=True)
            mask = mask_pin[q_lo:q_hi].to(dev, non_blocking=True)
            _, emb, _ = txt_model(ids, mask, pos_d, need_sequence=False)   # (as BiEncoder.forward calls it)
            q_all = indexer.gather_queries(emb, q_counts)
            if world > 1:   # the results are consumed on rank 0: one device -> host copy, one Python list build
                return indexer.search(q_all, k, host_rank=0) if api == "search" else indexer.search_knn(q_all, k, host_rank=0)
            return indexer.search(q_all, k) if api == "search" else indexer.search_knn(q_all, k)

    # ---- value: device-resident inputs, CUDA events, per-kernel accounting on
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    _lib.prof_reset()
    _lib.prof_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        scores, ids_out = step_device()
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    _lib.prof_enable(False)
    prof = _lib.prof_read()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_step = ms_total / args.steps
    value = nq / (ms_step * 1e-3)
    flagged = indexer.pending_flags() if world > 1 else indexer.index.last_flagged
    if world > 1 and flagged:   # (never seen on this workload) the lazy result would not be certified: redo it eagerly
        scores, ids_out = indexer.search_device(indexer.gather_queries(
            txt_model(ids_d[q_lo:q_hi], mask_d[q_lo:q_hi], pos_d, need_sequence=False)[1], q_counts), k)
    recall = recall_at(ids_out, gt)
    sorted_ok = bool((scores[:, 1:] <= scores[:, :-1]).all().item())

    # ---- e2e: pinned host token ids -> host results through the reference-facing calls
    e2e = None
    if not args.no_e2e:
        def time_e2e(api):
            for _ in range(max(1, min(args.warmup, 3))):
                step_e2e(api)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                res = step_e2e(api)
            barrier()
            dt = (time.perf_counter() - t0) / args.steps
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return dt, res
        dt, res = time_e2e("search")
        dt_knn, res_knn = time_e2e("search_knn")
        if rank == 0:
            assert res[0].shape == (nq, k) and res[1].shape == (nq, k) and res[1].dtype.name == "int64"
            assert len(res_knn) == nq and len(res_knn[0][0]) == k
        e2e = {"value": nq / dt, "unit": UNIT,
               "h2d_bytes_per_step": int((q_hi - q_lo) * L * 8 * 2),
               "d2h_bytes_per_step": int(nq * k * (4 + 8)),
               "ms_per_step": dt * 1e3,
               "api": "BertEncoder.forward + index.search (faiss-level: numpy scores + int64 labels on the host)",
               "search_knn_value": nq / dt_knn, "search_knn_ms_per_step": dt_knn * 1e3,
               "search_knn_note": "same, through DenseFlatIndexer.search_knn: adds the Python [(db_id list, scores)] "
                                  "materialisation of faiss_indexers.py:85-87 (host-side, nq * k object references)",
               "timer": "host wall clock around the API calls (sync on both sides)"}

    # ---- the same device-resident step with fp16 towers (what `fp16: true` of the reference's shipped eval config selects,
    # config/flickr30k_eval_config.json:19); the headline above is bf16, the dtype north_star names
    fp16_line = None
    if not args.no_e2e:
        txt_model.compute_dtype = torch.float16
        for _ in range(2):
            step_device()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            s16, i16 = step_device()
        f1.record()
        barrier()
        ms16 = f0.elapsed_time(f1)
        if world > 1:
            t = torch.tensor([ms16], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms16 = float(t.item())
        fp16_line = {"value": nq / (ms16 / args.steps * 1e-3), "unit": UNIT, "ms_per_step": ms16 / args.steps,
                     "recall_planted": recall_at(i16, gt), "dtype": "fp16 towers (bf16 is the headline dtype)"}
        txt_model.compute_dtype = torch.bfloat16

    # ---- certificate margins (diagnostic): error bound E of the coarse pass vs the score gap between rank k and k'
    cert = None
    if rank == 0:
        ix = indexer.index
        kp = k + max(k // 2, 32)
        kp = (max(kp, 64) + 31) // 32 * 32
        qs = emb_all[:512].contiguous()
        s_kp, _ = ix.search_device(qs, min(kp, 1024))
        q16 = qs.half().float()
        nq16, rq = q16.norm(dim=1).double(), (qs - q16).norm(dim=1).double()
        rmax, xmax = (float(v) for v in ix._xstats.tolist())
        E = nq16 * rmax + rq * (xmax + rmax) + D * 2.384185791015625e-07 * nq16 * xmax
        gap = (s_kp[:, k - 1] - s_kp[:, -1]).double()
        cert = {"coarse_k": kp, "x_residual_max": rmax, "x_norm_max": xmax, "E_median": float(E.median()),
                "gap_k_to_coarse_k_median": float(gap.median()), "gap_over_E_min": float((gap / E).min()),
                "gap_over_E_median": float((gap / E).median())}

    # ---- online regime (HBM-bound: one 128-query tile per index pass), coarse kernel only
    online = None
    if rank == 0:
        q128 = emb_all[:128].contiguous()
        for _ in range(3):
            indexer.index.search_device(q128, k)
        torch.cuda.synchronize()
        _lib.prof_reset()
        _lib.prof_enable(True)
        for _ in range(10):
            indexer.index.search_device(q128, k)
        torch.cuda.synchronize()
        _lib.prof_enable(False)
        p128 = _lib.prof_read()
        c = p128["coarse_score_topk"]
        tot = sum(v["ms"] for v in p128.values())
        gbs = c["bytes"] / (c["ms"] * 1e-3) / 1e9 if c["ms"] > 0 else 0.0
        online = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                  "traffic": None, "kernel": "coarse_score_topk", "queries_per_pass": 128,
                  "shard_rows": int(hi - lo), "us_per_launch": 1e3 * c["ms"] / max(1, c["timed"]),
                  "search_us_total": 1e3 * tot / 10, "peak_source": peaks["source"]}
        if world == 1 and not args.no_e2e:
            # retrieve_query / rerank.py regime end to end (host token ids -> host top-k): eager calls vs ONE CUDA graph
            from lightningdot_b200.online import GraphedRetriever
            lat = {}
            for qb in (1, 128):
                gr = GraphedRetriever(txt_model, indexer.index, batch=qb, seq_len=L, k=k)
                qi, qm = ids_h[:qb], mask_h[:qb]

                def eager():
                    with torch.no_grad():
                        _, e, _ = txt_model(qi.to(dev), qm.to(dev), pos_d, need_sequence=False)
                    return indexer.index.search(e, k)
                for fn, tag in ((lambda: gr.search(qi, qm), "graph"), (eager, "eager")):
                    for _ in range(3):
                        res = fn()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    for _ in range(20):
                        res = fn()
                    lat[f"{tag}_us_per_call_batch{qb}"] = 1e6 * (time.perf_counter() - t0) / 20
                same = bool((gr.search(qi, qm)[1] == eager()[1]).all())
                lat[f"graph_equals_eager_batch{qb}"] = same
                del gr
            online["e2e_latency"] = lat
    # ---- index-build side of the path: UNITER-base image tower over synthetic region features
    index_build = None
    if rank == 0 and world == 1 and not args.no_e2e:
        from lightningdot_b200.bi_encoder import UniterEncoder
        img_model = UniterEncoder(TowerConfig(vocab_size=synth.VOCAB), project_dim=D)
        img_model.load_state_dict(synth.random_tower_state("img", seed=43), strict=True)
        img_model.to(dev).eval()
        nb_img, regions = 4096, 36
        ib = {k_: (v.to(dev) if torch.is_tensor(v) else v) for k_, v in synth.image_batch(nb_img, regions, seed=5).items()}

        def encode_images():
            with torch.no_grad():
                return img_model(ib["input_ids"], ib["attention_mask"], ib["position_ids"], ib["img_feat"],
                                 ib["img_pos_feat"], None, ib["gather_index"], need_sequence=False)[1]
        for _ in range(2):
            encode_images()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            img_emb = encode_images()
        e1.record()
        torch.cuda.synchronize()
        ms_img = e0.elapsed_time(e1) / 5
        index_build = {"images_per_s": nb_img / (ms_img * 1e-3), "ms_per_batch": ms_img, "batch_images": nb_img,
                       "regions": regions, "feat_dim": 2048,
                       "model_tflops": 6.45e9 * nb_img / (ms_img * 1e-3) / 1e12,
                       "note": "UniterEncoder.forward on device-resident fp32 region features (36 x 2048 + 7 box "
                               "coordinates per image), 6.45 GFLOP per image (SURVEY 8); a 1M-image index = "
                               f"{1e6 / (nb_img / (ms_img * 1e-3)):.1f} s on one GPU"}
        del img_model, ib, img_emb
    barrier()

    # ---- BASELINE configs[4]: the train_itm.py step at 512 pairs per GPU (global batch 512 x N, 4096 at N = 8), dropout on,
    # symmetric in-batch NLL over the GLOBAL batch (embedding all-gather), gradient average, clip, AdamW
    train_step = None
    if args.train_steps > 0 and not args.no_e2e:
        import importlib.util
        spec = importlib.util.spec_from_file_location("bench_train", os.path.join(ROOT, "scripts", "bench_train.py"))
        bt = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bt)
        x_host_keep = x.cpu() if (world == 1 and not args.no_cpu_baseline) else None   # (the CPU legs below need the rows)
        del indexer, x
        x = x_host_keep
        torch.cuda.empty_cache()
        try:
            train_step = bt.measure(bt.default_args(steps=args.train_steps, warmup=3), rank, world, local_rank, dev)
        except Exception as exc:   # noqa: BLE001 - the headline line must not depend on the secondary leg
            import traceback
            traceback.print_exc(file=sys.stderr)
            train_step = {"error": f"{type(exc).__name__}: {exc}"[:300]} if rank == 0 else None
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines from the per-launch event accounting of the timed region
    total_kernel_ms = sum(v["ms"] for v in prof.values())
    shares = {name: {"ms_per_step": v["ms"] / args.steps, "share": v["ms"] / total_kernel_ms if total_kernel_ms else 0.0,
                     "launches_per_step": v["launches"] / args.steps}
              for name, v in prof.items() if v["launches"]}
    lin, coarse = prof["linear_tcgen05"], prof["coarse_score_topk"]

    def tensor_roof(v, name):
        tf = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0.0
        return {"bound": "tensor", "achieved": tf, "peak": peaks["tc_sustained"], "unit": "TFLOP/s",
                "frac": tf / peaks["tc_sustained"], "traffic": None, "kernel": name,
                "us_per_launch": 1e3 * v["ms"] / max(1, v["timed"]), "peak_source": peaks["source"] + " (sustained bf16)"}

    dominant = max(prof.items(), key=lambda kv: kv[1]["ms"])[0]
    roof_lin = tensor_roof(lin, "linear_tcgen05 (linear_tc_kernel + linear_ln2_kernel: O-proj+LN, FFN-up+GELU, FFN-down+LN, head)")
    roof_qa = tensor_roof(prof["qkv_attention"], "qkv_attention (qkv_attn_kernel: Q|K|V projection + attention, one kernel)") \
        if prof.get("qkv_attention", {}).get("launches") else None
    roof_search = tensor_roof(coarse, "coarse_score_topk (coarse_ts_kernel<EpiTopK>)")
    roof_search["queries_per_pass"] = nq
    roofline = roof_lin if dominant == "linear_tcgen05" else roof_search
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):   # dram bytes per launch from the committed ncu --set full captures
        with open(traffic_file) as f:
            tr = json.load(f)
        shapes = tr.get("linear_tcgen05_shapes")
        if shapes:   # per-shape captures at the bench's own M (scripts/ncu_shapes.sh): launch-weighted mean over a layer
            per = [v["dram_bytes"] for v in shapes.values()]
            roof_lin["traffic"] = sum(per) / len(per)
            roof_lin["traffic_per_shape"] = shapes
        else:
            roof_lin["traffic"] = tr.get("linear_tcgen05")
        roof_search["traffic"] = tr.get("coarse_score_topk")
        if online:
            online["traffic"] = tr.get("coarse_score_topk_online")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, world),
        "e2e": e2e,
        "gpu_launches": int(sum(v["launches"] for v in prof.values())),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_qkv_attention": roof_qa,
        "roofline_search": roof_search,
        "roofline_online": online,
        "index_build": index_build,
        "fp16_towers": fp16_line,
        "train_step": train_step,
        "kernel_shares": shares,
        "parity": {"recall_planted": recall, "scores_sorted": sorted_ok, "flagged_queries_last_step": int(flagged),
                   "certificate": cert},
    }

    # ---- CPU baseline + parity check against the oracle on a bounded sample (rank 0, N = 1)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], par = cpu_baseline(args, sd, ids_h, mask_h, pos_h, x, emb_all, ids_out, scores, gt)
        line["parity"].update(par)
    _emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, sd, ids_h, mask_h, pos_h, x_dev, emb_all, ids_gpu, scores_gpu, gt):
    import numpy as np
    import torch
    from oracle import flatip
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    s, k = min(args.cpu_sample, ids_h.shape[0]), args.k
    x_host = x_dev.cpu()   # (already on the host when the train_step leg released the device copy)
    # (a) the timed leg: the reference's CPU path on the sample
    t0 = time.perf_counter()
    emb_cpu = cpu_text_tower(sd, ids_h[:s], mask_h[:s], pos_h)
    t1 = time.perf_counter()
    cs, ci = cpu_search_f32(emb_cpu, x_host, k)
    t2 = time.perf_counter()
    qps = s / (t2 - t0)
    base = {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"first {s} of the {ids_h.shape[0]} queries through the fp32 torch-CPU restatement of the reference "
                       f"text tower ({s / (t1 - t0):.1f} q/s) + fp32 sgemm/top-{k} flat-IP search over the full "
                       f"{x_host.shape[0]} x {D} index ({s / (t2 - t1):.1f} q/s)")}
    # (b) parity at the index boundary: the oracle's exact search on the GPU-produced embeddings (same inputs)
    m = min(args.parity_sample, ids_h.shape[0])
    q_np = emb_all[:m].cpu().numpy()
    x_np = x_host.numpy()
    os_, oi = flatip.search_blocked(q_np, x_np, k, row_block=125000)   # fp64-accumulated exact search, (score desc, id asc)
    gi = ids_gpu[:m].cpu().numpy()
    gs = scores_gpu[:m].cpu().numpy()
    rel = float(np.max(np.abs(gs - os_) / np.maximum(np.abs(os_), 1e-30)))
    tower_cos = float(torch.nn.functional.cosine_similarity(emb_all[:s].cpu(), emb_cpu, dim=1).min().item())
    gt_h = gt[:m].cpu().numpy()
    rec_cpu = {str(t): float(np.mean((oi[:, :t] == gt_h[:, None]).any(axis=1))) for t in (1, 5, 10)}
    rec_gpu = {str(t): float(np.mean((gi[:, :t] == gt_h[:, None]).any(axis=1))) for t in (1, 5, 10)}
    par = {"oracle_sample_queries": m, "ids_identical_to_oracle": bool(np.array_equal(gi, oi)),
           "max_rel_score_err_vs_oracle": rel, "recall_sample_oracle": rec_cpu, "recall_sample_gpu": rec_gpu,
           "tower_embedding_min_cosine_vs_fp32_cpu": tower_cos,
           "note": "index-boundary parity (same embeddings into both searches); end-to-end Recall@1/5/10 identity through the "
                   "16-bit towers is tests/test_gpu_configs0.py (BASELINE configs[0] at full size vs the reference's own run)"}
    return base, par


_emit = print


def main():
    args = parse_args()
    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner, torchrun notes) is
    # sent to stderr by pointing fd 1 at fd 2 for the duration of the run; the line itself goes to the saved fd.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    global _emit
    _emit = lambda text: (real_stdout.write(text + "\n"), real_stdout.flush())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
