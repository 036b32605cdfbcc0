"""Turn the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py r01

  profiles/<round>_launches_bench.csv   every launch of `bench.py --steps 2 --warmup 1` aggregated per kernel
                                        (count, total / mean device time, share) from the gpu__time_duration pass
  profiles/<round>_ncu_<name>.txt       key raw metrics + stall mix of the `--set full` capture of one kernel
  profiles/traffic.json                 dram bytes per launch of the hot kernels (bench.py's roofline.traffic)
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
           "sm__inst_executed.sum.per_cycle_elapsed", "l1tex__data_bank_conflicts_pipe_lsu.sum",
           "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second"]


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").strip()[:110]


def launches(tag, name="launches_bench", what="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"):
    src = os.path.join(SRC, name + ".csv")
    if not os.path.exists(src):
        return
    with open(src) as f:
        lines = [l for l in f if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = (short(row["Kernel Name"]), row["Grid Size"], row["Block Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    dst = os.path.join(OUT, f"{tag}_{name}.csv")
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([f"# ncu --metrics gpu__time_duration.sum --clock-control none ... {what} "
                    "(cold-cache, serialised: compare SHARES, not absolutes)"])
        w.writerow(["kernel", "grid", "block", "launches", "total_ms", "mean_us", "share"])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k[0], k[1], k[2], v[0], f"{v[1] / 1e6:.3f}", f"{v[1] / v[0] / 1e3:.1f}", f"{v[1] / tot:.4f}"])
    print("wrote", dst)


def ncu_report(tag, rep, name, traffic_key=None, kernel_index=0, traffic=None):
    path = os.path.join(SRC, rep)
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none --import-source on ({rep}); kernel {kernel_index} of the capture"]
    if 2 + kernel_index >= len(rows):
        return
    data = rows[2 + kernel_index]
    col = {h: i for i, h in enumerate(hdr)}
    lines.append("kernel: " + data[col["Kernel Name"]][:200])
    vals = {}
    for m in METRICS:
        if m in col:
            lines.append(f"{m} [{units[col[m]]}] = {data[col[m]]}")
            vals[m] = (data[col[m]], units[col[m]])
    stalls = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_stalls.py"), path, str(kernel_index), "12"],
                            capture_output=True, text=True).stdout
    lines.append("")
    lines.append("# warp stall samples (source page, SASS): mix + the 12 hottest instructions")
    lines.extend(stalls.splitlines()[1:])
    dst = os.path.join(OUT, f"{tag}_ncu_{name}.txt")
    with open(dst, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", dst)
    if traffic is not None and traffic_key:
        def to_bytes(v, u):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
            return float(v) * scale
        r, w = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
        if r and w:
            traffic[traffic_key] = to_bytes(*r) + to_bytes(*w)


SHAPES = {  # scripts/ncu_shapes.sh captures at M = 320 000 tokens: name -> (N, K, reads a residual)
    "qkv": (2304, 768, False), "oproj_ln": (768, 768, True), "ffn1_gelu": (3072, 768, False), "ffn2_ln": (768, 3072, True),
    "qkv_attn": (768, 768, False)}   # (fused projection + attention: reads x + the stacked weight, writes ctx only)


def shape_traffic(tag, traffic, M=320000):
    """Per-shape dram bytes of one launch at the bench's own size next to the algorithmic bytes (A + W + residual + out,
    16-bit): profiles/traffic.json["linear_tcgen05_shapes"] - what bench.py's roofline.traffic is read from."""
    out = {}
    for name, (N, K, res) in SHAPES.items():
        t = {}
        ncu_report(tag, f"shape_{name}.ncu-rep", f"shape_{name}_m{M}", "x", 0, t)
        if "x" in t:
            alg = 2.0 * (M * K + N * K + M * N + (M * N if res else 0)) + (2.0 * 2 * N * K if name == "qkv_attn" else 0.0)
            out[name] = {"M": M, "N": N, "K": K, "dram_bytes": t["x"], "algorithmic_bytes": alg, "ratio": t["x"] / alg}
    if out:
        traffic["linear_tcgen05_shapes"] = out


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    launches(tag, "launches_train", "python scripts/bench_train.py --steps 1 --warmup 1 (warm-up step included)")
    traffic = {}
    ncu_report(tag, "prof_coarse_10k.ncu-rep", "coarse_score_topk_nq10k", "coarse_score_topk", 0, traffic)
    ncu_report(tag, "prof_coarse_128.ncu-rep", "coarse_score_topk_nq128", "coarse_score_topk_online", 0, traffic)
    for i, nm in enumerate(["linear_qkv", "linear_oproj", "linear_ffn1_gelu", "linear_ffn2", "linear_next"]):
        ncu_report(tag, "prof_linear.ncu-rep", nm, "linear_tcgen05" if nm == "linear_ffn1_gelu" else None, i, traffic)
    ncu_report(tag, "prof_attention.ncu-rep", "attention_fwd", None, 0, None)
    # training-step kernels (scripts/ncu_train.sh)
    for rep, nm in (("train_ln_bwd16.ncu-rep", "train_ln_bwd_dropout"), ("train_gelu_pre.ncu-rep", "train_ffn_up_fused_epilogue"),
                    ("train_dgrad_gelu.ncu-rep", "train_ffn_up_dgrad"), ("train_attention_bwd.ncu-rep", "train_attention_bwd")):
        ncu_report(tag, rep, nm, None, 0, None)
    shape_traffic(tag, traffic)
    old = os.path.join(OUT, "traffic.json")
    if os.path.exists(old):   # keep entries this run had no capture for
        with open(old) as f:
            traffic = {**json.load(f), **traffic}
    if traffic:
        with open(os.path.join(OUT, "traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
        print("wrote traffic.json", traffic)


if __name__ == "__main__":
    main()
