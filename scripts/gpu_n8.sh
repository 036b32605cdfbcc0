#!/bin/bash
# 8-GPU validation: bench line at N = 8 (BASELINE configs[3] sharded + configs[4] train_step at global batch 4096) and N = 1 on the same box
N="${1:-8}"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo rc=$?; tail -2 gpurun_out/bench_n$N.err
python bench.py --no-cpu-baseline --train-steps 0 > gpurun_out/bench_n1_samebox.json 2>/dev/null
python - <<PY
import json
for n in ("n$N", "n1_samebox"):
    b = json.load(open(f"gpurun_out/bench_{n}.json"))
    print(n, round(b["value"]), round(b["ms_per_step"], 3), "e2e", round(b["e2e"]["value"]), "knn", round(b["e2e"]["search_knn_value"]),
          {k: round(v["ms_per_step"], 2) for k, v in b["kernel_shares"].items()})
    if b.get("train_step"):
        t = b["train_step"]; print("  train", round(t["ms_per_step"], 2), "ms", round(t["pairs_per_s"]), "pairs/s", t["kernel_ms_per_step"])
    if b.get("roofline_online"):
        o = b["roofline_online"]; print("  online", round(o["frac"], 3), o["us_per_launch"], o["search_us_total"])
PY
