#!/bin/bash
# One gpurun call: GPU tests, the bench line, the launch list and the ncu captures of the two hot kernels.
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tests] [bench] [launches] [ncu]'
set -u
what="${*:-tests bench launches ncu}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [[ "$what" == *tests* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
  tail -n 15 gpurun_out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
fi
if [[ "$what" == *bench* ]]; then
  timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  tail -c 6000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
fi
if [[ "$what" == *refarm* ]]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
  cat gpurun_out/bench_ref.json; tail -n 5 gpurun_out/bench_ref.err
fi
if [[ "$what" == *launches* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.out 2>&1; echo "launches rc=$?"
fi
if [[ "$what" == *ncu* ]]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:coarse_ts -s 1 -c 1 -f -o gpurun_out/prof_coarse_10k \
    python scripts/profile_search.py 1000000 10000 1 > gpurun_out/ncu_coarse_10k.out 2>&1; echo "ncu coarse10k rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:coarse_ts -s 1 -c 1 -f -o gpurun_out/prof_coarse_128 \
    python scripts/profile_search.py 1000000 128 1 > gpurun_out/ncu_coarse_128.out 2>&1; echo "ncu coarse128 rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_ -s 8 -c 4 -f -o gpurun_out/prof_linear \
    python scripts/profile_tower.py 4096 1 > gpurun_out/ncu_linear.out 2>&1; echo "ncu linear rc=$?"
fi
if [[ "$what" == *trainprof* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_train.csv \
    python scripts/bench_train.py --steps 1 --warmup 1 > gpurun_out/launches_train.out 2>&1; echo "train launches rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 14 -c 1 -f -o gpurun_out/prof_attention \
    python scripts/profile_tower.py 10000 1 > gpurun_out/ncu_attention.out 2>&1; echo "ncu attention rc=$?"
fi
ls -la gpurun_out | head -40
