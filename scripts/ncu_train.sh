#!/bin/bash
# ncu --set full captures of the training step's own kernels (one launch each, steady-state launch picked with -s):
# LayerNorm backward with fused dropout, the GELU + pre-activation forward GEMM, the GELU' dgrad, attention backward.
# usage: gpurun --timeout 900 -- 'bash scripts/ncu_train.sh'
mkdir -p gpurun_out
cap() {  # name, kernel regex, launches to skip
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s "$3" -c 1 -f \
    -o "gpurun_out/train_$1" python scripts/bench_train.py --steps 1 --warmup 1 --no-graph > "gpurun_out/train_$1.out" 2>&1; echo "ncu $1 rc=$?"
}
cap ln_bwd16 'ln_bwd16_kernel' 60
cap gelu_pre 'linear_tc_kernel<.int.1, .int.0, .int.2, .int.0, .int.0, .int.0, .int.1>' 30
cap dgrad_gelu 'linear_tc_kernel<.int.2, .int.0, .int.2, .int.0, .int.1, .int.0, .int.0>' 30
cap attention_bwd 'attention_bwd_kernel<.int.48' 14
ls -la gpurun_out | grep train_
