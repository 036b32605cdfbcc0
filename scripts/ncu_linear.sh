#!/bin/bash
# ncu --set full captures of the encoder GEMM kernels at a large M (bench_linear.py shapes): GELU epilogue, fused LN, plain.
# usage: gpurun --timeout 900 -- 'bash scripts/ncu_linear.sh [tokens]'
M="${1:-131072}"
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s "$3" -c 1 -f -o "gpurun_out/prof_$1" \
    python scripts/bench_linear.py "$M" 0.01 > "gpurun_out/ncu_$1.out" 2>&1; echo "ncu $1 rc=$?"
}
cap lin_gelu 'linear_tc_kernel<.int.1, .int.0, .int.2' 3
cap lin_plain 'linear_tc_kernel<.int.0, .int.0, .int.2, .int.0, .int.0, .int.0' 3
