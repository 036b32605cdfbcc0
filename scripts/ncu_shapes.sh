#!/bin/bash
# ncu --set full capture of every encoder GEMM shape AT THE BENCH'S OWN SIZE (M = 320 000 tokens = 10k captions x 32), one
# launch per shape, so that roofline.traffic can be read per shape (profiles/traffic.json).
# usage: gpurun --timeout 900 -- 'bash scripts/ncu_shapes.sh [tokens]'
M="${1:-320000}"
mkdir -p gpurun_out
cap() {  # name, kernel regex, case
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s 3 -c 1 -f \
    -o "gpurun_out/shape_$1" python scripts/bench_linear.py "$M" 0.01 "$3" > "gpurun_out/shape_$1.out" 2>&1; echo "ncu $1 rc=$?"
}
cap qkv 'linear_tc_kernel' qkv
cap qkv_attn 'qkv_attn_kernel' qkv+attn
cap oproj_ln 'linear_ln' o+ln
cap ffn1_gelu 'linear_tc_kernel' ffn1+gelu
cap ffn2_ln 'linear_ln' ffn2+ln
timeout 300 python scripts/bench_linear.py "$M" 1.0 > gpurun_out/bench_linear.txt 2>&1; cat gpurun_out/bench_linear.txt
