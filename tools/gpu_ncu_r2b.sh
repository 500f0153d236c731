#!/bin/bash
# Round 2, second pass: --set full captures of the encoder products (q|k|v, out-proj, fc2 = three consecutive launches of the
# CTA-pair kernel with 224-column tiles; fc1 = the 256-column one with GELU), the fused prefill q|k|v and the prefill attention,
# after the epilogue changes (256-bit stores, prefetched residual rows, dimension-major RoPE table).
mkdir -p gpurun_out
CMD="python tools/profile_step.py 64 4 1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/plain.log; exit 1; }
cap() {  # name regex skip count [env...]
  env "${@:5}" ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s "$3" -c "$4" -f -o "gpurun_out/full_r2b_$1" $CMD > "gpurun_out/ncu_$1.log" 2>&1
  echo "$1 exit $?"
}
cap enc224 'gemm_tc2_kernel<\(int\)224, \(int\)0>' 3 3
cap fc1 'gemm_tc2_kernel<\(int\)256, \(int\)0>' 1 1
cap qkv_fused 'gemm_tc2_kernel<\(int\)256, \(int\)4>' 2 1
cap attn_pre 'fa_tc_kernel<\(int\)128' 2 1
ls -la gpurun_out | grep full_r2b
