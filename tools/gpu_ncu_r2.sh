#!/bin/bash
# Round 2 evidence: ncu launch list of one bench-shaped step + --set full captures of the kernels that carry the step or changed
# this round (plain run first; ncu only if it exits 0).  Outputs under gpurun_out/: r2_launches.csv, full_r2_<name>.ncu-rep
mkdir -p gpurun_out
CMD="python tools/profile_step.py 64 4 1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
cap() {  # name regex skip count [env...]
  env "${@:5}" ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s "$3" -c "$4" -f -o "gpurun_out/full_r2_$1" $CMD > "gpurun_out/ncu_$1.log" 2>&1
  echo "$1 exit $?"
}
cap mel 'mel_kernel' 0 1
cap qkv_fused 'gemm_tc2_kernel<\(int\)256, \(int\)4>' 2 1
cap dec_attn 'decode_attn_mma_kernel' 30 1
cap attn_pre 'fa_tc_kernel<\(int\)128' 2 1
cap conv1 'conv1_kernel' 1 1
cap conv2 'gemm_tc_kernel<\(int\)160, \(int\)0>' 3 1
cap skinny 'gemm_skinny_kernel' 40 3
cap lmhead 'lmhead_argmax_kernel' 1 1
cap mega 'megastep_kernel' 1 1 Q3ASR_MEGA=1
ls -la gpurun_out | grep full_r2
