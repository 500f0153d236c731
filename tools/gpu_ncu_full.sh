#!/bin/bash
# ncu --set full captures of one launch per hot kernel on the bench workload's shapes (64 x 30 s clips, 0.6B; 3 decode steps so
# the run stays short).  The plain run goes first; ncu only if it exits 0.  Outputs: gpurun_out/full_<name>.ncu-rep
mkdir -p gpurun_out
CMD="python tools/profile_step.py 64 3 1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/plain.log; exit 1; }
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s "$3" -c "$4" -f -o "gpurun_out/full_$1" $CMD > "gpurun_out/ncu_$1.log" 2>&1
  echo "$1 exit $?"
}
cap mel 'mel_kernel' 0 1
cap conv2 'gemm_tc_kernel<\(int\)160, \(int\)0>' 3 1
cap fc1 'gemm_tc_kernel<\(int\)256, \(int\)0>' 3 1
cap gateup 'gemm_tc_kernel<\(int\)256, \(int\)1>' 2 1
cap attn_pre 'fa_tc_kernel<\(int\)128' 2 1
cap attn_enc 'fa_tc_kernel<\(int\)64' 2 1
cap rope 'qknorm_rope_kv_kernel' 2 1
cap dec_attn 'decode_attn_mma_kernel' 30 1
cap skinny 'gemm_skinny_kernel' 40 3
cap dec_gateup 'gemm_tc_kernel<\(int\)64, \(int\)1>' 10 1
cap lm_head 'gemm_tc_kernel<\(int\)128, \(int\)3>' 1 1
ls -la gpurun_out | grep full_
