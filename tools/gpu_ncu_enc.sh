#!/bin/bash
# ncu --set full of the encoder / prefill GEMM families (plain run first)
mkdir -p gpurun_out
CMD="python tools/profile_step.py 64 2 1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/plain.log; exit 1; }
cap() {
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s "$3" -c "$4" -f -o "gpurun_out/full_$1" $CMD > "gpurun_out/ncu_$1.log" 2>&1
  echo "$1 exit $?"
}
cap tc2_224 'gemm_tc2_kernel<\(int\)224' 4 6
cap tc2_256 'gemm_tc2_kernel<\(int\)256' 4 6
cap tc_160 'gemm_tc_kernel<\(int\)160' 3 2
cap tc_256_0 'gemm_tc_kernel<\(int\)256, \(int\)0>' 3 1
