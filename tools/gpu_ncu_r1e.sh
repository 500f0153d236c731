#!/bin/bash
# ncu --set full captures of the kernels that changed late in round 1 (plain run first; ncu only if it exits 0)
mkdir -p gpurun_out
CMD="python tools/profile_step.py 64 3 1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -n 5 gpurun_out/plain.log; exit 1; }
cap() {  # name regex skip count [env...]
  env "${@:5}" ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s "$3" -c "$4" -f -o "gpurun_out/full_$1" $CMD > "gpurun_out/ncu_$1.log" 2>&1
  echo "$1 exit $?"
}
cap mel 'mel_kernel' 0 1
cap mel_clamp 'mel_clamp_kernel' 0 1
cap dec_attn 'decode_attn_mma_kernel' 30 1
cap skinny 'gemm_skinny_kernel' 40 3
PROFILE_KNOBS=1 PROFILE_RATE=24000 $CMD > gpurun_out/plain_knobs.log 2>&1 && {
  cap sample '::sample_kernel' 1 1 PROFILE_KNOBS=1 PROFILE_RATE=24000
  cap resample '::resample_kernel' 3 1 PROFILE_KNOBS=1 PROFILE_RATE=24000
}
tail -n 2 gpurun_out/plain_knobs.log
ls -la gpurun_out | grep full_
