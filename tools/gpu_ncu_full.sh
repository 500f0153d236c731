#!/bin/bash
# ncu --set full captures of one launch per hot kernel, 64 x 30 s clips, 3 tokens (plain run first; ncu only if it exits 0).
mkdir -p gpurun_out
CMD="python tools/profile_step.py 64 3 1"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s "$3" -c "$4" -f -o "gpurun_out/full_$1" $CMD > "gpurun_out/ncu_$1.log" 2>&1
  echo "$1 exit $?"
}
cap mel 'mel_kernel' 0 1
cap gateup 'gemm_tc_kernel<256, 1>' 2 1
cap fc1 'gemm_tc_kernel<256, 0>' 3 1
cap attn_pre 'flash_attn_kernel<128, true>' 2 1
cap attn_enc 'flash_attn_kernel<64, false>' 2 1
cap rope 'qknorm_rope_kv_kernel' 2 1
cap dec_attn 'decode_attn_fused_kernel' 30 1
cap conv1 'conv1_kernel' 3 1
cap skinny 'gemm_skinny_kernel<64, 0>' 40 3
ls -la gpurun_out | grep full_
