#!/bin/bash
for size in "0.6B 64 30" "1.7B 64 15"; do
  for st in 4 6 8; do
    Q3ASR_SK_STAGES=$st timeout 200 python tools/decode_time.py $size 96 "stages=$st" 2>&1 | tail -n 1
  done
  for dk in 4 8 16; do
    Q3ASR_SK_DEEP_KB=$dk timeout 200 python tools/decode_time.py $size 96 "policy deep_kb=$dk" 2>&1 | tail -n 1
  done
done
