#!/bin/bash
# decode step time at small batches: multi-kernel chain vs the persistent kernel (Q3ASR_MEGA=1, one or two sub-batches)
mkdir -p gpurun_out
for b in 1 8 16; do
  python tools/decode_time.py 0.6B $b 30 64 "chain b=$b"
  Q3ASR_MEGA=1 Q3ASR_MEGA_SUBS=1 python tools/decode_time.py 0.6B $b 30 64 "mega subs=1 b=$b"
  Q3ASR_MEGA=1 Q3ASR_MEGA_SUBS=2 python tools/decode_time.py 0.6B $b 30 64 "mega subs=2 b=$b"
done 2>&1 | tee gpurun_out/small_batch.log
