#!/bin/bash
for i in 1 2; do
  Q3ASR_PKG=$PWD/ab_old/pkg timeout 200 python tools/decode_time.py 0.6B 64 30 128 "old (fd4ca33)" 2>&1 | tail -n 1
  timeout 200 python tools/decode_time.py 0.6B 64 30 128 "new" 2>&1 | tail -n 1
done
Q3ASR_PKG=$PWD/ab_old/pkg timeout 200 python tools/decode_time.py 1.7B 64 15 128 "old (fd4ca33)" 2>&1 | tail -n 1
timeout 200 python tools/decode_time.py 1.7B 64 15 128 "new" 2>&1 | tail -n 1
