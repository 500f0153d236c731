#!/bin/bash
# One GPU session: parity tests file by file (each under its own timeout so a hung kernel cannot eat the
# whole lease), logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for t in "$@"; do
  name=$(basename "$t" .py)
  timeout 600 python -m pytest "$t" -q -m gpu -x --no-header -p no:cacheprovider > "gpurun_out/${name}.log" 2>&1
  echo "$name exit $?" | tee -a gpurun_out/summary.txt
  tail -n 25 "gpurun_out/${name}.log"
done
