#!/bin/bash
# mel kernel check: parity tests, then the stage time for the bench batch (64 x 30 s)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mel.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/test_gpu_mel.log 2>&1
echo "test_gpu_mel exit $? $(tail -n 1 gpurun_out/test_gpu_mel.log)"
timeout 120 python tools/mel_time.py 2>&1 | tee gpurun_out/mel_time.log
