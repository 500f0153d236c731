#!/bin/bash
# decode attention variants: parity tests, then the decode step time of the bench batch with the split variant against two warps
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_golden.py tests/test_gpu_compaction.py -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/attn_split_tests.log 2>&1
echo "tests exit $? $(tail -n 1 gpurun_out/attn_split_tests.log)"
for w in 2 1; do Q3ASR_DECODE_ATTN_WARPS=$w python tools/decode_time.py 0.6B 64 30 128 "attn warps=$w"; done
python tools/decode_time.py 0.6B 64 30 128 "default"
python tools/decode_time.py 0.6B 48 30 128 "default b=48"
python tools/decode_time.py 1.7B 64 15 128 "default 1.7B"
Q3ASR_DECODE_ATTN_WARPS=2 python tools/decode_time.py 1.7B 64 15 128 "1.7B warps=2"
