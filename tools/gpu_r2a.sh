#!/bin/bash
# round 2, first GPU pass: megastep parity, full GPU suite, bench A/B (persistent decode kernel on / off)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
echo "== megastep tests" > gpurun_out/r2a_mega.log
timeout -k 5 600 python -m pytest tests/test_gpu_megastep.py -x -q -s >> gpurun_out/r2a_mega.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_mega.log
tail -5 gpurun_out/r2a_mega.log
echo "== golden tests" > gpurun_out/r2a_golden.log
timeout -k 5 900 python -m pytest tests/test_gpu_golden.py -q -s >> gpurun_out/r2a_golden.log 2>&1
echo "rc=$?" >> gpurun_out/r2a_golden.log
tail -5 gpurun_out/r2a_golden.log
echo "== bench mega on"
timeout -k 5 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipelined > gpurun_out/r2a_bench_mega.json 2> gpurun_out/r2a_bench_mega.err
echo "rc=$?"; cut -c1-600 gpurun_out/r2a_bench_mega.json
echo "== bench mega off"
Q3ASR_MEGA=0 timeout -k 5 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipelined > gpurun_out/r2a_bench_nomega.json 2> gpurun_out/r2a_bench_nomega.err
echo "rc=$?"; cut -c1-600 gpurun_out/r2a_bench_nomega.json
echo "== full gpu suite"
timeout -k 5 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_golden.py --deselect tests/test_gpu_megastep.py > gpurun_out/r2a_suite.log 2>&1
echo "rc=$?"; tail -8 gpurun_out/r2a_suite.log
