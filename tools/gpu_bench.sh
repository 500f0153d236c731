#!/bin/bash
# bench + ncu launch list + ncu full captures of the top kernels (each ncu pass only after the same command exited 0).
mkdir -p gpurun_out
python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
python tools/profile_step.py 8 3 1 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 8 3 1 > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"; tail -n 3 gpurun_out/plain.log
python tools/profile_step.py 8 2 1 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'mel_kernel|gemm_tc_kernel|flash_attn_kernel|decode_attn_kernel|conv1_kernel' -c 40 -o gpurun_out/prof_r1 python tools/profile_step.py 8 2 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -n 3 gpurun_out/ncu_full.log
ls -la gpurun_out
