#!/bin/bash
# one --set full capture of the mel kernel on the bench batch (plain run first)
mkdir -p gpurun_out
python tools/mel_time.py > gpurun_out/mel_time.log 2>&1 || { echo plain run failed; tail -n 5 gpurun_out/mel_time.log; exit 1; }
cat gpurun_out/mel_time.log
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:mel_kernel -s 2 -c 1 -f -o gpurun_out/full_mel python tools/mel_time.py > gpurun_out/ncu_mel.log 2>&1
echo "ncu exit $?"
