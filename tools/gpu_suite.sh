#!/bin/bash
# the GPU test suite file by file (each under its own timeout), logs in gpurun_out/suite_*.log, one summary line per file
mkdir -p gpurun_out
rm -f gpurun_out/suite_summary.txt
for t in ${@:-tests/test_gpu_*.py tests/test_host_cpp.py tests/test_sampler.py tests/test_aligner.py tests/test_audio_io.py}; do
  name=$(basename "$t" .py)
  timeout -k 5 900 python -m pytest "$t" -q -m gpu -s --no-header -p no:cacheprovider > "gpurun_out/suite_${name}.log" 2>&1
  echo "$name exit $? : $(tail -n 1 gpurun_out/suite_${name}.log)" | tee -a gpurun_out/suite_summary.txt
done
