#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_suite.sh tests/test_gpu_model.py tests/test_gpu_golden.py tests/test_gpu_compaction.py tests/test_gpu_megastep.py tests/test_gpu_pool.py
for v in 0 1; do
  Q3ASR_NO_QKV_FUSE=$v timeout -k 5 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipelined --no-extras > gpurun_out/r2d_bench_nofuse$v.json 2> gpurun_out/r2d_bench_nofuse$v.err
  echo "bench NO_QKV_FUSE=$v rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2d_bench_nofuse$v.json"))
print(d["ms_per_step"], d["stage_ms_per_step"], d["value"])
for k,v in sorted(d["kernel_families"].items(), key=lambda kv:-kv[1]["est_ms_in_step"]):
    if k.startswith("pre_"): print("   ", k, round(v["ms_per_step"],3), v["launches_per_step"], v["tflops"] and round(v["tflops"]), v["gbs"] and round(v["gbs"]))
PY
done
