#!/bin/bash
# quick regression: the test files given as arguments, then a short bench without the CPU legs / extras; prints the family table
mkdir -p gpurun_out
bash tools/gpu_suite.sh "$@"
timeout -k 5 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipelined --no-extras > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/quick_bench.json"))
print(round(d["ms_per_step"],2), {k: round(v,2) for k,v in d["stage_ms_per_step"].items()}, round(d["value"]))
for k,v in sorted(d["kernel_families"].items(), key=lambda kv:-kv[1]["est_ms_in_step"]):
    if not k.startswith("dec_") and k != "decode_graph_steps": print("   ", f"{k:12s}", round(v["ms_per_step"],3), v["launches_per_step"], v["tflops"] and round(v["tflops"]), v["gbs"] and round(v["gbs"]))
PY
