#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_suite.sh
timeout -k 5 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2e_bench_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "pipelined", d["e2e_pipelined"] and d["e2e_pipelined"]["value"], "ms", d["ms_per_step"], d["stage_ms_per_step"])
print("roofline", {k: d["roofline"][k] for k in ("kernel","achieved","frac","ms_per_launch")})
print("mel", d["roofline_mel"] and {k: d["roofline_mel"][k] for k in ("achieved","frac")})
print("cpu", d["cpu_baseline"] and {k: d["cpu_baseline"][k] for k in ("value","cores")})
print("parity", d.get("parity"))
for k in ("e2e_pool","strong_scaling","config4","config5","error"):
    if k in d: print(k, d[k])
PY
