#!/bin/bash
# round-end check: every GPU test file (own timeout each), the bench as the driver runs it, then the ncu launch list of a short pass
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for t in tests/test_*.py; do
  name=$(basename "$t" .py)
  timeout 600 python -m pytest "$t" -q -m gpu -x --no-header -p no:cacheprovider > "gpurun_out/${name}.log" 2>&1
  echo "$name exit $? $(tail -n 1 gpurun_out/${name}.log)" | tee -a gpurun_out/summary.txt
done
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt; tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], d["stage_ms_per_step"])
    print("roofline", {k: d["roofline"][k] for k in ("kernel", "achieved", "peak", "frac", "traffic")}, d["roofline"].get("in_graph"))
    print("roofline_mel", {k: d["roofline_mel"][k] for k in ("achieved", "frac", "traffic")})
    print("cpu", d["cpu_baseline"])
    print("clocks", d["clocks"])
except Exception as e:
    print("no bench json", e)
PY
python tools/profile_step.py 64 4 1 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches64.csv python tools/profile_step.py 64 4 1 > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
