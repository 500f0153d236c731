#!/bin/bash
# tests (each file under its own timeout) + one bench run, no ncu
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
for t in "$@"; do
  name=$(basename "$t" .py)
  timeout 240 python -m pytest "$t" -q -m gpu -x --no-header -p no:cacheprovider > "gpurun_out/${name}.log" 2>&1
  echo "$name exit $?" | tee -a gpurun_out/summary.txt
  tail -n 12 "gpurun_out/${name}.log"
done
timeout 240 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], d["stage_ms_per_step"])
    for k, v in sorted(d["kernel_families"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
        print(f"{k:20s} {v['ms_per_step']:9.3f} ms x{v['launches_per_step']:4d} tf={v['tflops'] and round(v['tflops'],1)} gbs={v['gbs'] and round(v['gbs'],1)}")
except Exception as e:
    print("no bench json", e)
PY
