#!/bin/bash
# the other BASELINE configs on one GPU (per-GPU share of the 8-GPU configurations), no CPU leg
mkdir -p gpurun_out
python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "config3 exit $?"
python bench.py --model 0.6B --clips-per-gpu 1 --no-cpu-baseline --steps 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "config2 exit $?"
python bench.py --model 1.7B --clip-seconds 15 --clips-per-gpu 64 --max-tokens 448 --no-cpu-baseline --steps 3 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "config4 exit $?"
python bench.py --model 1.7B --clip-seconds 30 --clips-per-gpu 15 --max-tokens 128 --no-cpu-baseline --steps 3 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "config5 exit $?"
python - <<'PY'
import json
for f in ("bench", "bench_cfg2", "bench_cfg4", "bench_cfg5"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        r = d.get("roofline") or {}
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()},
              "| roofline", r.get("kernel"), round(r.get("frac", 0), 3), (r.get("in_graph") or {}).get("frac"))
    except Exception as e:
        print(f, "failed", e)
PY
