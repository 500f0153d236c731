#!/bin/bash
# A/B of environment switches on the bench workload: tools/gpu_ab.sh "NAME=VAL ..." "NAME=VAL ..." (one bench run per argument,
# "" = defaults); prints the stage times and the kernel families of each run.  No CPU baseline, no extras.
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  echo "=== run $i: ${envs:-defaults}"
  env $envs timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err || tail -n 5 gpurun_out/ab_$i.err
  python - "$i" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/ab_{sys.argv[1]}.json"))
    print("value %.0f e2e %.0f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
    r = d["roofline"]; print("roofline", r["kernel"], "frac %.3f" % r["frac"], "ms/launch %.4f" % r["ms_per_launch"], "standalone", r.get("standalone"))
    fam = d["kernel_families"]
    print(" ".join(f"{k}={v['ms_per_step']:.3f}" for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms_per_step"]) if k != "decode_graph_steps"))
except Exception as e:
    print("no bench json", e)
PY
done
