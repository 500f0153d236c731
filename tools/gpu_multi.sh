#!/bin/bash
# multi-GPU pass (gpurun --gpus N): the pool on distinct devices, then the bench under torchrun with the pool / strong-scaling /
# 1.7B sub-records.  Usage: bash tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt
timeout -k 5 600 python -m pytest tests/test_gpu_pool.py -q -s -m gpu > gpurun_out/multi_pool_n$N.log 2>&1
echo "pool tests rc=$? : $(tail -n 1 gpurun_out/multi_pool_n$N.log)"
timeout -k 5 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2d_bench_n$N.json 2> gpurun_out/r2d_bench_n$N.err
echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2d_bench_n$N.json"))
print("n", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pipelined", d["e2e_pipelined"] and round(d["e2e_pipelined"]["value"]), "ms", d["ms_per_step"])
for k in ("e2e_pool","strong_scaling","config2","throughput_mode","config4","config5","error"):
    if k in d: print(k, {kk: vv for kk, vv in d[k].items() if kk != "how"} if isinstance(d[k], dict) else d[k])
PY
tail -n 3 gpurun_out/r2d_bench_n$N.err
