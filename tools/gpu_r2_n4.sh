#!/bin/bash
# the default bench at 4 and 2 GPUs (final build)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 150 $TR --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/h4_bench_n4.json 2> gpurun_out/h4_err.log
timeout 150 $TR --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/h4_bench_n2.json 2>> gpurun_out/h4_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/h4_*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("fraction_of_copy_only_ceiling"))
PY
