#!/bin/bash
# the default bench on all GPUs of the box (final build)
set -x
mkdir -p gpurun_out
N=${1:-8}
timeout 200 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/g8_bench_n$N.json 2> gpurun_out/g8_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/g8_*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("fraction_of_copy_only_ceiling"), d.get("segments_check",{}).get("pixels_within_1_and_identical"))
PY
tail -3 gpurun_out/g8_err.log
