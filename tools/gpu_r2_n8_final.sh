#!/bin/bash
# round-2 final multi-GPU call (8 GPUs, final build): default bench, configs[3] for real, segment mode with the device exchange
set -x
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node $N --master-port 29514 bench.py --gpus $N --steps 10 --warmup 5 > gpurun_out/f8_bench_n$N.json 2> gpurun_out/f8_err.log
timeout 240 $TR --nproc-per-node $N --master-port 29512 bench.py --gpus $N --config batch4096 --total $((512*N)) --steps 3 --warmup 2 > gpurun_out/f8_batch4096_n$N.json 2>> gpurun_out/f8_err.log
timeout 120 $TR --nproc-per-node $N --master-port 29517 bench.py --gpus $N --segments --steps 10 --warmup 3 > gpurun_out/f8_segments_n$N.json 2>> gpurun_out/f8_err.log
timeout 120 $TR --nproc-per-node 2 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 5 > gpurun_out/f8_bench_n2.json 2>> gpurun_out/f8_err.log
tail -n 12 gpurun_out/f8_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/f8_*.json")):
    try:
        d=json.load(open(f))
        keys=[k for k in ("value","ms_per_step","parity","segments_check") if k in d]
        print(f.split('/')[-1], {k:d[k] for k in keys}, "e2e", (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("copy_only_ceiling"))
    except Exception as e:
        print(f, "ERR", e)
PY
