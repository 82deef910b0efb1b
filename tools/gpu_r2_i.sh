#!/bin/bash
# round-2 GPU call I (2 GPUs): device-resident segment exchange, batch config parity subset
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_segments.py -x -q -m gpu > gpurun_out/i_test_segments.log 2>&1
echo "exit $?" >> gpurun_out/i_test_segments.log
$TR --nproc-per-node 2 --master-port 29516 bench.py --gpus 2 --segments --steps 10 --warmup 3 > gpurun_out/i_segments_n2.json 2> gpurun_out/i_err.log
WEFAX_SEG_HOST=1 $TR --nproc-per-node 2 --master-port 29517 bench.py --gpus 2 --segments --steps 10 --warmup 3 > gpurun_out/i_segments_n2_host.json 2>> gpurun_out/i_err.log
python bench.py --segments --steps 10 --warmup 3 > gpurun_out/i_segments_n1.json 2>> gpurun_out/i_err.log
$TR --nproc-per-node 2 --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_n2.json 2>> gpurun_out/i_err.log
$TR --nproc-per-node 2 --master-port 29519 bench.py --gpus 2 --config batch4096 --total 64 --steps 2 --warmup 2 > gpurun_out/i_batch64_n2.json 2>> gpurun_out/i_err.log
tail -n 4 gpurun_out/i_test_segments.log
grep -v "Warning\|OMP_NUM\|^\*\*\*\|NCCL version\|^$" gpurun_out/i_err.log | tail -20
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/i_*.json")):
    try:
        s=open(f).read(); d=json.loads(s[s.index('{'):])
        print(f.split('/')[-1], round(d["value"]), round(d["ms_per_step"],4), "e2e", (d.get("e2e") or {}).get("value"), d.get("parity"), d.get("segments_check"))
    except Exception as e:
        print(f, "ERR", e)
PY
