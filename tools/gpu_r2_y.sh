#!/bin/bash
# round-2 call Y: settled-bits kernel with 256 threads in batches
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/y_test_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/y_test_all.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/y_bench_b64.json 2>> gpurun_out/y_bench.err
tail -n 3 gpurun_out/y_test_all.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/y_bench*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), {k:round(v["ms"]*1000,1) for k,v in (d.get("stage_parts") or {}).items()}, round(d["stages"]["sync_search"]["ms"]*1000,1))
PY
tail -3 gpurun_out/y_bench.err
