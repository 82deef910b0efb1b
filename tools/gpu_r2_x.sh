#!/bin/bash
# round-2 call X: graph replay decided at the top of the call
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/x_test.log 2>&1
echo "tests exit $?" >> gpurun_out/x_test.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/x_bench_b64.json 2>> gpurun_out/x_bench.err
tail -n 3 gpurun_out/x_test.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/x_bench*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("fraction_of_copy_only_ceiling"))
PY
tail -3 gpurun_out/x_bench.err
