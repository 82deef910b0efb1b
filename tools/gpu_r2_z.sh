#!/bin/bash
# round-2 call Z: settled-bits kernel thread count in batches (A/B), batch-64
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batch or noisy or waves" > gpurun_out/z_test.log 2>&1
echo "tests exit $?" >> gpurun_out/z_test.log
for T in 1024 512; do
WEFAX_SETTLED_THREADS=$T python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-depth 1 --duration 600 --batch 64 > gpurun_out/z_bench_b64_t$T.json 2>> gpurun_out/z_bench.err
done
tail -n 3 gpurun_out/z_test.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/z_bench*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), {k:round(v["ms"]*1000,1) for k,v in (d.get("stage_parts") or {}).items()})
PY
tail -3 gpurun_out/z_bench.err
