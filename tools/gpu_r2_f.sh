#!/bin/bash
# round-2 GPU call F: sync pulse scan (N2), recursive notch, table-driven settled-bit chain
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_measurements.jsonl
python -m pytest tests -x -q -m gpu > gpurun_out/f_test_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/f_test_all.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
tail -n 30 gpurun_out/f_test_all.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/f_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], round(d["value"]), round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, {k:v["ms"] for k,v in d["stage_parts"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
