#!/bin/bash
# round-2 last call: the frozen build - all GPU tests, smoke, both bench arms
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_measurements.jsonl
python -m pytest tests -x -q -m gpu > gpurun_out/last_test_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/last_test_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/last_smoke.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/last_bench_reference.json 2> gpurun_out/last_bench.err
python bench.py > gpurun_out/last_bench.json 2>> gpurun_out/last_bench.err
tail -n 3 gpurun_out/last_test_all.log gpurun_out/last_smoke.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/last_bench*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("fraction_of_copy_only_ceiling"), d.get("parity"), d.get("cpu_baseline",{}).get("kind"), (d.get("roofline") or {}).get("frac"))
PY
tail -3 gpurun_out/last_bench.err
