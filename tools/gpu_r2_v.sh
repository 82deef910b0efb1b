#!/bin/bash
# round-2 call V: short way for repeated decode(out=res) calls (host overhead between graph replays)
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/v_test_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/v_test_all.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/v_bench_200.json 2>> gpurun_out/v_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/v_bench_b64.json 2>> gpurun_out/v_bench.err
tail -n 3 gpurun_out/v_test_all.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/v_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), "e2e", round((d.get("e2e") or {}).get("value") or 0,1), (d.get("e2e") or {}).get("fraction_of_copy_only_ceiling"))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/v_bench.err
