#!/bin/bash
# round-2 GPU call E: grey_raster v5 (8 columns, FFMA2 raster)
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_measurements.jsonl
python -m pytest tests -x -q -m gpu > gpurun_out/e_test_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/e_test_all.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
for u in 10 13 16 20 26 32; do
WEFAX_GR_LINES=$u python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_u$u.json 2>> gpurun_out/e_bench.err
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/e_bench_b64.json 2>> gpurun_out/e_bench.err
ncu --set full --clock-control none --import-source on -k regex:'grey_raster' -s 2 -c 1 -o gpurun_out/e_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/e_ncu.log 2>&1
tail -n 3 gpurun_out/e_test_all.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/e_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], round(d["value"]), round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items() if k in ("grey_raster","filtfilt","percentiles","sync_search")})
    except Exception as e:
        print(f, "ERR", e)
PY
