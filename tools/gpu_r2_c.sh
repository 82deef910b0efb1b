#!/bin/bash
# round-2 GPU call C: grey_raster v3 (4 columns, rolled loop), lean percentile collect, two-level sync chain
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/c_test_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/c_test_all.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
for u in 8 12 16 20 28; do
WEFAX_GR_LINES=$u python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_u$u.json 2>> gpurun_out/c_bench.err
done
WEFAX_PCT_COLLECT=old python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_oldcollect.json 2>> gpurun_out/c_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/c_bench_b64.json 2>> gpurun_out/c_bench.err
ncu --set full --clock-control none --import-source on -k regex:'grey_raster|pct_collect2|sync_chain' -s 6 -c 3 -o gpurun_out/c_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c_ncu.log 2>&1
tail -n 3 gpurun_out/c_test_all.log
python - <<'PY'
import json
for f in ("c_bench","c_bench_u8","c_bench_u12","c_bench_u16","c_bench_u20","c_bench_u28","c_bench_oldcollect","c_bench_b64"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"]), round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()}, {k:v["ms"] for k,v in d["stage_parts"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
