#!/bin/bash
# round-2 GPU call B: grey_raster v2 (templated offsets, class launches), persistent notch
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py -x -q -m gpu > gpurun_out/b_test_fused.log 2>&1
echo "fused tests exit $?" >> gpurun_out/b_test_fused.log
python -m pytest tests -x -q -m gpu > gpurun_out/b_test_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/b_test_all.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
for u in 10 14 18 24; do
WEFAX_GR_LINES=$u python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_u$u.json 2>> gpurun_out/b_bench.err
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/b_bench_b64.json 2>> gpurun_out/b_bench.err
ncu --set full --clock-control none --import-source on -k regex:'grey_raster|notch_sym' -s 4 -c 2 -o gpurun_out/b_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
tail -n 3 gpurun_out/b_test_fused.log gpurun_out/b_test_all.log
python - <<'PY'
import json
for f in ("b_bench","b_bench_u10","b_bench_u14","b_bench_u18","b_bench_u24","b_bench_b64"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"]), round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
