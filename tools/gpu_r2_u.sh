#!/bin/bash
# round-2 call U: fused sweep on widths that are not multiples of 8 (whole last column group, inline stores, 128 registers)
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/u_test.log 2>&1
echo "tests exit $?" >> gpurun_out/u_test.log
for L in 60 90 100 120 180 240; do
  python bench.py --steps 8 --warmup 4 --no-cpu-baseline --e2e-depth 1 --lpm $L > gpurun_out/u_bench_lpm$L.json 2>> gpurun_out/u_bench.err
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/u_bench_b64.json 2>> gpurun_out/u_bench.err
WEFAX_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:grey_raster -s 2 -c 1 -o gpurun_out/u_prof_raster60 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-depth 1 --lpm 60 > gpurun_out/u_ncu.log 2>&1
tail -n 3 gpurun_out/u_test.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/u_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), {k:round(v["ms"]*1000,1) for k,v in (d.get("stages") or {}).items() if k in ("grey_raster","percentiles","sync_search")})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/u_bench.err
