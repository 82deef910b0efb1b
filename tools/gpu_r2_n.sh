#!/bin/bash
# round-2 call N: ring-buffer raster window (compare with 94.8 us of call M), batch numbers with the new kernels
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/n_test.log 2>&1
echo "tests exit $?" >> gpurun_out/n_test.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
$B > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
python bench.py --config batch4096 --total 512 --steps 3 --warmup 2 > gpurun_out/n_batch512_n1.json 2>> gpurun_out/n_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/n_bench_b64.json 2>> gpurun_out/n_bench.err
WEFAX_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:grey_raster -s 2 -c 1 -o gpurun_out/n_prof_raster python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-depth 1 > gpurun_out/n_ncu.log 2>&1
tail -n 3 gpurun_out/n_test.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n_b*.json")):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), "e2e", round((d.get("e2e") or {}).get("value") or 0,1), d.get("parity"), {k:round(v["ms"]*1000,1) for k,v in (d.get("stages") or {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/n_bench.err
