#!/bin/bash
# round-2 call O: three-tile TMA pass
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/o_test_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/o_test_all.log
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
$B > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err
WEFAX_TMA_PIPE=0 $B > gpurun_out/o_bench_pipe0.json 2>> gpurun_out/o_bench.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/o_bench_b64.json 2>> gpurun_out/o_bench.err
WEFAX_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_fast_tma3 -s 3 -c 1 -o gpurun_out/o_prof_tma3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-depth 1 > gpurun_out/o_ncu.log 2>&1
tail -n 3 gpurun_out/o_test_all.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/o_b*.json")):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), "e2e", round((d.get("e2e") or {}).get("value") or 0,1), d.get("parity"), {k:round(v["ms"]*1000,1) for k,v in (d.get("stages") or {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/o_bench.err
