#!/bin/bash
# round-2 call L: tight chain loop, side stream, new defaults: tests, A/B, ncu
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/l_test_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/l_test_all.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
$B > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
WEFAX_SIDE=0 $B > gpurun_out/l_bench_noside.json 2>> gpurun_out/l_bench.err
WEFAX_GRAPH=0 $B > gpurun_out/l_bench_nograph.json 2>> gpurun_out/l_bench.err
WEFAX_GRAPH=0 WEFAX_SIDE=0 $B > gpurun_out/l_bench_nograph_noside.json 2>> gpurun_out/l_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/l_bench_b64.json 2>> gpurun_out/l_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sample-rate 48000 --duration 1200 > gpurun_out/l_bench_48k.json 2>> gpurun_out/l_bench.err
WEFAX_GRAPH=0 WEFAX_SIDE=0 ncu --set full --clock-control none --import-source on -s 40 -c 20 -o gpurun_out/l_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-depth 1 > gpurun_out/l_ncu.log 2>&1
tail -n 3 gpurun_out/l_test_all.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/l_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), "e2e", round((d.get("e2e") or {}).get("value") or 0,1), d.get("parity"), {k:round(v["ms"]*1000,1) for k,v in (d.get("stages") or {}).items()}, {k:round(v["ms"]*1000,1) for k,v in (d.get("stage_parts") or {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/l_bench.err
