#!/bin/bash
# round-2 call K: bitmap chain, always-median collect, warp-autonomous middle kernel, TMA envelope pass: tests, A/B, ncu
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/k_test_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/k_test_all.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
$B > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
WEFAX_MID_WARP=0 $B > gpurun_out/k_bench_midcta.json 2>> gpurun_out/k_bench.err
WEFAX_TMA_ENV=0 $B > gpurun_out/k_bench_envdirect.json 2>> gpurun_out/k_bench.err
WEFAX_PCT_COLLECT=2 $B > gpurun_out/k_bench_collect2.json 2>> gpurun_out/k_bench.err
WEFAX_PCT_NCTA=64 $B > gpurun_out/k_bench_ncta64.json 2>> gpurun_out/k_bench.err
WEFAX_PCT_NCTA=128 $B > gpurun_out/k_bench_ncta128.json 2>> gpurun_out/k_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/k_bench_b64.json 2>> gpurun_out/k_bench.err
WEFAX_GRAPH=0 ncu --set full --clock-control none --import-source on -s 40 -c 20 -o gpurun_out/k_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-depth 1 > gpurun_out/k_ncu.log 2>&1
tail -n 3 gpurun_out/k_test_all.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/k_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), "e2e", round((d.get("e2e") or {}).get("value") or 0,1), d.get("parity"), {k:round(v["ms"]*1000,1) for k,v in (d.get("stages") or {}).items()}, {k:round(v["ms"]*1000,1) for k,v in (d.get("stage_parts") or {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/k_bench.err
