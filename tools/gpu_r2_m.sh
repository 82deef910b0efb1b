#!/bin/bash
# round-2 call M: chain with next-position table, L2 eviction hints: tests, A/B, ncu
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/m_test_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/m_test_all.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
$B > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
WEFAX_L2_HINT=1 $B > gpurun_out/m_bench_hint1.json 2>> gpurun_out/m_bench.err
WEFAX_L2_HINT=2 $B > gpurun_out/m_bench_hint2.json 2>> gpurun_out/m_bench.err
WEFAX_GRAPH=0 $B > gpurun_out/m_bench_nograph.json 2>> gpurun_out/m_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/m_bench_b64.json 2>> gpurun_out/m_bench.err
WEFAX_L2_HINT=2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/m_bench_b64_hint2.json 2>> gpurun_out/m_bench.err
WEFAX_GRAPH=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -s 150 -c 80 --csv --log-file gpurun_out/m_launches_warm.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-depth 1 > gpurun_out/m_ncu_launches.log 2>&1
WEFAX_GRAPH=0 WEFAX_L2_HINT=2 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none -s 150 -c 80 --csv --log-file gpurun_out/m_launches_warm_hint2.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-depth 1 > gpurun_out/m_ncu_launches2.log 2>&1
tail -n 3 gpurun_out/m_test_all.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/m_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), "e2e", round((d.get("e2e") or {}).get("value") or 0,1), d.get("parity"), {k:round(v["ms"]*1000,1) for k,v in (d.get("stages") or {}).items()}, {k:round(v["ms"]*1000,1) for k,v in (d.get("stage_parts") or {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/m_bench.err
