#!/bin/bash
# round-2 call P (final state): full tests, smoke, bench (both arms), launch list, ncu full step, sanitizers
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_measurements.jsonl
python -m pytest tests -x -q -m gpu > gpurun_out/p_test_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/p_test_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/p_bench_reference.json 2>> gpurun_out/p_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/p_bench_b64.json 2>> gpurun_out/p_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sample-rate 48000 --duration 1200 > gpurun_out/p_bench_48k.json 2>> gpurun_out/p_bench.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-depth 1 > gpurun_out/p_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -s 40 -c 20 -o gpurun_out/p_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-depth 1 > gpurun_out/p_ncu.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitizer_smoke.py > gpurun_out/p_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/p_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitizer_smoke.py > gpurun_out/p_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/p_racecheck.log
tail -n 3 gpurun_out/p_test_all.log gpurun_out/p_smoke.log gpurun_out/p_memcheck.log gpurun_out/p_racecheck.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/p_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), "e2e", (d.get("e2e") or {}).get("value"), d.get("parity"), {k:v["ms"] for k,v in (d.get("stages") or {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/p_bench.err
