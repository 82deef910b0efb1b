#!/bin/bash
# round-2 GPU call H: percentile rework (flags-first collect, fused sample+bracket, smem final), list-based chain, lanes heuristics
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_measurements.jsonl
python -m pytest tests -x -q -m gpu > gpurun_out/h_test_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/h_test_all.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
WEFAX_PCT_BRACKET=old python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_oldbracket.json 2>> gpurun_out/h_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/h_bench_b64.json 2>> gpurun_out/h_bench.err
WEFAX_LANES=2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/h_bench_b64_l2.json 2>> gpurun_out/h_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 --e2e-depth 1 > gpurun_out/h_bench_b64_d1.json 2>> gpurun_out/h_bench.err
ncu --set full --clock-control none --import-source on -k regex:'pct_|sync_chain|hilbert_mid|fft_fast_strided' -s 8 -c 8 -o gpurun_out/h_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/h_ncu.log 2>&1
tail -n 5 gpurun_out/h_test_all.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/h_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), {k:v["ms"] for k,v in d["stages"].items() if k in ("percentiles","sync_search","grey_raster","filtfilt")}, {k:v["ms"] for k,v in d["stage_parts"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/h_bench.err
