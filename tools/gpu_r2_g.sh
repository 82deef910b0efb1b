#!/bin/bash
# round-2 GPU call G: depth-first lanes for batches; remaining tests
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tones.py tests/test_gpu_fm.py tests/test_gpu_segments.py tests/test_gpu_fused.py -x -q -m gpu > gpurun_out/g_test_rest.log 2>&1
echo "rest tests exit $?" >> gpurun_out/g_test_rest.log
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "batch or waves or device_resident" > gpurun_out/g_test_batch.log 2>&1
echo "batch tests exit $?" >> gpurun_out/g_test_batch.log
for cfg in "0 2 1" "1 2 1" "1 2 2" "1 3 1" "1 4 1" "1 1 1"; do
set -- $cfg
WEFAX_DEPTH_FIRST=$1 WEFAX_LANES=$2 WEFAX_LANE_WAVE=$3 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/g_bench_b64_df$1_l$2_w$3.json 2>> gpurun_out/g_bench.err
done
WEFAX_DEPTH_FIRST=1 WEFAX_LANES=2 WEFAX_LANE_WAVE=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --duration 600 --batch 512 > gpurun_out/g_bench_b512_df1.json 2>> gpurun_out/g_bench.err
tail -n 5 gpurun_out/g_test_rest.log gpurun_out/g_test_batch.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/g_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/g_bench.err
