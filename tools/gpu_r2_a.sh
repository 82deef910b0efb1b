#!/bin/bash
# round-2 GPU call A: new fused / notch kernels: tests, A/B bench, ncu
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
python -m pytest tests/test_gpu_fused.py -x -q -m gpu > gpurun_out/a_test_fused.log 2>&1
echo "fused tests exit $?" >> gpurun_out/a_test_fused.log
python -m pytest tests -x -q -m gpu > gpurun_out/a_test_all.log 2>&1
echo "all tests exit $?" >> gpurun_out/a_test_all.log
python bench.py --steps 20 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
WEFAX_FUSED=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_nofused.json 2>> gpurun_out/a_bench.err
WEFAX_NOTCH_SYM=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_nosym.json 2>> gpurun_out/a_bench.err
for u in 12 20 32; do
WEFAX_GR_LINES=$u python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench_u$u.json 2>> gpurun_out/a_bench.err
done
ncu --set full --clock-control none --import-source on -k regex:'grey_raster|notch_sym' -s 4 -c 2 -o gpurun_out/a_prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/a_ncu2.log 2>&1
tail -3 gpurun_out/a_test_fused.log gpurun_out/a_test_all.log
python - <<'PY'
import json
for f in ("a_bench","a_bench_nofused","a_bench_nosym","a_bench_u12","a_bench_u20","a_bench_u32"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"]), round(d["ms_per_step"],4), {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
