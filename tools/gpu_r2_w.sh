#!/bin/bash
# round-2 call W: the bench lines of the final build that go into profiles/
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err
python bench.py --config batch4096 --total 512 --steps 3 --warmup 2 > gpurun_out/w_batch512_n1.json 2>> gpurun_out/w_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --duration 600 --batch 64 > gpurun_out/w_bench_b64.json 2>> gpurun_out/w_bench.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --sample-rate 48000 --duration 1200 > gpurun_out/w_bench_48k.json 2>> gpurun_out/w_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/w_*.json")):
    d=json.load(open(f)); print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("fraction_of_copy_only_ceiling"), d.get("parity"))
PY
tail -3 gpurun_out/w_bench.err
