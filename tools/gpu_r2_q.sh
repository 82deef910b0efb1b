#!/bin/bash
# round-2 call Q: the fused sweep per LPM width class (why is the mixed batch's raster 2x slower per sample?)
set -x
mkdir -p gpurun_out
for L in 60 90 100 120 180 240; do
  python bench.py --steps 8 --warmup 4 --no-cpu-baseline --e2e-depth 1 --lpm $L > gpurun_out/q_bench_lpm$L.json 2>> gpurun_out/q_bench.err
done
WEFAX_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:grey_raster -s 2 -c 1 -o gpurun_out/q_prof_raster60 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-depth 1 --lpm 60 > gpurun_out/q_ncu.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/q_bench*.json")):
    try:
        d=json.load(open(f))
        print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), {k:round(v["ms"]*1000,1) for k,v in (d.get("stages") or {}).items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 gpurun_out/q_bench.err
