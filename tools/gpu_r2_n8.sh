#!/bin/bash
# round-2 multi-GPU call: configs[3]/[4] for real, copy ceiling, scaling of the default bench, segments
set -x
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi topo -m > gpurun_out/n8_topo.txt 2>&1
lscpu | head -30 > gpurun_out/n8_lscpu.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# copy-only ceiling of the host at 1 and N GPUs
python tools/copy_ceiling.py > gpurun_out/n8_copy_1.json 2> gpurun_out/n8_err.log
$TR --nproc-per-node $N --master-port 29511 tools/copy_ceiling.py > gpurun_out/n8_copy_$N.json 2>> gpurun_out/n8_err.log
# configs[3]: 512 recordings per GPU (4096 over 8), and the same share on one GPU
python bench.py --config batch4096 --total 512 --steps 3 --warmup 2 > gpurun_out/n8_batch512_n1.json 2>> gpurun_out/n8_err.log
$TR --nproc-per-node $N --master-port 29512 bench.py --gpus $N --config batch4096 --total $((512*N)) --steps 3 --warmup 2 > gpurun_out/n8_batch4096_n$N.json 2>> gpurun_out/n8_err.log
$TR --nproc-per-node $N --master-port 29513 bench.py --gpus $N --config noisy4096 --total $((512*N)) --steps 3 --warmup 2 > gpurun_out/n8_noisy4096_n$N.json 2>> gpurun_out/n8_err.log
# the default bench at N GPUs (segments_check included) and at 2
$TR --nproc-per-node $N --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/n8_bench_n$N.json 2>> gpurun_out/n8_err.log
$TR --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n8_bench_n2.json 2>> gpurun_out/n8_err.log
# configs[2]: one 20-min 48 kHz recording in segments over N GPUs, 2 GPUs, 1 GPU
python bench.py --segments --steps 10 --warmup 3 > gpurun_out/n8_segments_n1.json 2>> gpurun_out/n8_err.log
$TR --nproc-per-node 2 --master-port 29516 bench.py --gpus 2 --segments --steps 10 --warmup 3 > gpurun_out/n8_segments_n2.json 2>> gpurun_out/n8_err.log
$TR --nproc-per-node $N --master-port 29517 bench.py --gpus $N --segments --steps 10 --warmup 3 > gpurun_out/n8_segments_n$N.json 2>> gpurun_out/n8_err.log
# the one GPU test that needs two GPUs
python -m pytest tests/test_gpu_segments.py -x -q -m gpu > gpurun_out/n8_test_segments.log 2>&1
tail -n 3 gpurun_out/n8_test_segments.log
tail -n 20 gpurun_out/n8_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n8_*.json")):
    try:
        d=json.load(open(f))
        keys=[k for k in ("value","ms_per_step","e2e_ceiling_msamples_s","duplex_d2h_gbs","d2h_only_gbs","h2d_only_gbs","parity","segments_check") if k in d]
        print(f.split('/')[-1], {k:d[k] for k in keys}, "e2e", (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
