#!/bin/bash
# the default bench at 2 GPUs: affinity only where pinned memory is allocated (compare 0.770 ms bound / 0.738 ms never bound)
set -x
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/ab_bench_n2_rebind.json 2> gpurun_out/ab_err.log
python -c "
import json; d=json.load(open('gpurun_out/ab_bench_n2_rebind.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('fraction_of_copy_only_ceiling'), d.get('segments_check',{}).get('start_frame_equal'))"
