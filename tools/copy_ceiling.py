#!/usr/bin/env python
"""What the HOST can move: N processes (one per GPU, launched by torchrun or alone), each copying the end-to-end
traffic of one decode step - 79 MB of pinned PCM up, 198 MB of pinned results down, full duplex on two streams -
with NO kernels in between.  This is the ceiling the end-to-end (`e2e`) figure of bench.py can reach on this box.

    python tools/copy_ceiling.py                                             # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 tools/copy_ceiling.py # 8 GPUs at once
Prints one JSON line on rank 0: aggregate GB/s each way and the equivalent decode rate in Msamples/s."""
import json
import os
import time

import torch


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n = 39_690_000                                   # samples of the 60-min recording
    up_bytes, down_bytes = 2 * n, n + 4 * (n // 5512) * 5512
    h_up = torch.empty(up_bytes, dtype=torch.uint8, pin_memory=True)
    h_dn = torch.empty(down_bytes, dtype=torch.uint8, pin_memory=True)
    d_up = torch.empty(up_bytes, dtype=torch.uint8, device="cuda")
    d_dn = torch.empty(down_bytes, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    steps = 30

    def run(do_up, do_dn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            if do_up:
                with torch.cuda.stream(s_up):
                    d_up.copy_(h_up, non_blocking=True)
            if do_dn:
                with torch.cuda.stream(s_dn):
                    h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    run(True, True)
    t_up, t_dn, t_both = run(True, False), run(False, True), run(True, True)
    if rank == 0:
        print(json.dumps({
            "n_gpus": world, "h2d_bytes_per_step": up_bytes, "d2h_bytes_per_step": down_bytes,
            "h2d_only_gbs": world * steps * up_bytes / t_up / 1e9,
            "d2h_only_gbs": world * steps * down_bytes / t_dn / 1e9,
            "duplex_h2d_gbs": world * steps * up_bytes / t_both / 1e9,
            "duplex_d2h_gbs": world * steps * down_bytes / t_both / 1e9,
            "e2e_ceiling_msamples_s": world * steps * n / t_both / 1e6,
            "note": "pinned host buffers, one process per GPU, no kernels: the end-to-end decode rate cannot exceed "
                    "e2e_ceiling_msamples_s on this host"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
