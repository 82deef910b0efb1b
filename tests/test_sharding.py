"""The N > 1 path on CPU: partitioning of recordings over ranks and the host gather,
with world_size 2 over gloo.  The per-rank decode is replaced by the oracle here (no GPU
in this container); on the GPU box the same functions drive Decoder.decode."""
import hashlib
import os
import socket

import numpy as np
import pytest

from wefax_b200 import sharding, synth


def test_assign_is_a_balanced_partition():
    keys = [(6_615_000, 11025, 1)] * 10 + [(330750, 11025, 1)] * 7 + [(480000, 48000, 2)] * 3
    for world in (1, 2, 3, 4, 8):
        shards = sharding.assign(keys, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(keys)))
        sizes = [len(s) for s in shards]
        assert max(sizes) - min(sizes) <= 1
        # every bucket is spread, not dumped on one rank
        for k in set(keys):
            per_rank = [sum(1 for i in s if keys[i] == k) for s in shards]
            assert max(per_rank) - min(per_rank) <= 1
    with pytest.raises(ValueError):
        sharding.assign(keys, 0)


def test_local_batches_group_by_key():
    keys = [("a",), ("b",), ("a",), ("a",), ("b",), ("c",)]
    got = sharding.local_batches(keys, 0, 2) + sharding.local_batches(keys, 1, 2)
    seen = sorted(i for _, idx in got for i in idx)
    assert seen == list(range(6))
    for k, idx in got:
        assert all(keys[i] == k for i in idx)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _specs():
    return [dict(duration_s=6.0 + 2.0 * (k % 2), lpm=(120, 240)[k % 2], seed=100 + k, noise_sigma=0.02) for k in range(5)]


def _decode_with_oracle(spec):
    from oracle import wefax_oracle as O
    pcm = synth.synth_recording(**spec)
    o = O.decode(pcm, 11025, spec["lpm"])
    img = o.get("output_image")
    return dict(start_frame=o.get("start_frame"), error=o["error"], n=int(pcm.shape[0]),
                image_sha=hashlib.sha256(img.tobytes()).hexdigest() if img is not None else None)


def _worker(rank, world, port, queue):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    specs = _specs()
    keys = [(int(round(s["duration_s"] * 11025)), 11025, 1) for s in specs]

    def decode_batch(key, idx):
        return {i: _decode_with_oracle(specs[i]) for i in idx}

    local = sharding.decode_sharded(keys, decode_batch, rank, world)
    merged = sharding.gather_results(local, dst=0)
    dist.barrier()
    if rank == 0:
        queue.put(merged)
    else:
        assert merged is None
    dist.destroy_process_group()


def test_two_ranks_over_gloo_equal_single_rank():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    merged = queue.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    specs = _specs()
    assert sorted(merged) == list(range(len(specs)))
    for i, spec in enumerate(specs):
        assert merged[i] == _decode_with_oracle(spec)
