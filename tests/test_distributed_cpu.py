"""Multi-GPU host logic on CPU (gloo, world_size 2): the shard boundaries plonky_b200.distributed.ShardedMsm uses and
the product's own collective step (distributed.exchange_partials, the all-gather whose layout
plk_msm_combine_partials_dev consumes) run here on CPU tensors.  The CUDA side of the sharded path (partial + combine
kernels) is parity-tested on one GPU in tests/test_gpu_sharded.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from plonky_b200.sharding import shard_range, partial_layout
from plonky_b200 import distributed as pkd


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 1 << 20, (1 << 22) + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    limbs = 16                                              # one Tweedle XYZZ partial = 4 * 4 u64
    off, total = partial_layout(world, limbs)
    partial = torch.full((limbs,), rank + 1, dtype=torch.int64)
    gathered = torch.zeros(total, dtype=torch.int64)
    pkd.exchange_partials(partial, gathered)            # the product's collective step, gloo instead of NCCL
    ok = all(bool((gathered[off(r):off(r) + limbs] == r + 1).all()) for r in range(world))
    # a scalar vector is split exactly like the generator table
    n = 1001
    lo, hi = shard_range(n, world, rank)
    counts = torch.tensor([hi - lo], dtype=torch.int64)
    dist.all_reduce(counts)
    q.put((rank, ok and int(counts.item()) == n))
    dist.destroy_process_group()


def test_all_gather_layout_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
