"""torchrun helper (GPU box, N >= 2 GPUs): the domain-split NTT over NCCL equals the single-GPU transform.
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_ntt_check.py [log_n]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonky_b200 as pk  # noqa: E402
from plonky_b200.distributed import DistributedNtt, fft_dev  # noqa: E402


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    pk._check(pk.lib().plk_set_device(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world, rank = dist.get_world_size(), dist.get_rank()
    n = 1 << log_n
    rng = np.random.Generator(np.random.PCG64(1234))          # same data on every rank
    a = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    x = torch.from_numpy(a.view(np.int64)).cuda()
    mode = os.environ.get("PLK_DIST_NTT_EXCHANGE", "p2p")
    d = DistributedNtt(pk.TWEEDLEDEE_BASE, log_n, p2p=(False if mode == "nccl" else True))
    if rank == 0:
        print("exchange:", d.exchange_kind)
    rows = d.input_rows(x)
    for inverse in (False, True, False, True):          # twice: both receive buffers of the peer-store path
        out = d.forward(rows, inverse=inverse).clone()
        plan = pk.fft_precompute(pk.TWEEDLEDEE_BASE, n)
        want = torch.empty_like(x)
        fft_dev(plan, x, want, inverse=inverse)
        # this rank's slice of the natural-order result: X[k' + M k_1], k' in its column block
        M = 1 << d.log_m
        mine = want.view(1 << d.log_r1, M, 4)[:, rank * d.cols:(rank + 1) * d.cols]
        ok = torch.equal(out, mine.contiguous())
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"domain-split NTT 2^{log_n} over {world} GPUs inverse={inverse}: {'bit-exact' if flag.item() else 'MISMATCH'}")
        assert flag.item() == 1
    # timing (max over ranks), data resident
    for _ in range(3):
        d.forward(rows)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 10
    for _ in range(K):
        d.forward(rows)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / K], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"domain-split NTT 2^{log_n} over {world} GPUs: {t.item():.3f} ms/transform = {n / (t.item() * 1e-3):.4g} elements/s")
    d.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
