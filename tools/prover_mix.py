#!/usr/bin/env python3
"""Replay of the prover's L1 call mix (BASELINE config 5 substitute).

`Circuit::generate_proof` (src/plonk.rs:84-373) on n gates issues, against ONE fixed generator table
(pedersen_g, w = 11) and the n / 8n FFT precomputations:
    group 1   9 x IFFT(n)  [wire values -> polynomials]      plonk.rs:94
              9 x FFT(8n)  [zero-padded LDE]                 plonk.rs:96
              9 x MSM(n)   [wire commitments]                plonk.rs:100
              9 x IFFT(n)                                    plonk.rs:117
    group 2   1 x IFFT(n), 1 x MSM(n)  [Z]                   plonk.rs:136-142
    group 3   1 x FFT(8n), 1 x IFFT(8n) [vanishing poly]     plonk.rs:388,455
              divide_by_z_h: coset FFT(8n) + coset IFFT(8n)  plonk.rs:170 -> polynomial.rs:330-380
              7 x MSM(n)   [t chunks]                        plonk.rs:192
    group 4   1 x MSM(n)   [PI quotient]                     plonk.rs:231
Totals: 18 MSM(n), 19 (I)FFT(n), 13 (I)FFT(8n) per proof (the remaining 8n transforms sit in the PI
quotient's polynomial arithmetic; they are replayed as plain FFT(8n)).  The full prover cannot run here
(Rust + an #[ignore]d test); this replays only the device work of the path, data resident,
with the reference's dependency groups as barriers.  Multi-GPU = replicas: the independent items of a
group are dealt round-robin to the ranks (no data-path collective); commitments are all-gathered.

  python tools/prover_mix.py [--log-n 16] [--reps 5]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/prover_mix.py
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import plonky_b200 as pk  # noqa: E402
from plonky_b200 import distributed as pkd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--ipa", action="store_true", help="also time the Halo IPA rounds (halo.rs:63-124) on rank 0")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pk._check(pk.lib().plk_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << args.log_n
    curve, field = pk.TWEEDLEDEE, pk.TWEEDLEDUM_BASE          # Circuit<Tweedledee>: scalars live in TweedledumBase
    pts = pkd.points_generate_dev(curve, 77, n)                # same generators on every rank (replicated table)
    table = pkd.msm_precompute_affine_dev(curve, pts, 11)
    plan_n = pk.fft_precompute(field, n)
    plan_8n = pk.fft_precompute(field, 8 * n)
    rng = np.random.Generator(np.random.PCG64(5))
    def rnd(rows, m):
        a = rng.integers(0, 1 << 62, size=(rows, m, 4), dtype=np.uint64)
        return torch.from_numpy(a.view(np.int64)).cuda()
    vals = rnd(9, n)
    buf_n = torch.empty_like(vals)
    buf_8n = torch.empty((9, 8 * n, 4), dtype=torch.int64, device="cuda")
    outs = torch.zeros((18, 3, 4), dtype=torch.int64, device="cuda")
    zeros = torch.zeros((18, 8), dtype=torch.uint8, device="cuda")
    zbytes = torch.zeros(32, dtype=torch.uint8, device="cuda")
    gathered = torch.zeros((world, 18, 3, 4), dtype=torch.int64, device="cuda")

    def mine(k):                       # items of a k-item group owned by this rank
        return [i for i in range(k) if i % world == rank]

    def proof():
        # group 1
        my = mine(9)
        if world == 1:
            # one batched launch sequence per kind, like values_to_polynomials / commit_polynomials
            pkd.fft_dev(plan_n, vals, buf_n, inverse=True)
            pkd.fft_dev(plan_8n, buf_n, buf_8n)
            pkd.msm_execute_batch_dev(table, buf_n, outs[:9], zbytes[:9])
            pkd.fft_dev(plan_n, vals, buf_n, inverse=True)
        else:
            for i in my:
                pkd.fft_dev(plan_n, vals[i], buf_n[i], inverse=True)
                pkd.fft_dev(plan_8n, buf_n[i], buf_8n[i])
                pkd.msm_execute_dev(table, buf_n[i], outs[i], zeros[i])
                pkd.fft_dev(plan_n, vals[i], buf_n[i], inverse=True)
        sync()
        # group 2
        if rank == 0:
            pkd.fft_dev(plan_n, vals[0], buf_n[0], inverse=True)
            pkd.msm_execute_dev(table, buf_n[0], outs[9], zeros[9])
        sync()
        # group 3
        if rank == 0:
            pkd.fft_dev(plan_8n, buf_n[0], buf_8n[0])
            pkd.fft_dev(plan_8n, buf_8n[0], buf_8n[1], inverse=True)
            pkd.fft_dev(plan_8n, buf_8n[1], buf_8n[2], coset=True)
            pkd.fft_dev(plan_8n, buf_8n[2], buf_8n[3], inverse=True, coset=True)
        sync()
        if world == 1:
            pkd.msm_execute_batch_dev(table, buf_8n[3, :7 * n].view(7, n, 4), outs[10:17], zbytes[10:17])
        else:
            for i in mine(7):
                pkd.msm_execute_dev(table, buf_8n[3, i * n:(i + 1) * n], outs[10 + i], zeros[10 + i])
        for i in mine(8):              # the remaining 8n transforms of the quotient arithmetic
            pkd.fft_dev(plan_8n, buf_8n[i], buf_8n[(i + 1) % 9], inverse=bool(i & 1))
        sync()
        # group 4
        if rank == 0:
            pkd.msm_execute_dev(table, buf_n[1], outs[17], zeros[17])
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), outs.view(-1))

    def sync():
        if world > 1:
            dist.barrier()

    for _ in range(2):
        proof()
    torch.cuda.synchronize()
    sync()
    l0 = pk.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        proof()
    e1.record()
    torch.cuda.synchronize()
    sync()
    t = torch.tensor([e0.elapsed_time(e1) / args.reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ipa = None
    if args.ipa and rank == 0:
        # batch_opening_proof's loop (halo.rs:63-124): log2(n) rounds of 2 variable-base MSMs + 2 inner products,
        # then the fold of a, b and G; vectors stay on the device, one challenge per round comes from the host
        import time
        a = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
        b = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
        g = pk.points_generate(curve, 77, n)
        u = rng.integers(1, 1 << 62, size=4, dtype=np.uint64)
        u_inv = pk.field_op(field, "inverse", u.reshape(1, 4))[0]
        host_table = pk.msm_precompute_affine(curve, g, 11)

        def run(table_mode):
            best = None
            for _ in range(3):
                st = pk.HaloIpaRounds(curve, a, b, precomputation=host_table) if table_mode else pk.HaloIpaRounds(curve, a, b, g)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                while len(st) > 1:
                    st.round_lr()
                    st.fold(u, u_inv)
                st.read()                         # halo_a[0], halo_b[0], halo_g[0].to_affine()
                dt = (time.perf_counter() - t0) * 1e3
                best = dt if best is None else min(best, dt)
                st.close()
            return best

        ipa = {"rounds": args.log_n, "ms_all_rounds_table_mode": run(True), "ms_all_rounds_folding_mode": run(False),
               "timing": "host wall clock around the synchronous C-ABI calls (2 per round + final read), best of 3"}
    if rank == 0:
        print(json.dumps({"ipa": ipa, "workload": f"prover L1 call mix, n = 2^{args.log_n} gates (18 MSM(n), 19 FFT(n), 13 FFT(8n))",
                          "n_gpus": world, "ms_per_proof_mix": float(t.item()), "mode": "replicas, round-robin within dependency groups",
                          "launches_per_proof_rank0": (pk.kernel_launch_count() - l0) // args.reps}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
