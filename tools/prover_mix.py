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
              vanishing_points over the 8n LDE points        plonk.rs:393-452 (ten gate evaluators, gates/mod.rs:46-125)
              divide_by_z_h: coset FFT(8n) + coset IFFT(8n)  plonk.rs:170 -> polynomial.rs:330-380
              7 x MSM(n)   [t chunks]                        plonk.rs:192
    group 4   1 x MSM(n)   [PI quotient]                     plonk.rs:231
Totals: 18 MSM(n), 19 (I)FFT(n), 13 (I)FFT(8n) per proof (the remaining 8n transforms sit in the PI
quotient's polynomial arithmetic; they are replayed as plain FFT(8n)).  The full prover cannot run here
(Rust + an #[ignore]d test); this replays only the device work of the path with the
reference's dependency groups as barriers: the wire values arrive from HOST memory (H2D inside the timed
region) and every group ends with the D2H read of its commitments (what the Fiat-Shamir challenger needs
before the next group can start).  Multi-GPU = replicas: the independent items of a group are dealt
round-robin to the ranks (no data-path collective); commitments are all-gathered.

  python tools/prover_mix.py [--log-n 16] [--reps 5]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/prover_mix.py
bench.py imports `bench()` for its `prover_mix` sub-object.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Mix:
    def __init__(self, cx, log_n):
        torch, np, pk, pkd = cx.torch, cx.np, cx.pk, cx.pkd
        self.cx = cx
        self.n = n = 1 << log_n
        self.world, self.rank = cx.world, cx.rank
        self.curve, self.field = pk.TWEEDLEDEE, pk.TWEEDLEDUM_BASE          # Circuit<Tweedledee>: scalars live in TweedledumBase
        self.pts = pkd.pedersen_generators_dev(self.curve, 0, n)             # same generators on every rank (replicated table)
        self.table = pkd.msm_precompute_affine_dev(self.curve, self.pts, 11)
        self.plan_n = pk.fft_precompute(self.field, n)
        self.plan_8n = pk.fft_precompute(self.field, 8 * n)
        rng = np.random.Generator(np.random.PCG64(5))
        a = rng.integers(0, 1 << 62, size=(9, n, 4), dtype=np.uint64)
        self.vals_h = torch.from_numpy(a.view(np.int64)).pin_memory()        # Witness wire values, host
        self.vals = torch.empty((9, n, 4), dtype=torch.int64, device="cuda")
        self.buf_n = torch.zeros_like(self.vals)
        self.buf_8n = torch.zeros((9, 8 * n, 4), dtype=torch.int64, device="cuda")
        self.outs = torch.zeros((18, 3, 4), dtype=torch.int64, device="cuda")
        self.zeros = torch.zeros((18, 8), dtype=torch.uint8, device="cuda")
        self.zbytes = torch.zeros(32, dtype=torch.uint8, device="cuda")
        self.gathered = torch.zeros((self.world, 18, 3, 4), dtype=torch.int64, device="cuda")
        self.outs_h = torch.zeros((18, 3, 4), dtype=torch.int64).pin_memory()
        self.h2d_bytes = self.vals_h.numel() * 8
        self.buf_n2 = torch.zeros_like(self.vals)                    # target of group 1's second IFFT batch
        self.side = torch.cuda.Stream()                              # independent items of a dependency group overlap on two streams
        # inputs of the pointwise vanishing evaluation that a Circuit holds precomputed: constants_8n, s_sigma_values_8n,
        # subgroup_8n (plonk.rs:47-63); wires_8n / Z(8n) are the LDE outputs of groups 1 and 3
        b = rng.integers(0, 1 << 62, size=(12, 8 * n, 4), dtype=np.uint64)
        self.consts_8n = torch.from_numpy(b[:6].view(np.int64)).cuda()
        self.sigma_8n = torch.from_numpy(b[6:].view(np.int64)).cuda()
        self.params = torch.from_numpy(rng.integers(0, 1 << 62, size=(11, 4), dtype=np.uint64).view(np.int64)).cuda()   # k_is[6], alpha, beta, gamma, zeta, a
        self.subgroup_8n = torch.from_numpy(pk.fft_subgroup(self.plan_8n).view(np.int64)).cuda()
        self.van = torch.zeros((8 * n, 4), dtype=torch.int64, device="cuda")
        self.scr_8n = torch.zeros((5, 8 * n, 4), dtype=torch.int64, device="cuda")        # operands of the quotient arithmetic's 8n transforms
        self.wires_8n = torch.zeros((9, 8 * n, 4), dtype=torch.int64, device="cuda")      # every rank's copy of the nine wire LDEs

    def mine(self, k):
        """this rank's contiguous block [lo, hi) of a k-item group (the split ShardedMsm uses, plonky_b200/sharding.py)"""
        from plonky_b200.sharding import shard_range
        return shard_range(k, self.world, self.rank)

    def owner(self, k, i):
        from plonky_b200.sharding import shard_range
        for r in range(self.world):
            lo, hi = shard_range(k, self.world, r)
            if lo <= i < hi:
                return r
        raise IndexError(i)

    def group_end(self, lo, hi):
        """dependency barrier: the group's commitments reach the host (rank 0 holds all of them for N > 1)"""
        cx = self.cx
        self.marks.append(time.perf_counter())
        if self.world > 1:
            cx.dist.all_gather_into_tensor(self.gathered.view(-1), self.outs.view(-1))
        self.outs_h[lo:hi].copy_(self.outs[lo:hi], non_blocking=True)
        cx.torch.cuda.current_stream().synchronize()
        self.marks.append(time.perf_counter())

    def proof(self, single=False):
        """one proof's device work; single=True forces the one-GPU batched path (the N > 1 self-check)"""
        pkd = self.cx.pkd
        n = self.n
        world = 1 if single else self.world
        t, pn, p8 = self.table, self.plan_n, self.plan_8n
        self.marks = [time.perf_counter()]
        self.vals.copy_(self.vals_h, non_blocking=True)
        # group 1
        torch = self.cx.torch
        main = torch.cuda.current_stream()
        if world == 1:
            # one batched launch sequence per kind, like values_to_polynomials / commit_polynomials; the LDE transforms and the
            # commitments both depend only on the coefficients, so they run on two streams
            pkd.fft_dev(pn, self.vals, self.buf_n, inverse=True)
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                pkd.fft_dev(p8, self.buf_n, self.buf_8n)
                pkd.fft_dev(pn, self.vals, self.buf_n2, inverse=True)
            pkd.msm_execute_batch_dev(t, self.buf_n, self.outs[:9], self.zbytes[:9])
            main.wait_stream(self.side)
        else:
            lo, hi = self.mine(9)
            if hi > lo:                   # this rank's block of the nine wires, batched like the single-GPU path
                pkd.fft_dev(pn, self.vals[lo:hi], self.buf_n[lo:hi], inverse=True)
                pkd.fft_dev(p8, self.buf_n[lo:hi], self.buf_8n[lo:hi])
                pkd.msm_execute_batch_dev(t, self.buf_n[lo:hi], self.outs[lo:hi], self.zbytes[lo:hi])
                pkd.fft_dev(pn, self.vals[lo:hi], self.buf_n[lo:hi], inverse=True)
            # every replica needs all nine wire LDEs for the vanishing polynomial: the blocks are disjoint, so a sum over
            # the ranks of zero-filled copies is their concatenation (one 9 x 8n x 32 B all-reduce per proof)
            self.wires_8n.zero_()
            if hi > lo:
                self.wires_8n[lo:hi].copy_(self.buf_8n[lo:hi])
            self.cx.dist.all_reduce(self.wires_8n)
        if not single:
            self.group_end(0, 9)
        # group 2 (Z): every rank keeps its own copy of the chain input so that group 3 can be dealt out
        pkd.fft_dev(pn, self.vals[0], self.buf_n[0], inverse=True)
        if world == 1 or self.rank == 0:
            pkd.msm_execute_dev(t, self.buf_n[0], self.outs[9], self.zeros[9])
        if not single:
            self.group_end(9, 10)
        # group 3: FFT(8n), IFFT(8n), divide_by_z_h's coset pair -- a chain, replicated on every rank
        pkd.fft_dev(p8, self.buf_n[0], self.buf_8n[0])
        if world == 1:
            # t chunks: dense scalars (the 8n evaluations of the chain's first transform stand in for the quotient's chunks);
            # the seven commitments overlap with the vanishing evaluation and the rest of the chain on the second stream
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                pkd.msm_execute_batch_dev(t, self.buf_8n[0, :7 * n].view(7, n, 4), self.outs[10:17], self.zbytes[10:17])
        # the pointwise vanishing evaluation over the 8n points (wires: the LDEs of group 1; Z: the LDE just computed)
        self.vanishing(self.buf_8n if world == 1 else self.wires_8n)
        pkd.fft_dev(p8, self.buf_8n[0], self.buf_8n[1], inverse=True)
        pkd.fft_dev(p8, self.buf_8n[1], self.buf_8n[2], coset=True)
        pkd.fft_dev(p8, self.buf_8n[2], self.buf_8n[3], inverse=True, coset=True)
        if world == 1:
            pass
        else:
            lo, hi = self.mine(7)
            if hi > lo:
                pkd.msm_execute_batch_dev(t, self.buf_8n[0, lo * n:hi * n].view(hi - lo, n, 4), self.outs[10 + lo:10 + hi], self.zbytes[10 + lo:10 + hi])
        for i in (range(8) if world == 1 else range(*self.mine(8))):     # the remaining 8n transforms of the quotient arithmetic
            pkd.fft_dev(p8, self.scr_8n[i % 5], self.scr_8n[(i + 1) % 5], inverse=bool(i & 1))
        if world == 1:
            main.wait_stream(self.side)
        if not single:
            self.group_end(10, 17)
        # group 4
        if world == 1 or self.rank == 0:
            pkd.msm_execute_dev(t, self.buf_n[0], self.outs[17], self.zeros[17])
        if not single:
            self.group_end(17, 18)

    def vanishing(self, wires):
        import ctypes as C
        pk = self.cx.pk
        pk._check(pk.lib().plk_vanishing_points_dev(self.field, self.n, C.c_void_p(wires.data_ptr()), C.c_void_p(self.consts_8n.data_ptr()),
                                                    C.c_void_p(self.sigma_8n.data_ptr()), C.c_void_p(self.buf_8n[0].data_ptr()),
                                                    C.c_void_p(self.subgroup_8n.data_ptr()), C.c_void_p(self.params.data_ptr()),
                                                    C.c_void_p(self.van.data_ptr()), C.c_void_p(self.cx.torch.cuda.current_stream().cuda_stream)))

    def commitments(self):
        """(18, 3, 4) uint64 on the host: for N > 1 item i comes from the rank that owns it"""
        np = self.cx.np
        self.cx.torch.cuda.synchronize()
        if self.world == 1:
            return self.outs.cpu().numpy().view(np.uint64).copy()
        g = self.gathered.cpu().numpy().view(np.uint64)
        own = [self.owner(9, i) for i in range(9)] + [0] + [self.owner(7, i) for i in range(7)] + [0]
        return np.stack([g[own[i], i] for i in range(18)])


def bench(cx, log_n=16, reps=5, with_cpu=False):
    torch, np, pk = cx.torch, cx.np, cx.pk
    m = Mix(cx, log_n)
    for _ in range(2):
        m.proof()
    cx.barrier()
    l0 = pk.kernel_launch_count()
    t0 = time.perf_counter()
    for _ in range(reps):
        m.proof()
    cx.barrier()
    ms = cx.max_over_ranks((time.perf_counter() - t0) * 1e3 / reps)
    launches = (pk.kernel_launch_count() - l0) // reps
    # where one proof's time goes (last repetition, this rank): per dependency group, host time spent issuing its work and
    # the wait for the group's commitments
    mk = m.marks
    groups = [{"issue_ms": (mk[2 * g + 1] - mk[2 * g]) * 1e3, "wait_and_read_ms": (mk[2 * g + 2] - mk[2 * g + 1]) * 1e3} for g in range((len(mk) - 1) // 2)]
    got = m.commitments()
    # the vanishing evaluation alone (device resident, CUDA events)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wires = m.buf_8n if cx.world == 1 else m.wires_8n
    m.vanishing(wires)
    e0.record()
    for _ in range(5):
        m.vanishing(wires)
    e1.record()
    torch.cuda.synchronize()
    van_ms = e0.elapsed_time(e1) / 5
    checks = []
    # (1) the chain FFT(8n) -> IFFT(8n) -> coset LDE -> coset IFFT is the identity on the zero-padded coefficients
    n = m.n
    ok = bool(torch.equal(m.buf_8n[3, :n], m.buf_n[0])) and not bool(m.buf_8n[3, n:].any())
    checks.append("8n transform chain (FFT, IFFT, coset LDE, coset IFFT) returns the zero-padded coefficients bit for bit")
    # (2) N > 1: the replica-distributed commitments equal the single-GPU batched path on rank 0's GPU
    if cx.world > 1:
        van_dist = m.van.clone()
        m.proof(single=True)
        torch.cuda.synchronize()
        alone = m.outs.cpu().numpy().view(np.uint64)
        same = bool(np.array_equal(alone, got)) and bool(torch.equal(van_dist, m.van))
        t = torch.tensor([1 if (same and ok) else 0], dtype=torch.int64, device="cuda")
        cx.dist.all_reduce(t, op=cx.dist.ReduceOp.MIN)
        ok = bool(t.item())
        checks.append("the 18 commitments gathered from the ranks and the 8n vanishing evaluations == the single-GPU batched path recomputed on every rank")
    res = {"workload": f"prover L1 call mix, n = 2^{log_n} gates (18 MSM(n), 19 FFT(n), 13 FFT(8n), the 8n-point vanishing evaluation), wire values from host memory, "
                       "one D2H of the commitments per dependency group",
           "n_gpus": cx.world, "ms_per_proof_mix": ms, "proofs_per_sec": 1e3 / ms,
           "mode": "replicas: contiguous blocks of every dependency group per rank (batched launches), the Z / quotient chain replicated" if cx.world > 1
           else "single GPU, batched launches, independent items of a dependency group on two streams",
           "h2d_bytes_per_proof": m.h2d_bytes, "d2h_bytes_per_proof": 18 * 3 * 4 * 8, "timing": "host wall clock, barrier + synchronize on both sides, max over ranks",
           "launches_per_proof_rank0": int(launches), "groups_rank0": groups,
           "vanishing_points_ms": van_ms, "vanishing_points_per_sec": 8 * m.n / (van_ms * 1e-3),
           "vanishing_note": "plk_vanishing_points_dev over the 8n = 2^19 points alone: 24 rows x 32 B read + 32 B written and ~230 Montgomery products per point"}
    # (3) N = 1: the CPU restatement of the reference on the same inputs -- parity of every commitment and the CPU time of the mix
    if with_cpu and cx.rank == 0 and cx.world == 1:
        cpu = cpu_mix(cx, m, got)
        ok = ok and cpu.pop("ok")
        checks.append("all 18 commitments and the 9 IFFT(n) outputs == the CPU restatement (msm_execute_parallel w=11, ifft) on the same inputs; "
                      "the vanishing evaluation at 30 sampled points of the 8n domain == the big-integer restatement of plonk.rs:393-452")
        res["cpu_baseline"] = cpu
    res["verified"] = ok
    res["verification"] = checks
    m.table.close()
    return res


def vanishing_sample_check(cx, m, samples=24):
    """The device's vanishing evaluation at a sample of the 8n points against the big-integer restatement of
    plonk.rs:393-452 / gates/mod.rs:46-125 (oracle/plonky_oracle.py), on the mix's own device-resident inputs."""
    np = cx.np
    import plonky_oracle as po
    f = po.FIELDS_BY_ID[m.field]
    mm = 8 * m.n

    def ints(t):            # (..., 4) int64 Montgomery limbs -> canonical python ints
        a = t.cpu().numpy().view(np.uint64).reshape(-1, 4)
        return [f.from_mont(sum(int(a[i, j]) << (64 * j) for j in range(4))) for i in range(a.shape[0])]
    rng = np.random.Generator(np.random.PCG64(99))
    idx = sorted(set([0, 1, 7, 8, mm - 8, mm - 1] + [int(v) for v in rng.integers(0, mm, size=samples)]))
    need = sorted(set(j for i in idx for j in (i, (i + 8) % mm, (i + 8 * po.GRID_WIDTH) % mm)))
    pos = {j: k for k, j in enumerate(need)}
    sel = cx.torch.tensor(need, device="cuda")

    class Rows:             # sparse view: rows[j][i] for the indices the sample touches
        def __init__(self, t):
            vals = [ints(t[r].index_select(0, sel)) for r in range(t.shape[0])]
            self.vals = vals
        def __getitem__(self, r):
            outer = self
            class Row:
                def __getitem__(self, i):
                    return outer.vals[r][pos[i]]
            return Row()
    wires, consts, sigma = Rows(m.buf_8n), Rows(m.consts_8n), Rows(m.sigma_8n)
    zrow = Rows(m.buf_8n[0:1])[0]
    sub = Rows(m.subgroup_8n.unsqueeze(0))[0]
    prm = ints(m.params)
    want = po.vanishing_points(f, m.n, wires, consts, sigma, zrow, sub, prm[:6], prm[6], prm[7], prm[8], prm[9], prm[10], indices=idx)
    got = ints(m.van.index_select(0, cx.torch.tensor(idx, device="cuda")))
    return got == want


def cpu_mix(cx, m, got):
    """The same call mix through oracle/ref_port.cpp on the host cores (bench.py's cpu_baseline leg): timing + parity."""
    np = cx.np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_port as rp
    try:
        rp.build(native=True)
        L = rp.lib(native=True)
    except Exception:
        L = rp.lib()
    cores = os.cpu_count() or 1
    L.ref_set_threads(cores)
    n = m.n
    vals = m.vals_h.numpy().view(np.uint64)
    gens = m.pts.cpu().numpy().view(np.uint64)
    table = rp.MsmTable(m.curve, gens, None, 11, L)
    pn, p8 = rp.FftPlan(m.field, n, L), rp.FftPlan(m.field, 8 * n, L)
    t0 = time.perf_counter()
    coeffs = [pn.run(vals[i], inverse=True) for i in range(9)]
    pad = np.zeros((8 * n, 4), dtype=np.uint64)
    for i in range(9):
        pad[:n] = coeffs[i]
        p8.run(pad)
    commits = [table.execute(coeffs[i], parallel=True) for i in range(9)]
    for i in range(9):
        pn.run(vals[i], inverse=True)
    commits.append(table.execute(coeffs[0], parallel=True))
    pn.run(vals[0], inverse=True)
    pad[:n] = coeffs[0]
    x0 = p8.run(pad)
    x = x0
    for k in range(3 + 8):                     # the chain's other 8n transforms + the quotient arithmetic's
        x = p8.run(x, inverse=bool(k & 1))
    commits += [table.execute(x0[i * n:(i + 1) * n], parallel=True) for i in range(7)]
    commits.append(table.execute(coeffs[0], parallel=True))
    cpu_ms = (time.perf_counter() - t0) * 1e3
    ok = True
    for i, (xy, z) in enumerate(commits):
        ok = ok and (z == (not got[i].any())) and (z or bool(np.array_equal(got[i][:2], xy)))
    dev_coeffs = m.buf_n.cpu().numpy().view(np.uint64)
    # buf_n holds the second IFFT(n) batch of group 1 (row 0 is overwritten identically by group 2)
    ok = ok and all(bool(np.array_equal(dev_coeffs[i], coeffs[i])) for i in range(9))
    ok = ok and vanishing_sample_check(cx, m)
    return {"ok": ok, "value": cpu_ms, "unit": "ms per proof mix", "cores": cores, "kind": "port",
            "sample": "the same 18 MSM(n) + 19 (I)FFT(n) + 13 (I)FFT(8n) through the C++ restatement, one run, table and plans untimed; "
                      "WITHOUT the vanishing evaluation (restated in Python only, used here to check 30 sampled points)"}


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    import plonky_b200 as pk
    from plonky_b200 import distributed as pkd

    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=16)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()

    class Cx:
        pass
    cx = Cx()
    cx.np, cx.torch, cx.dist, cx.pk, cx.pkd = np, torch, dist, pk, pkd
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    pk._check(pk.lib().plk_set_device(local))
    if cx.world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if cx.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if cx.world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    cx.barrier, cx.max_over_ranks = barrier, max_over_ranks
    res = bench(cx, args.log_n, args.reps, with_cpu=(cx.world == 1))
    if cx.rank == 0:
        print(json.dumps(res))
    if cx.world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
