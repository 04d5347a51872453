#!/usr/bin/env python3
"""k short MSMs against one table (the prover's commit_polynomials shape): run_batch's merged pipeline against the fork/join
path.  PLK_MSM_BATCH_MERGE is read once per process:
   python tools/time_msm_batch.py [--log-n 16] [--k 9];  PLK_MSM_BATCH_MERGE=16 python tools/time_msm_batch.py"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import plonky_b200 as pk
from plonky_b200 import distributed as pkd
from bench import rand_scalars_np

ap = argparse.ArgumentParser()
ap.add_argument("--log-n", type=int, default=16)
ap.add_argument("--k", type=int, default=9)
ap.add_argument("--curve", type=int, default=0)
args = ap.parse_args()
n, k = 1 << args.log_n, args.k
Lb = 6 if args.curve == 2 else 4
pts = pkd.pedersen_generators_dev(args.curve, 0, n)
t = pkd.msm_precompute_affine_dev(args.curve, pts, 11)
S = torch.from_numpy(np.stack([rand_scalars_np(n, 70 + i, args.curve) for i in range(k)]).view(np.int64)).cuda()
out = torch.zeros((k, 3, Lb), dtype=torch.int64, device="cuda")
oz = torch.zeros(max(k, 8), dtype=torch.uint8, device="cuda")
for _ in range(3):
    pkd.msm_execute_batch_dev(t, S, out, oz)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 20
e0.record()
for _ in range(K):
    pkd.msm_execute_batch_dev(t, S, out, oz)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
single = torch.zeros((3, Lb), dtype=torch.int64, device="cuda")
sz = torch.zeros(8, dtype=torch.uint8, device="cuda")
same = True
for i in range(k):
    pkd.msm_execute_dev(t, S[i], single, sz)
    torch.cuda.synchronize()
    same &= bool(torch.equal(single, out[i]))
e0.record()
for _ in range(K):
    pkd.msm_execute_dev(t, S[0], single, sz)
e1.record()
torch.cuda.synchronize()
print(f"k={k} n=2^{args.log_n} merge={os.environ.get('PLK_MSM_BATCH_MERGE', 'off')}: batch {ms:.3f} ms ({ms / k:.3f} ms per vector), "
      f"one execute alone {e0.elapsed_time(e1) / K:.3f} ms, points equal to single executes: {same}")
