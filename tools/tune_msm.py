#!/usr/bin/env python3
"""Sweep of the MSM execute knobs on one GPU (tuning aid): PLK_MSM_AFFINE_ROUNDS x PLK_MSM_AFF_PER_THREAD x PLK_MSM_TASK.
   PLK_MSM_OVERLAP_PARTS (read once per process) is swept from the shell: the printed point_crc must not change.
   python tools/tune_msm.py [--log-n 20] [--curve 0]"""
import argparse, itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import plonky_b200 as pk
from plonky_b200 import distributed as pkd
from bench import rand_scalars_np

ap = argparse.ArgumentParser()
ap.add_argument("--log-n", type=int, default=20)
ap.add_argument("--curve", type=int, default=0)
ap.add_argument("--rounds", default="0,1,2,3")
ap.add_argument("--per", default="0,16,32,64")
ap.add_argument("--task", default="0")
ap.add_argument("--window", default="0")
args = ap.parse_args()
n = 1 << args.log_n
Lb = 6 if args.curve == 2 else 4
pts = pkd.pedersen_generators_dev(args.curve, 0, n)
sc = [torch.from_numpy(rand_scalars_np(n, 7 + i, args.curve).view(np.int64)).cuda() for i in range(2)]
out = torch.zeros((3, Lb), dtype=torch.int64, device="cuda"); oz = torch.zeros(8, dtype=torch.uint8, device="cuda")
ref = None
for w, r, per, task in itertools.product(args.window.split(","), args.rounds.split(","), args.per.split(","), args.task.split(",")):
    if r == "0" and per != args.per.split(",")[0]:
        continue
    for k, v in (("PLK_MSM_WINDOW", w), ("PLK_MSM_AFFINE_ROUNDS", r), ("PLK_MSM_AFF_PER_THREAD", per), ("PLK_MSM_TASK", task)):
        if v == "0" and k != "PLK_MSM_AFFINE_ROUNDS":
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    t = pkd.msm_precompute_affine_dev(args.curve, pts, 11)
    for i in range(3):
        pkd.msm_execute_dev(t, sc[i & 1], out, oz)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K = 10
    for i in range(K):
        pkd.msm_execute_dev(t, sc[i & 1], out, oz)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    pk.set_profiling(True)
    pkd.msm_execute_dev(t, sc[0], out, oz)
    ph = pk.msm_last_phase_ms(t)
    pk.set_profiling(False)
    res = out.cpu().numpy().copy()
    if ref is None:
        ref = res
    ok = bool(np.array_equal(res, ref))
    print(f"window={w} rounds={r} per={per} task={task}: {ms:.3f} ms  phases={' '.join(f'{x:.3f}' for x in ph)}  info={pk.msm_table_info(t)} same_point={ok} "
          f"overlap_parts={os.environ.get('PLK_MSM_OVERLAP_PARTS', 'auto')} point_crc={int(res.view(np.uint64).sum() & np.uint64(0xffffffff)):08x}", flush=True)
    t.close()
