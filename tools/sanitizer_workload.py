#!/usr/bin/env python3
"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck), touching every kernel family: fixed-base and
variable-base MSM (shared-memory sort, quad tails), Halo IPA rounds in both modes, NTT / divide_by_z_h / polynomial
product, BLAKE3 generator derivation and the point codec.  Results are checked against the oracle.

    compute-sanitizer --tool memcheck  --error-exitcode 7 python tools/sanitizer_workload.py
    compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitizer_workload.py
Round 2 adds: the batched-affine bucket rounds (run with PLK_MSM_AFFINE_ROUNDS=2), the TMA / Stockham NTT pass (run with
PLK_NTT_TMA=1), the sharded partial + combine path, the vanishing-polynomial kernel and the one-call commitment.
Last runs: profiles/r1_compute_sanitizer.txt, profiles/r2_compute_sanitizer.txt
"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import plonky_b200 as pk, plonky_oracle as po, ref_port as rp
from helpers import mont_array, rand_scalars
c = po.TWEEDLEDEE
for n in (5000, 300):
    g = pk.blake_hash_usize_to_curve(c.cid, 0, n)
    s = mont_array(c.scalar, rand_scalars(c.scalar, 3, n))
    pre = pk.msm_precompute_affine(c.cid, g, 11)
    out, oz = pk.pedersen_hash(s, pre)
    want, wz = rp.MsmTable(c.cid, g, None, 11).execute(s, parallel=True)
    assert np.array_equal(out[:2], want) and oz == wz
    xyz = np.zeros((n, 3, 4), dtype=np.uint64); xyz[:, :2] = g; xyz[:, 2] = out[2]
    o2, z2 = pk.msm_parallel(c.cid, s, xyz, 8)
    assert np.array_equal(o2[:2], want)
n = 64
g = pk.blake_hash_usize_to_curve(c.cid, 0, n)
a = mont_array(c.scalar, rand_scalars(c.scalar, 5, n)); b = mont_array(c.scalar, rand_scalars(c.scalar, 6, n))
pre = pk.msm_precompute_affine(c.cid, g, 11)
for st in (pk.HaloIpaRounds(c.cid, a, b, g), pk.HaloIpaRounds(c.cid, a, b, precomputation=pre)):
    while len(st) > 1:
        st.round_lr(); st.fold(a[0], b[0])
    st.read()
f = po.TWEEDLEDEE_BASE
x = mont_array(f, rand_scalars(f, 9, 1 << 12))
plan = pk.fft_precompute(f.fid, 1 << 12)
y = pk.fft_with_precomputation_power_of_2(x, plan)
assert np.array_equal(pk.ifft_with_precomputation_power_of_2(y, plan), x)
pk.divide_by_z_h(x[:2048], 512, plan)
pk.polynomial_mul(f.fid, x[:100], x[:200])
enc = pk.points_to_bytes(c.cid, g); pk.points_from_bytes(c.cid, enc)
# ---- round 2 ----
# coset LDE (first pass of a zero-padded input: TMA path by default) and a 3-pass transform
x17 = mont_array(f, rand_scalars(f, 10, 1 << 14))
plan17 = pk.fft_precompute(f.fid, 1 << 17)
lde = pk.coset_lde(x17, plan17)
assert np.array_equal(pk.coset_ifft(lde, plan17)[:1 << 14], x17)
# sharded partial + combine on one device
import torch
from plonky_b200 import distributed as pkd
from plonky_b200.sharding import shard_range, partial_layout
n = 2048
g = pk.blake_hash_usize_to_curve(c.cid, 0, n)
s = mont_array(c.scalar, rand_scalars(c.scalar, 13, n))
off, total = partial_layout(2, 16)
gathered = torch.zeros(total, dtype=torch.int64, device="cuda")
tabs = []
for r in range(2):
    lo, hi = shard_range(n, 2, r)
    tabs.append(pkd.msm_precompute_affine_dev(c.cid, torch.from_numpy(g[lo:hi].view(np.int64)).cuda(), 11))
    pkd.msm_execute_partial_dev(tabs[-1], torch.from_numpy(s[lo:hi].view(np.int64)).cuda(), gathered[off(r):off(r) + 16])
o = torch.zeros((3, 4), dtype=torch.int64, device="cuda"); z = torch.zeros(8, dtype=torch.uint8, device="cuda")
pkd.msm_combine_partials_dev(c.cid, gathered, 2, o, z)
torch.cuda.synchronize()
want, wz = rp.MsmTable(c.cid, g, None, 11).execute(s, parallel=True)
assert np.array_equal(o.cpu().numpy().view(np.uint64)[:2], want)
# one-call commitment and the vanishing kernel
pre = pk.msm_precompute_affine(c.cid, g[:256], 11)
pk.coeffs_vec_to_commitments(np.stack([s[:256], s[256:512]]), pre, g[300], s[:2])
fs = c.scalar
deg = 16
rows = lambda k, seed: np.stack([mont_array(fs, rand_scalars(fs, seed + j, 8 * deg)) for j in range(k)])
p8 = pk.fft_precompute(fs.fid, 8 * deg)
small = [mont_array(fs, rand_scalars(fs, 80 + j, 6 if j == 0 else 1)) for j in range(6)]
pk.vanishing_poly(p8, deg, rows(9, 1), rows(6, 20), rows(6, 40), mont_array(fs, rand_scalars(fs, 60, deg)), *small)
print("sanitizer workload ok")
