#!/usr/bin/env python3
"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck), touching every kernel family: fixed-base and
variable-base MSM (shared-memory sort, quad tails), Halo IPA rounds in both modes, NTT / divide_by_z_h / polynomial
product, BLAKE3 generator derivation and the point codec.  Results are checked against the oracle.

    compute-sanitizer --tool memcheck  --error-exitcode 7 python tools/sanitizer_workload.py
    compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitizer_workload.py
Last run (round 1, B200): 0 errors, 0 hazards -- profiles/r1_compute_sanitizer.txt
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
print("sanitizer workload ok")
