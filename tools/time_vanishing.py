#!/usr/bin/env python3
"""Time plk_vanishing_points_dev alone at n = 2^16 (8n = 2^19 points), device resident.  PLK_VANISH_MB selects the variant."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import plonky_b200 as pk
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
m = 8 * n
field = pk.TWEEDLEDUM_BASE
rng = np.random.Generator(np.random.PCG64(3))
def rnd(*shape):
    return torch.from_numpy(rng.integers(0, 1 << 62, size=shape + (4,), dtype=np.uint64).view(np.int64)).cuda()
wires, consts, sigma, z, params = rnd(9, m), rnd(6, m), rnd(6, m), rnd(m), rnd(11)
plan = pk.fft_precompute(field, m)
sub = torch.from_numpy(pk.fft_subgroup(plan).view(np.int64)).cuda()
out = torch.zeros((m, 4), dtype=torch.int64, device="cuda")
def run():
    pk._check(pk.lib().plk_vanishing_points_dev(field, n, *[C.c_void_p(t.data_ptr()) for t in (wires, consts, sigma, z, sub, params, out)],
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"PLK_VANISH_MB={os.environ.get('PLK_VANISH_MB', 'default')}: {ms:.3f} ms for {m} points = {m / ms / 1e3:.1f} M points/s; checksum {int(out.sum().item()) & 0xffffffff:08x}")
