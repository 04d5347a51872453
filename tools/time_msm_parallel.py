#!/usr/bin/env python3
"""Time msm_parallel (variable base, host buffers through the C ABI) at the Halo IPA sizes (src/halo.rs:87-91)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plonky_b200 as pk

for logn in (10, 12, 15, 18):
    n = 1 << logn
    xy = pk.points_generate(pk.TWEEDLEDEE, 5, n)
    xyz = np.zeros((n, 3, 4), dtype=np.uint64)
    xyz[:, :2] = xy
    xyz[:, 2] = np.array([0x7379f083fffffffd, 0xf5601c89c3d86ba3, 0xffffffffffffffff, 0x3fffffffffffffff], dtype=np.uint64)
    rng = np.random.Generator(np.random.PCG64(1))
    s = rng.integers(0, 1 << 62, size=(n, 4), dtype=np.uint64)
    pk.msm_parallel(pk.TWEEDLEDEE, s, xyz, 8)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        pk.msm_parallel(pk.TWEEDLEDEE, s, xyz, 8)
    dt = (time.perf_counter() - t0) / reps
    print(f"msm_parallel 2^{logn}: {dt * 1e3:.3f} ms end to end (H2D of points + scalars included) = {n / dt:.4g} scalar-muls/s")
