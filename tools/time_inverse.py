#!/usr/bin/env python3
"""Single-thread latency of the two device inversions (Fermat ladder vs binary GCD), through plk_field_op with n = 1."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plonky_b200 as pk
for field, limbs in ((0, 4), (3, 6)):
    x = np.array([[0x123456789abcdef1] * (limbs - 1) + [0x0123456789abcdef]], dtype=np.uint64)
    for op in ("double", "inverse", "inverse_gcd"):
        pk.field_op(field, op, x)
        t0 = time.perf_counter()
        for _ in range(200):
            pk.field_op(field, op, x)
        print(f"field {field} {op:12s} {(time.perf_counter() - t0) / 200 * 1e6:8.1f} us per call (n = 1, copies included)")
