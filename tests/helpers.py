"""Shared helpers for the parity tests: conversions between Python ints and limb arrays."""
import json
import os

import numpy as np

import plonky_oracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def kats():
    with open(os.path.join(GOLDEN, "reference_kats.json")) as f:
        return json.load(f)


def ints_to_limbs(vals, nlimbs):
    out = np.zeros((len(vals), nlimbs), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(nlimbs):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(arr):
    arr = np.asarray(arr, dtype=np.uint64)
    flat = arr.reshape(-1, arr.shape[-1])
    return [sum(int(flat[i, j]) << (64 * j) for j in range(flat.shape[1])) for i in range(flat.shape[0])]


def mont_array(field: po.Field, canon_vals):
    """canonical ints -> (n, L) Montgomery limb array."""
    return ints_to_limbs([field.to_mont(v) for v in canon_vals], field.limbs)


def canon_list(field: po.Field, mont_arr):
    return [field.from_mont(v) for v in limbs_to_ints(mont_arr)]


def points_to_array(curve: po.Curve, pts):
    """list of affine (x, y) | None -> ((n,2,L) Montgomery limbs, zero flags)."""
    f = curve.base
    n = len(pts)
    xy = np.zeros((n, 2, f.limbs), dtype=np.uint64)
    zero = np.zeros(n, dtype=np.uint8)
    for i, P in enumerate(pts):
        if P is None:
            zero[i] = 1
        else:
            xy[i, 0] = ints_to_limbs([f.to_mont(P[0])], f.limbs)[0]
            xy[i, 1] = ints_to_limbs([f.to_mont(P[1])], f.limbs)[0]
    return xy, zero


def array_to_point(curve: po.Curve, xy, zero):
    if zero:
        return None
    f = curve.base
    x, y = limbs_to_ints(np.asarray(xy).reshape(2, f.limbs))
    return (f.from_mont(x), f.from_mont(y))


def rand_scalars(field: po.Field, seed: int, n: int):
    rng = po.SplitMix64(seed)
    return [po.rand_field_limbs(field, rng) for _ in range(n)]


def splitmix_hash(z: int) -> int:
    M = (1 << 64) - 1
    z = (z + 0x9E3779B97F4A7C15) & M
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return z ^ (z >> 31)
