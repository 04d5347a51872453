"""ctypes binding of oracle/libplonky_ref_port.so (the C++ restatement of the reference).

TEST INFRASTRUCTURE ONLY -- see the header of ref_port.cpp.  Arrays are numpy uint64, Montgomery
limbs, little-endian, shape (n, L) for field elements and (n, 2, L) for affine points.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OP = dict(add=0, sub=1, mul=2, square=3, neg=4, inverse=5, to_canonical=6, from_canonical=7, double=8, triple=9)
FIELD_LIMBS = {0: 4, 1: 4, 2: 4, 3: 6}
CURVE_BASE = {0: 0, 1: 1, 2: 3}
CURVE_SCALAR = {0: 1, 1: 0, 2: 2}


def build(native: bool = False) -> str:
    target = "native" if native else "all"
    subprocess.check_call(["make", "-s", "-C", _DIR, target])
    return os.path.join(_DIR, "libplonky_ref_port_native.so" if native else "libplonky_ref_port.so")


def lib(native: bool = False):
    global _LIB
    if _LIB is not None and not native:
        return _LIB
    path = os.path.join(_DIR, "libplonky_ref_port_native.so" if native else "libplonky_ref_port.so")
    if not os.path.exists(path):
        path = build(native)
    L = C.CDLL(path)
    u64p, u8p, u32p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
    L.ref_field_op.argtypes = [C.c_int, C.c_int, u64p, u64p, u64p, C.c_size_t]
    L.ref_batch_inverse.argtypes = [C.c_int, u64p, u64p, C.c_size_t]
    L.ref_div2.argtypes = [C.c_int, u64p, u64p]
    L.ref_reverse_bits.argtypes = [C.c_uint64, C.c_int]
    L.ref_reverse_bits.restype = C.c_uint64
    L.ref_primitive_root_of_unity.argtypes = [C.c_int, C.c_int, u64p]
    L.ref_fft.argtypes = [C.c_int, u64p, u64p, C.c_size_t, C.c_int]
    L.ref_fft_precompute.argtypes = [C.c_int, C.c_size_t]
    L.ref_fft_precompute.restype = C.c_void_p
    L.ref_fft_run.argtypes = [C.c_void_p, u64p, u64p, C.c_int]
    L.ref_fft_free.argtypes = [C.c_void_p]
    L.ref_msm_precompute.argtypes = [C.c_int, u64p, u8p, C.c_size_t, C.c_int]
    L.ref_msm_precompute.restype = C.c_void_p
    L.ref_msm_execute.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_int, u64p, u8p]
    L.ref_msm_free.argtypes = [C.c_void_p]
    L.ref_curve_mul.argtypes = [C.c_int, u64p, C.c_uint8, u64p, u64p, u8p]
    L.ref_affine_sum.argtypes = [C.c_int, u64p, u8p, C.c_size_t, C.c_int, u64p, u8p]
    L.ref_gen_points.argtypes = [C.c_int, C.c_uint64, C.c_size_t, u64p]
    L.ref_blake_hash_usize_to_curve.argtypes = [C.c_int, C.c_uint64, C.c_size_t, u64p]
    L.ref_to_digits.argtypes = [C.c_int, u64p, C.c_int, u32p]
    L.ref_set_threads.argtypes = [C.c_int]
    L.ref_ipa_round_lr.argtypes = [C.c_int, u64p, u64p, u64p, u8p, C.c_size_t, u64p, u8p, u64p]
    L.ref_ipa_fold.argtypes = [C.c_int, u64p, u64p, u64p, u8p, C.c_size_t, u64p, u64p, u64p, u64p, u64p, u8p]
    if not native:
        _LIB = L
    return L


def _p64(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _p8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _arr(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def field_op(fid: int, op: str, a, b=None, L=None):
    L = L or lib()
    a = _arr(a)
    out = np.empty_like(a)
    bp = _p64(_arr(b)) if b is not None else None
    rc = L.ref_field_op(fid, OP[op], _p64(a), bp, _p64(out), a.shape[0])
    if rc == -2:
        raise ZeroDivisionError("No inverse")
    assert rc == 0
    return out


def batch_inverse(fid: int, a):
    a = _arr(a)
    out = np.empty_like(a)
    rc = lib().ref_batch_inverse(fid, _p64(a), _p64(out), a.shape[0])
    if rc == -2:
        raise ZeroDivisionError("No inverse")
    assert rc == 0
    return out


def div2(limbs):
    a = _arr(limbs)
    out = np.empty_like(a)
    assert lib().ref_div2(a.shape[0], _p64(a), _p64(out)) == 0
    return out


def reverse_bits(n: int, bits: int) -> int:
    return int(lib().ref_reverse_bits(n, bits))


def primitive_root_of_unity(fid: int, k: int):
    out = np.empty(FIELD_LIMBS[fid], dtype=np.uint64)
    assert lib().ref_primitive_root_of_unity(fid, k, _p64(out)) == 0
    return out


def fft(fid: int, a, inverse: bool = False):
    a = _arr(a)
    out = np.empty_like(a)
    rc = lib().ref_fft(fid, _p64(a), _p64(out), a.shape[0], 1 if inverse else 0)
    if rc != 0:
        raise AssertionError("Not a power of two")
    return out


class FftPlan:
    def __init__(self, fid: int, n: int, L=None):
        self.L = L or lib()
        self.h = self.L.ref_fft_precompute(fid, n)
        if not self.h:
            raise AssertionError("Not a power of two")

    def run(self, a, inverse=False):
        a = _arr(a)
        out = np.empty_like(a)
        assert self.L.ref_fft_run(self.h, _p64(a), _p64(out), 1 if inverse else 0) == 0
        return out

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_fft_free(self.h)
            self.h = None


class MsmTable:
    """msm_precompute (curve_msm.rs:27) -> execute / execute_parallel."""

    def __init__(self, cid: int, points_xy, zero=None, w: int = 11, L=None):
        self.L = L or lib()
        self.cid = cid
        xy = _arr(points_xy)
        n = xy.shape[0]
        z = np.ascontiguousarray(zero if zero is not None else np.zeros(n, dtype=np.uint8), dtype=np.uint8)
        self.n = n
        self.h = self.L.ref_msm_precompute(cid, _p64(xy), _p8(z), n, w)
        assert self.h

    def execute(self, scalars, parallel=True):
        s = _arr(scalars)
        nl = FIELD_LIMBS[CURVE_BASE[self.cid]]
        out = np.empty((2, nl), dtype=np.uint64)
        oz = np.zeros(1, dtype=np.uint8)
        rc = self.L.ref_msm_execute(self.h, _p64(s), s.shape[0], 1 if parallel else 0, _p64(out), _p8(oz))
        if rc == -3:
            raise AssertionError("precomputation / scalars length mismatch")
        assert rc == 0
        return out, bool(oz[0])

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_msm_free(self.h)
            self.h = None


def curve_mul(cid: int, point_xy, zero: bool, scalar):
    xy = _arr(point_xy)
    s = _arr(scalar)
    out = np.empty_like(xy)
    oz = np.zeros(1, dtype=np.uint8)
    assert lib().ref_curve_mul(cid, _p64(xy), 1 if zero else 0, _p64(s), _p64(out), _p8(oz)) == 0
    return out, bool(oz[0])


def affine_sum(cid: int, points_xy, zero=None, mode: int = 2):
    nl = FIELD_LIMBS[CURVE_BASE[cid]]
    xy = _arr(points_xy).reshape(-1, 2, nl)
    n = xy.shape[0]
    z = np.ascontiguousarray(zero if zero is not None else np.zeros(n, dtype=np.uint8), dtype=np.uint8)
    out = np.empty((2, nl), dtype=np.uint64)
    oz = np.zeros(1, dtype=np.uint8)
    assert lib().ref_affine_sum(cid, _p64(xy), _p8(z), n, mode, _p64(out), _p8(oz)) == 0
    return out, bool(oz[0])


def gen_points(cid: int, seed: int, n: int, L=None):
    nl = FIELD_LIMBS[CURVE_BASE[cid]]
    out = np.empty((n, 2, nl), dtype=np.uint64)
    assert (L or lib()).ref_gen_points(cid, seed, n, _p64(out)) == 0
    return out


def blake_hash_usize_to_curve(cid: int, seed_start: int, n: int, L=None):
    """[blake_hash_usize_to_curve(seed) for seed in seed_start .. seed_start + n) (hash_to_curve.rs:53-76): pedersen_g."""
    nl = FIELD_LIMBS[CURVE_BASE[cid]]
    out = np.zeros((n, 2, nl), dtype=np.uint64)
    assert (L or lib()).ref_blake_hash_usize_to_curve(cid, seed_start, n, _p64(out)) == 0
    return out


def to_digits(cid: int, scalar, w: int):
    s = _arr(scalar)
    out = np.zeros(512, dtype=np.uint32)
    nd = lib().ref_to_digits(cid, _p64(s), w, out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return [int(v) for v in out[:nd]]


def ipa_round_lr(cid: int, a, b, g_xy, g_zero=None):
    """halo.rs:87-93 without blinding / U' terms: ((l_xy, l_zero), (r_xy, r_zero), ip_l, ip_r)."""
    nl = FIELD_LIMBS[CURVE_BASE[cid]]
    a, b, g = _arr(a), _arr(b), _arr(g_xy).reshape(-1, 2, nl)
    n = a.shape[0]
    z = np.ascontiguousarray(g_zero if g_zero is not None else np.zeros(n, dtype=np.uint8), dtype=np.uint8)
    lr = np.empty((2, 2, nl), dtype=np.uint64)
    lrz = np.zeros(2, dtype=np.uint8)
    ip = np.empty((2, 4), dtype=np.uint64)
    rc = lib().ref_ipa_round_lr(cid, _p64(a), _p64(b), _p64(g), _p8(z), n, _p64(lr), _p8(lrz), _p64(ip))
    if rc == -2:
        raise AssertionError("Not a power of two")
    assert rc == 0
    return (lr[0], bool(lrz[0])), (lr[1], bool(lrz[1])), ip[0], ip[1]


def ipa_fold(cid: int, a, b, g_xy, g_zero, u, u_inv):
    """halo.rs:117-123: returns (a', b', g'_xy, g'_zero) of half the length."""
    nl = FIELD_LIMBS[CURVE_BASE[cid]]
    a, b, g = _arr(a), _arr(b), _arr(g_xy).reshape(-1, 2, nl)
    n = a.shape[0]
    z = np.ascontiguousarray(g_zero if g_zero is not None else np.zeros(n, dtype=np.uint8), dtype=np.uint8)
    oa, ob = np.empty((n // 2, 4), dtype=np.uint64), np.empty((n // 2, 4), dtype=np.uint64)
    og = np.empty((n // 2, 2, nl), dtype=np.uint64)
    oz = np.zeros(n // 2, dtype=np.uint8)
    rc = lib().ref_ipa_fold(cid, _p64(a), _p64(b), _p64(g), _p8(z), n, _p64(_arr(u)), _p64(_arr(u_inv)), _p64(oa), _p64(ob),
                            _p64(og), _p8(oz))
    if rc == -2:
        raise AssertionError("Not a power of two")
    assert rc == 0
    return oa, ob, og, oz
