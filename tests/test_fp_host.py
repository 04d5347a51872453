"""The kernels' field / curve arithmetic (plonky_b200/csrc/fp.cuh, ec.cuh) compiled for the HOST with
the PTX carry flag emulated, checked against the big-integer oracle.  This pins the limb-level
algorithm (even/odd CIOS Montgomery product, XYZZ group law) without a GPU; the same source is what
the CUDA kernels inline (GPU parity is tests/test_gpu_*.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import plonky_oracle as po
from helpers import mont_array, canon_list, points_to_array, array_to_point, rand_scalars, ints_to_limbs, limbs_to_ints

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "host_arith.cpp")
SO = os.path.join(HERE, "native", "libhost_arith.so")


@pytest.fixture(scope="module")
def lib():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    deps = [SRC] + [os.path.join(HERE, "..", "plonky_b200", "csrc", f) for f in ("fp.cuh", "ec.cuh", "field_constants.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO, SRC])
    L = C.CDLL(SO)
    u64p = C.POINTER(C.c_uint64)
    L.host_field_op.argtypes = [C.c_int, C.c_int, u64p, u64p, u64p, C.c_size_t]
    L.host_curve_sum.argtypes = [C.c_int, C.c_int, u64p, C.c_size_t, u64p]
    L.host_curve_mul64.argtypes = [C.c_int, u64p, C.c_uint64, u64p]
    return L


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def fop(L, f, op, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    out = np.empty_like(a)
    bp = _p(np.ascontiguousarray(b, dtype=np.uint64)) if b is not None else None
    assert L.host_field_op(f.fid, op, _p(a), bp, _p(out), a.shape[0]) == 0
    return out


@pytest.mark.parametrize("name", list(po.FIELDS))
def test_field_ops(lib, name):
    f = po.FIELDS[name]
    vals = po.field_test_inputs(f.p, 32)          # 32-bit word boundaries: the kernels use u32 limbs
    vals = vals[::5] + rand_scalars(f, 3, 40)
    a = [x for x in vals for _ in vals]
    b = [y for _ in vals for y in vals]
    A, B = mont_array(f, a), mont_array(f, b)
    assert canon_list(f, fop(lib, f, 0, A, B)) == [(x + y) % f.p for x, y in zip(a, b)]
    assert canon_list(f, fop(lib, f, 1, A, B)) == [(x - y) % f.p for x, y in zip(a, b)]
    assert canon_list(f, fop(lib, f, 2, A, B)) == [(x * y) % f.p for x, y in zip(a, b)]
    V = mont_array(f, vals)
    assert canon_list(f, fop(lib, f, 3, V)) == [x * x % f.p for x in vals]
    assert canon_list(f, fop(lib, f, 4, V)) == [(-x) % f.p for x in vals]
    assert canon_list(f, fop(lib, f, 8, V)) == [2 * x % f.p for x in vals]
    nz = [x for x in vals if x][:40]
    assert canon_list(f, fop(lib, f, 5, mont_array(f, nz))) == [f.inv(x) for x in nz]
    nz_all = [x for x in vals if x]
    assert canon_list(f, fop(lib, f, 9, mont_array(f, nz_all))) == [f.inv(x) for x in nz_all]      # binary-GCD inverse
    raw = ints_to_limbs(vals, f.limbs)
    m = fop(lib, f, 7, raw)
    assert limbs_to_ints(m) == [f.to_mont(x) for x in vals]
    assert limbs_to_ints(fop(lib, f, 6, m)) == vals
    # raw Montgomery products of arbitrary reduced limb patterns (inputs are "already Montgomery")
    ra, rb = rand_scalars(f, 11, 300), rand_scalars(f, 12, 300)
    rinv = f.inv(f.R)
    got = limbs_to_ints(fop(lib, f, 2, ints_to_limbs(ra, f.limbs), ints_to_limbs(rb, f.limbs)))
    assert got == [x * y * rinv % f.p for x, y in zip(ra, rb)]


@pytest.mark.parametrize("name", list(po.CURVES))
def test_curve_ops(lib, name):
    c = po.CURVES[name]
    f = c.base
    G = c.gen
    rng = po.SplitMix64(21 + c.cid)
    base = po.rand_points(c, rng, 6)
    cases = [
        [G, G], [G, c.double(G)], [G, G, G], [], [G, c.neg(G)], [None, G], [G, None, c.double(G)],
        base, base + [c.neg(base[2]), base[3], base[3]], [G, c.neg(G), G],
    ]
    for pts in cases:
        want = None
        for P in pts:
            want = c.add(want, P)
        xy, _ = points_to_array(c, pts)          # identity encoded as x = y = 0
        out = np.zeros((2, f.limbs), dtype=np.uint64)
        for op in (0, 1):
            assert lib.host_curve_sum(c.cid, op, _p(xy), len(pts), _p(out)) == 0
            got = None if not out.any() else array_to_point(c, out, False)
            assert got == want, (name, op, pts)
        want2 = None
        for P in pts:
            want2 = c.add(want2, c.double(P))
        assert lib.host_curve_sum(c.cid, 2, _p(xy), len(pts), _p(out)) == 0
        got = None if not out.any() else array_to_point(c, out, False)
        assert got == want2
    for k in (0, 1, 2, 3, 0xFFFFFFFFFFFFFFFF, 0x123456789ABCDEF):
        xy, _ = points_to_array(c, [base[0]])
        out = np.zeros((2, f.limbs), dtype=np.uint64)
        assert lib.host_curve_mul64(c.cid, _p(xy), k, _p(out)) == 0
        got = None if not out.any() else array_to_point(c, out, False)
        assert got == c.mul(k, base[0])
