"""plonky_b200 -- host-side mirror of plonky's L1 "bulk kernel" surface over the sm_100a C ABI.

The reference (0xPolygonZero/plonky) is Rust; its toolchain is not available in this image, so the
host side above the C ABI (include/plonky_b200.h) is mirrored here with the SAME function names,
argument meaning and error behaviour as the reference functions it fronts:

    msm_precompute / msm_execute / msm_execute_parallel / msm_parallel   src/curve/curve_msm.rs:27-157
    pedersen_hash                                                        src/plonk_util.rs:193-198
    fft_precompute / fft / fft_with_precomputation /
    fft_with_precomputation_power_of_2 / ifft_with_precomputation_power_of_2   src/fft.rs:42-156
    coset_lde / coset_ifft / divide_by_z_h                               src/polynomial.rs:330-380

Data: numpy uint64 arrays of little-endian limbs in Montgomery form (exactly the reference's `limbs`):
field vectors (n, L); projective points (n, 3, L) + uint8 zero flags; affine points (n, 2, L) + flags.
Panics of the reference (assert_eq!, log2_strict, "No inverse") surface as PlonkyPanic.

There is NO CPU fallback: importing works without a GPU (so the ABI can be inspected), but every
compute call goes to libplonky_b200.so and fails loudly if the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

__all__ = [
    "PlonkyPanic", "CudaError", "lib", "library_path",
    "TWEEDLEDEE_BASE", "TWEEDLEDUM_BASE", "BLS12_377_SCALAR", "BLS12_377_BASE",
    "TWEEDLEDEE", "TWEEDLEDUM", "BLS12_377",
    "MsmPrecomputation", "msm_precompute", "msm_precompute_affine", "msm_execute", "msm_execute_parallel",
    "msm_execute_batch", "msm_parallel", "pedersen_hash", "coeffs_vec_to_commitments", "commit_polynomials", "msm_table_info",
    "FftPrecomputation", "fft_precompute", "fft", "fft_with_precomputation", "fft_with_precomputation_power_of_2",
    "ifft_with_precomputation_power_of_2", "fft_batch", "coset_lde", "coset_ifft", "divide_by_z_h",
    "field_op", "field_to_bytes", "field_from_bytes", "batch_multiplicative_inverse", "batch_to_affine", "affine_summation_best", "affine_multisummation_best", "curve_mul", "points_generate", "kernel_launch_count",
    "polynomial_mul", "eval_domain", "from_evaluations", "permutation_polynomial", "vanishing_points", "vanishing_poly", "fft_subgroup",
    "HaloIpaRounds", "blake_hash_usize_to_curve", "blake_hash_base_field_to_curve", "points_to_bytes", "points_from_bytes",
]

# ids of include/plonky_b200.h
TWEEDLEDEE_BASE, TWEEDLEDUM_BASE, BLS12_377_SCALAR, BLS12_377_BASE = 0, 1, 2, 3
TWEEDLEDEE, TWEEDLEDUM, BLS12_377 = 0, 1, 2
class _IdMap(dict):
    """id -> value table whose misses are ValueError (unknown field / curve id), like PLK_EINVAL."""

    def __init__(self, what, *a):
        super().__init__(*a)
        self.what = what

    def __missing__(self, key):
        raise ValueError(f"unknown {self.what} id {key!r}")


FIELD_LIMBS = _IdMap("field", {0: 4, 1: 4, 2: 4, 3: 6})
CURVE_BASE_FIELD = _IdMap("curve", {0: 0, 1: 1, 2: 3})
CURVE_SCALAR_FIELD = _IdMap("curve", {0: 1, 1: 0, 2: 2})

PLK_OK, PLK_EINVAL, PLK_ELENGTH, PLK_ENOTPOW2, PLK_ESIZE, PLK_EZERO, PLK_ECUDA, PLK_ENOMEM, PLK_ETOOBIG = range(9)


class PlonkyPanic(AssertionError):
    """The reference panics here (assert_eq!, log2_strict, expect("No inverse"))."""


class CudaError(RuntimeError):
    pass


_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def library_path() -> str:
    return os.path.join(_HERE, "libplonky_b200.so")


def lib():
    """Load libplonky_b200.so (built in-tree by __graft_entry__.build() / plonky_b200/csrc/Makefile)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise CudaError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp, u64p, u8p, sz = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.c_size_t
    L.plk_status_string.restype = C.c_char_p
    L.plk_status_string.argtypes = [C.c_int]
    L.plk_last_error_message.restype = C.c_char_p
    L.plk_abi_version.restype = C.c_int
    L.plk_device_count.argtypes = [C.POINTER(C.c_int)]
    L.plk_set_device.argtypes = [C.c_int]
    L.plk_msm_precompute.argtypes = [C.c_int, u64p, u8p, sz, C.c_uint, C.POINTER(vp)]
    L.plk_msm_precompute_affine.argtypes = [C.c_int, u64p, u8p, sz, C.c_uint, C.POINTER(vp)]
    L.plk_msm_precompute_affine_dev.argtypes = [C.c_int, vp, sz, C.c_uint, C.POINTER(vp)]
    L.plk_msm_table_len.argtypes = [vp]
    L.plk_msm_table_len.restype = sz
    L.plk_msm_table_window.argtypes = [vp]
    L.plk_msm_table_window.restype = C.c_uint
    L.plk_msm_table_info.argtypes = [vp, C.POINTER(C.c_uint), C.c_int]
    L.plk_msm_free.argtypes = [vp]
    L.plk_msm_free.restype = None
    L.plk_msm_execute.argtypes = [vp, u64p, sz, u64p, u8p]
    L.plk_msm_execute_batch.argtypes = [vp, u64p, sz, sz, u64p, u8p]
    L.plk_msm_parallel.argtypes = [C.c_int, u64p, u64p, u8p, sz, C.c_uint, u64p, u8p]
    L.plk_msm_execute_dev.argtypes = [vp, vp, sz, vp, vp, vp]
    L.plk_msm_execute_batch_dev.argtypes = [vp, vp, sz, sz, vp, vp, vp]
    L.plk_msm_execute_partial_dev.argtypes = [vp, vp, sz, vp, vp]
    L.plk_msm_combine_partials_dev.argtypes = [C.c_int, vp, sz, vp, vp, vp]
    L.plk_commit_batch.argtypes = [vp, u64p, sz, sz, u64p, u64p, C.c_uint8, u64p, u8p]
    L.plk_msm_parallel_dev.argtypes = [C.c_int, vp, vp, sz, vp, vp, vp]
    L.plk_msm_partial_limbs.argtypes = [C.c_int]
    L.plk_msm_partial_limbs.restype = sz
    L.plk_fft_precompute.argtypes = [C.c_int, sz, C.POINTER(vp)]
    L.plk_fft_set_direct_log.argtypes = [vp, C.c_int]
    L.plk_fft_size.argtypes = [vp]
    L.plk_fft_size.restype = sz
    L.plk_fft_free.argtypes = [vp]
    L.plk_fft_free.restype = None
    L.plk_fft_pow2.argtypes = [vp, u64p, u64p, sz]
    L.plk_ifft_pow2.argtypes = [vp, u64p, u64p, sz]
    L.plk_fft.argtypes = [vp, u64p, sz, u64p]
    L.plk_fft_batch.argtypes = [vp, u64p, sz, sz, C.c_int, u64p]
    L.plk_coset_lde.argtypes = [vp, u64p, sz, u64p, u64p]
    L.plk_coset_ifft.argtypes = [vp, u64p, u64p, u64p]
    L.plk_divide_by_z_h.argtypes = [vp, u64p, sz, sz, u64p]
    L.plk_fft_dev.argtypes = [vp, vp, sz, sz, C.c_uint, vp, vp]
    L.plk_vanishing_points.argtypes = [C.c_int, sz] + [u64p] * 12
    L.plk_vanishing_points_dev.argtypes = [C.c_int, sz] + [vp] * 8
    L.plk_vanishing_poly.argtypes = [vp, sz] + [u64p] * 11
    L.plk_fft_subgroup.argtypes = [vp, u64p]
    L.plk_permutation_polynomial.argtypes = [C.c_int, sz, C.c_uint, u64p, u64p, sz, u64p, sz, sz, u64p, u64p, u64p, u64p]
    L.plk_poly_mul.argtypes = [C.c_int, u64p, sz, u64p, sz, u64p, sz, C.POINTER(sz)]
    L.plk_fft_dist_phase_a.argtypes = [vp, vp, vp, sz, sz, C.c_uint, C.c_uint, vp, vp, vp]
    L.plk_fft_dist_phase_b.argtypes = [vp, vp, C.c_uint, C.c_uint, C.c_uint, vp]
    L.plk_fft_dist_phase_a_p2p.argtypes = [vp, vp, vp, sz, sz, C.c_uint, C.c_uint, vp, C.POINTER(vp), vp]
    L.plk_ipc_alloc.argtypes = [sz, C.POINTER(vp), u8p]
    L.plk_ipc_open.argtypes = [u8p, C.POINTER(vp)]
    L.plk_ipc_close.argtypes = [vp]
    L.plk_ipc_free.argtypes = [vp]
    L.plk_copy_dev.argtypes = [vp, vp, sz, vp]
    L.plk_field_op.argtypes = [C.c_int, C.c_int, u64p, u64p, u64p, sz]
    L.plk_batch_inverse.argtypes = [C.c_int, u64p, u64p, sz]
    L.plk_field_to_bytes.argtypes = [C.c_int, u64p, sz, u8p]
    L.plk_field_from_bytes.argtypes = [C.c_int, u8p, sz, u64p]
    L.plk_batch_to_affine.argtypes = [C.c_int, u64p, u8p, sz, u64p, u8p]
    L.plk_affine_summation.argtypes = [C.c_int, u64p, u8p, sz, u64p, u8p]
    L.plk_affine_multisummation.argtypes = [C.c_int, u64p, u8p, u64p, sz, u64p, u8p]
    L.plk_curve_mul.argtypes = [C.c_int, u64p, u8p, u64p, sz, u64p, u8p]
    L.plk_points_generate_dev.argtypes = [C.c_int, C.c_uint64, sz, vp, vp]
    L.plk_points_generate.argtypes = [C.c_int, C.c_uint64, sz, u64p]
    L.plk_ipa_new.argtypes = [C.c_int, u64p, u64p, u64p, u8p, sz, C.POINTER(vp)]
    L.plk_ipa_new_with_table.argtypes = [vp, u64p, u64p, sz, C.POINTER(vp)]
    L.plk_ipa_len.argtypes = [vp]
    L.plk_ipa_len.restype = sz
    L.plk_ipa_free.argtypes = [vp]
    L.plk_ipa_free.restype = None
    L.plk_ipa_round_lr.argtypes = [vp, u64p, u8p, u64p, u8p, u64p, u64p]
    L.plk_ipa_fold.argtypes = [vp, u64p, u64p]
    L.plk_ipa_read.argtypes = [vp, u64p, u64p, u64p, u8p]
    L.plk_blake_hash_usize_to_curve.argtypes = [C.c_int, C.c_uint64, sz, u64p]
    L.plk_blake_hash_usize_to_curve_dev.argtypes = [C.c_int, C.c_uint64, sz, vp, vp]
    L.plk_blake_hash_base_field_to_curve.argtypes = [C.c_int, u64p, sz, u64p]
    L.plk_point_compressed_bytes.argtypes = [C.c_int]
    L.plk_point_compressed_bytes.restype = sz
    L.plk_points_compress.argtypes = [C.c_int, u64p, u8p, sz, u8p]
    L.plk_points_decompress.argtypes = [C.c_int, u8p, sz, u64p, u8p, u8p]
    L.plk_kernel_launch_count.restype = C.c_uint64
    L.plk_set_profiling.argtypes = [C.c_int]
    L.plk_measure_mul_throughput.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.plk_msm_last_phase_ms.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
    L.plk_fft_last_pass_ms.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
    L.plk_fft_num_passes.argtypes = [vp]
    _LIB = L
    return L


def _check(status: int):
    if status == PLK_OK:
        return
    L = lib()
    msg = (L.plk_last_error_message() or b"").decode() or L.plk_status_string(status).decode()
    if status in (PLK_ELENGTH, PLK_ENOTPOW2, PLK_ESIZE, PLK_EZERO, PLK_ETOOBIG):
        raise PlonkyPanic(msg)
    if status == PLK_EINVAL:
        raise ValueError(msg)
    raise CudaError(msg)


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _p64(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _p8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8)) if a is not None else None


def _zero_flags(zero, n) -> Optional[np.ndarray]:
    if zero is None:
        return None
    z = np.ascontiguousarray(zero, dtype=np.uint8)
    assert z.shape == (n,)
    return z


def kernel_launch_count() -> int:
    return int(lib().plk_kernel_launch_count())


MSM_PHASES = ("count", "scan", "scatter", "accumulate", "bucket_sum", "range", "final")


def set_profiling(enabled: bool):
    """Record CUDA events between the kernels of every MSM execute / transform (measurement only)."""
    lib().plk_set_profiling(1 if enabled else 0)


def measure_mul_throughput(field: int) -> float:
    """Montgomery products / s the device sustains in `field` (measurement only)."""
    v = C.c_double()
    _check(lib().plk_measure_mul_throughput(field, C.byref(v)))
    return float(v.value)


def msm_table_info(pre: "MsmPrecomputation") -> dict:
    """The geometry the device chose for a table: window bits, windows, accumulation mode, products per bucket addition."""
    buf = (C.c_uint * 5)()
    k = lib().plk_msm_table_info(pre.handle, buf, 5)
    assert k == 5
    return {"c": int(buf[0]), "nwin": int(buf[1]), "mode": ("xyzz", "batched-affine + xyzz")[int(buf[2])], "products_per_add": int(buf[3]),
            "affine_rounds": int(buf[4])}


def msm_last_phase_ms(pre: "MsmPrecomputation"):
    buf = (C.c_float * 12)()
    n = lib().plk_msm_last_phase_ms(pre.handle, buf, 12)
    return [float(buf[i]) for i in range(max(n, 0))]


def fft_last_pass_ms(pre: "FftPrecomputation"):
    buf = (C.c_float * 12)()
    n = lib().plk_fft_last_pass_ms(pre.handle, buf, 12)
    return [float(buf[i]) for i in range(max(n, 0))]


# ------------------------------------------------------------------------------------------------
# MSM (src/curve/curve_msm.rs)
# ------------------------------------------------------------------------------------------------
class MsmPrecomputation:
    """MsmPrecomputation<C> (curve_msm.rs:16-25): here a device-resident table behind an opaque handle.
    Keeps (curve, generators, w) so that it stays clonable / serialisable like the reference struct."""

    def __init__(self, curve: int, handle, n: int, w: int, generators=None, zero=None, affine=False):
        self.curve, self.handle, self.n, self.w = curve, handle, n, w
        self.generators, self.zero, self.affine = generators, zero, affine

    def __len__(self):
        return self.n

    def close(self):
        if self.handle:
            lib().plk_msm_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def msm_precompute(curve: int, generators, w: int, zero=None, keep_host_copy: bool = False) -> MsmPrecomputation:
    """msm_precompute(generators: &[ProjectivePoint<C>], w) (curve_msm.rs:27-38).
    generators: (n, 3, L) projective Montgomery limbs, zero: optional (n,) flags."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    g = _u64(generators).reshape(-1, 3, Lb)
    n = g.shape[0]
    z = _zero_flags(zero, n)
    h = C.c_void_p()
    _check(lib().plk_msm_precompute(curve, _p64(g), _p8(z), n, w, C.byref(h)))
    return MsmPrecomputation(curve, h, n, w, g if keep_host_copy else None, z if keep_host_copy else None)


def msm_precompute_affine(curve: int, generators_xy, w: int, zero=None) -> MsmPrecomputation:
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    g = _u64(generators_xy).reshape(-1, 2, Lb)
    n = g.shape[0]
    z = _zero_flags(zero, n)
    h = C.c_void_p()
    _check(lib().plk_msm_precompute_affine(curve, _p64(g), _p8(z), n, w, C.byref(h)))
    return MsmPrecomputation(curve, h, n, w, affine=True)


def msm_execute(precomputation: MsmPrecomputation, scalars) -> Tuple[np.ndarray, bool]:
    """msm_execute(&precomputation, scalars) (curve_msm.rs:63-100).  Returns the ProjectivePoint as
    ((3, L) limbs x, y, z with z = ONE, zero flag)."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[precomputation.curve]]
    s = _u64(scalars).reshape(-1, 4)
    out = np.zeros((3, Lb), dtype=np.uint64)
    oz = np.zeros(1, dtype=np.uint8)
    _check(lib().plk_msm_execute(precomputation.handle, _p64(s), s.shape[0], _p64(out), _p8(oz)))
    return out, bool(oz[0])


def msm_execute_parallel(precomputation: MsmPrecomputation, scalars):
    """msm_execute_parallel (curve_msm.rs:102-157): same value as msm_execute."""
    return msm_execute(precomputation, scalars)


def pedersen_hash(xs, pedersen_g_msm_precomputation: MsmPrecomputation):
    """pedersen_hash(xs, &precomputation) (src/plonk_util.rs:193-198)."""
    return msm_execute_parallel(pedersen_g_msm_precomputation, xs)


def msm_execute_batch(precomputation: MsmPrecomputation, scalars_k):
    """k scalar vectors against one table (commit_polynomials, src/plonk_util.rs:215-231)."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[precomputation.curve]]
    s = _u64(scalars_k)
    k, n = s.shape[0], s.shape[1]
    s = s.reshape(k, n, 4)
    out = np.zeros((k, 3, Lb), dtype=np.uint64)
    oz = np.zeros(k, dtype=np.uint8)
    _check(lib().plk_msm_execute_batch(precomputation.handle, _p64(s), n, k, _p64(out), _p8(oz)))
    return out, oz.astype(bool)


def coeffs_vec_to_commitments(coefficients_vec, msm_precomputation: MsmPrecomputation, blinding_point_xy, blinding_factors=None,
                              blinding_point_zero: bool = False):
    """PolynomialCommitment::coeffs_vec_to_commitments / commit_polynomials (src/poly_commit.rs:52-66,
    src/plonk_util.rs:215-231) in one device call: ((k, 2, L) affine commitments, (k,) zero flags).  blinding_factors:
    (k, 4) scalars for `blinding = true` (the caller draws them, poly_commit.rs:39-43) or None for `blinding = false`."""
    curve = msm_precomputation.curve
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    s = _u64(coefficients_vec)
    k, n = s.shape[0], s.shape[1]
    s = s.reshape(k, n, 4)
    h = _u64(blinding_point_xy).reshape(2, Lb)
    bl = None
    if blinding_factors is not None:
        bl = _u64(blinding_factors).reshape(k, 4)
    out = np.zeros((k, 2, Lb), dtype=np.uint64)
    oz = np.zeros(k, dtype=np.uint8)
    _check(lib().plk_commit_batch(msm_precomputation.handle, _p64(s), n, k, _p64(bl) if bl is not None else None, _p64(h),
                                  1 if blinding_point_zero else 0, _p64(out), _p8(oz)))
    return out, oz.astype(bool)


commit_polynomials = coeffs_vec_to_commitments


def msm_parallel(curve: int, scalars, generators, w: int, zero=None):
    """msm_parallel(scalars, generators, w) (curve_msm.rs:54-61)."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    s = _u64(scalars).reshape(-1, 4)
    g = _u64(generators).reshape(-1, 3, Lb)
    if g.shape[0] != s.shape[0]:
        raise PlonkyPanic("precomputation / scalars length mismatch")
    z = _zero_flags(zero, g.shape[0])
    out = np.zeros((3, Lb), dtype=np.uint64)
    oz = np.zeros(1, dtype=np.uint8)
    _check(lib().plk_msm_parallel(curve, _p64(s), _p64(g), _p8(z), s.shape[0], w, _p64(out), _p8(oz)))
    return out, bool(oz[0])


# ------------------------------------------------------------------------------------------------
# NTT (src/fft.rs)
# ------------------------------------------------------------------------------------------------
class FftPrecomputation:
    """FftPrecomputation<F> (fft.rs:28-40): device-resident twiddle tables behind an opaque handle;
    reconstructible from (field, degree)."""

    def __init__(self, field: int, degree: int, handle):
        self.field, self.degree, self.handle = field, degree, handle

    def size(self) -> int:
        return int(lib().plk_fft_size(self.handle))

    def set_direct_log(self, log2_entries: int):
        """0: inter-pass twiddles always on the fly (no N-entry table); default 24 (see plk_fft_set_direct_log)."""
        _check(lib().plk_fft_set_direct_log(self.handle, log2_entries))

    def close(self):
        if self.handle:
            lib().plk_fft_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fft_precompute(field: int, degree: int) -> FftPrecomputation:
    """fft_precompute(degree) (fft.rs:47-59)."""
    h = C.c_void_p()
    _check(lib().plk_fft_precompute(field, degree, C.byref(h)))
    return FftPrecomputation(field, degree, h)


def fft_with_precomputation_power_of_2(coefficients, precomputation: FftPrecomputation) -> np.ndarray:
    """fft.rs:103-156."""
    L = FIELD_LIMBS[precomputation.field]
    a = _u64(coefficients).reshape(-1, L)
    out = np.empty_like(a)
    _check(lib().plk_fft_pow2(precomputation.handle, _p64(a), _p64(out), a.shape[0]))
    return out


def ifft_with_precomputation_power_of_2(points, precomputation: FftPrecomputation) -> np.ndarray:
    """fft.rs:82-101."""
    L = FIELD_LIMBS[precomputation.field]
    a = _u64(points).reshape(-1, L)
    out = np.empty_like(a)
    _check(lib().plk_ifft_pow2(precomputation.handle, _p64(a), _p64(out), a.shape[0]))
    return out


def fft_with_precomputation(coefficients, precomputation: FftPrecomputation) -> np.ndarray:
    """fft.rs:61-80: zero-pad to the next power of two."""
    L = FIELD_LIMBS[precomputation.field]
    a = _u64(coefficients).reshape(-1, L)
    out = np.empty((precomputation.size(), L), dtype=np.uint64)
    _check(lib().plk_fft(precomputation.handle, _p64(a), a.shape[0], _p64(out)))
    return out


def fft(field: int, coefficients) -> np.ndarray:
    """fft(coefficients) (fft.rs:42-45)."""
    L = FIELD_LIMBS[field]
    a = _u64(coefficients).reshape(-1, L)
    pre = fft_precompute(field, a.shape[0])
    try:
        return fft_with_precomputation(a, pre)
    finally:
        pre.close()


def fft_batch(rows, precomputation: FftPrecomputation, inverse: bool = False) -> np.ndarray:
    """values_to_polynomials / polynomials_to_values_padded (src/plonk_util.rs:169-190): k rows at once;
    forward rows shorter than the plan size are zero-padded."""
    L = FIELD_LIMBS[precomputation.field]
    a = _u64(rows)
    k, n_in = a.shape[0], a.shape[1]
    out = np.empty((k, precomputation.size(), L), dtype=np.uint64)
    _check(lib().plk_fft_batch(precomputation.handle, _p64(a), n_in, k, 1 if inverse else 0, _p64(out)))
    return out


def coset_lde(coefficients, precomputation: FftPrecomputation, shift=None) -> np.ndarray:
    """c_i * g^i, zero-pad, FFT (src/polynomial.rs:336-347); shift=None uses MULTIPLICATIVE_SUBGROUP_GENERATOR."""
    L = FIELD_LIMBS[precomputation.field]
    a = _u64(coefficients).reshape(-1, L)
    out = np.empty((precomputation.size(), L), dtype=np.uint64)
    sp = _p64(_u64(shift)) if shift is not None else None
    _check(lib().plk_coset_lde(precomputation.handle, _p64(a), a.shape[0], sp, _p64(out)))
    return out


def coset_ifft(evaluations, precomputation: FftPrecomputation, shift=None) -> np.ndarray:
    """IFFT then scale coefficient i by g^-i (src/polynomial.rs:368-378)."""
    L = FIELD_LIMBS[precomputation.field]
    a = _u64(evaluations).reshape(-1, L)
    if a.shape[0] != precomputation.size():
        raise PlonkyPanic("Number of points does not match size of subgroup in precomputation")
    out = np.empty_like(a)
    sp = _p64(_u64(shift)) if shift is not None else None
    _check(lib().plk_coset_ifft(precomputation.handle, _p64(a), sp, _p64(out)))
    return out


def divide_by_z_h(coefficients, n_gates: int, precomputation: FftPrecomputation) -> np.ndarray:
    """Polynomial::divide_by_z_h (src/polynomial.rs:330-380)."""
    L = FIELD_LIMBS[precomputation.field]
    a = _u64(coefficients).reshape(-1, L)
    out = np.empty((precomputation.size(), L), dtype=np.uint64)
    _check(lib().plk_divide_by_z_h(precomputation.handle, _p64(a), a.shape[0], n_gates, _p64(out)))
    return out


# ------------------------------------------------------------------------------------------------
# helpers on the same arithmetic
# ------------------------------------------------------------------------------------------------
_OPS = dict(add=0, sub=1, mul=2, square=3, neg=4, inverse=5, to_canonical=6, from_canonical=7, double=8, inverse_gcd=9)


def field_op(field: int, op: str, a, b=None) -> np.ndarray:
    L = FIELD_LIMBS[field]
    a = _u64(a).reshape(-1, L)
    out = np.empty_like(a)
    bp = _p64(_u64(b).reshape(-1, L)) if b is not None else None
    _check(lib().plk_field_op(field, _OPS[op], _p64(a), bp, _p64(out), a.shape[0]))
    return out


def field_to_bytes(field: int, x) -> np.ndarray:
    """Field ToBytes (src/serialization.rs:17-21): (n, 8 L) bytes, to_canonical_u8_vec per element."""
    L = FIELD_LIMBS[field]
    a = _u64(x).reshape(-1, L)
    out = np.zeros((a.shape[0], 8 * L), dtype=np.uint8)
    _check(lib().plk_field_to_bytes(field, _p64(a), a.shape[0], _p8(out)))
    return out


def field_from_bytes(field: int, data) -> np.ndarray:
    """Field FromBytes (src/serialization.rs:23-30): ValueError("Out of range") where the reference returns io::Error."""
    L = FIELD_LIMBS[field]
    d = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, 8 * L)
    out = np.zeros((d.shape[0], L), dtype=np.uint64)
    _check(lib().plk_field_from_bytes(field, _p8(d), d.shape[0], _p64(out)))
    return out


def batch_multiplicative_inverse(field: int, x) -> np.ndarray:
    """Field::batch_multiplicative_inverse (src/field/field.rs:251-278); panics on a zero input."""
    L = FIELD_LIMBS[field]
    a = _u64(x).reshape(-1, L)
    out = np.empty_like(a)
    _check(lib().plk_batch_inverse(field, _p64(a), _p64(out), a.shape[0]))
    return out


def batch_to_affine(curve: int, points_xyz, zero=None):
    """ProjectivePoint::batch_to_affine (src/curve/curve.rs:216-232)."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    g = _u64(points_xyz).reshape(-1, 3, Lb)
    n = g.shape[0]
    z = _zero_flags(zero, n)
    out = np.zeros((n, 2, Lb), dtype=np.uint64)
    oz = np.zeros(n, dtype=np.uint8)
    _check(lib().plk_batch_to_affine(curve, _p64(g), _p8(z), n, _p64(out), _p8(oz)))
    return out, oz


def affine_multisummation_best(curve: int, summations, zeros=None):
    """affine_multisummation_best(summations: Vec<Vec<AffinePoint>>) (src/curve/curve_summations.rs:24-35):
    a list of (n_i, 2, L) arrays -> (k, 3, L) normalised projective sums + zero flags."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    lists = [_u64(s).reshape(-1, 2, Lb) for s in summations]
    k = len(lists)
    offsets = np.zeros(k + 1, dtype=np.uint64)
    for i, l in enumerate(lists):
        offsets[i + 1] = offsets[i] + np.uint64(l.shape[0])
    pts = np.concatenate(lists) if k and int(offsets[-1]) else np.zeros((0, 2, Lb), dtype=np.uint64)
    z = None
    if zeros is not None:
        z = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.uint8) for x in zeros]), dtype=np.uint8)
    out = np.zeros((k, 3, Lb), dtype=np.uint64)
    oz = np.zeros(k, dtype=np.uint8)
    _check(lib().plk_affine_multisummation(curve, _p64(pts), _p8(z), _p64(offsets), k, _p64(out), _p8(oz)))
    return out, oz


def affine_summation_best(curve: int, summation, zero=None):
    """affine_summation_best (src/curve/curve_summations.rs:18-22)."""
    out, oz = affine_multisummation_best(curve, [summation], None if zero is None else [zero])
    return out[0], bool(oz[0])


def curve_mul(curve: int, points_xyz, scalars, zero=None):
    """CurveScalar * ProjectivePoint, elementwise over n pairs (src/curve/curve_multiplication.rs:63-70)."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    g = _u64(points_xyz).reshape(-1, 3, Lb)
    s = _u64(scalars).reshape(-1, 4)
    if g.shape[0] != s.shape[0]:
        raise PlonkyPanic("points / scalars length mismatch")
    z = _zero_flags(zero, g.shape[0])
    out = np.zeros_like(g)
    oz = np.zeros(g.shape[0], dtype=np.uint8)
    _check(lib().plk_curve_mul(curve, _p64(g), _p8(z), _p64(s), g.shape[0], _p64(out), _p8(oz)))
    return out, oz


def points_generate(curve: int, seed: int, n: int) -> np.ndarray:
    """Synthetic generators P_i = [splitmix64(seed + i)] G as (n, 2, L) affine Montgomery limbs."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    out = np.zeros((n, 2, Lb), dtype=np.uint64)
    _check(lib().plk_points_generate(curve, seed, n, _p64(out)))
    return out


# ------------------------------------------------------------------------------------------------
# Halo inner-product-argument rounds (src/halo.rs:63-124)
# ------------------------------------------------------------------------------------------------
class HaloIpaRounds:
    """halo_a, halo_b, halo_g of batch_opening_proof (halo.rs:53-57) kept on the device for all rounds.

    round_lr()        -> halo.rs:87-93 without the blinding / U' terms (the caller adds [l_j] H + [<a,b>] U')
    fold(u, u_inv)    -> halo.rs:117-123
    read()            -> current (a, b, g_xy, g_zero); after the last fold halo_a[0], halo_b[0], halo_g[0].to_affine()
    """

    def __init__(self, curve: int, a, b, g_xy=None, g_zero=None, precomputation: "MsmPrecomputation" = None):
        Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
        a, b = _u64(a).reshape(-1, 4), _u64(b).reshape(-1, 4)
        n = a.shape[0]
        self.curve, self.Lb = curve, Lb
        self.handle = C.c_void_p()
        self.precomputation = precomputation          # keeps the borrowed table alive
        if precomputation is not None:
            # table mode: rounds run against the fixed-base table of pedersen_g, G is never folded
            if b.shape[0] != n:
                raise PlonkyPanic("halo_a / halo_b / halo_g length mismatch")
            _check(lib().plk_ipa_new_with_table(precomputation.handle, _p64(a), _p64(b), n, C.byref(self.handle)))
            return
        g = _u64(g_xy).reshape(-1, 2, Lb)
        if b.shape[0] != n or g.shape[0] != n:
            raise PlonkyPanic("halo_a / halo_b / halo_g length mismatch")        # debug_assert_eq!, halo.rs:67-69
        z = _zero_flags(g_zero, n)
        _check(lib().plk_ipa_new(curve, _p64(a), _p64(b), _p64(g), _p8(z), n, C.byref(self.handle)))

    def __len__(self):
        return int(lib().plk_ipa_len(self.handle))

    def round_lr(self):
        """((l_xyz, l_zero), (r_xyz, r_zero), ip_l, ip_r)"""
        l = np.zeros((3, self.Lb), dtype=np.uint64)
        r = np.zeros((3, self.Lb), dtype=np.uint64)
        lz, rz = np.zeros(1, dtype=np.uint8), np.zeros(1, dtype=np.uint8)
        ipl, ipr = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
        _check(lib().plk_ipa_round_lr(self.handle, _p64(l), _p8(lz), _p64(r), _p8(rz), _p64(ipl), _p64(ipr)))
        return (l, bool(lz[0])), (r, bool(rz[0])), ipl, ipr

    def fold(self, u, u_inv):
        u, u_inv = _u64(u).reshape(4), _u64(u_inv).reshape(4)
        _check(lib().plk_ipa_fold(self.handle, _p64(u), _p64(u_inv)))

    def read(self, with_g: bool = True):
        n = len(self)
        a, b = np.zeros((n, 4), dtype=np.uint64), np.zeros((n, 4), dtype=np.uint64)
        if not with_g:
            _check(lib().plk_ipa_read(self.handle, _p64(a), _p64(b), None, None))
            return a, b, None, None
        g = np.zeros((n, 2, self.Lb), dtype=np.uint64)
        z = np.zeros(n, dtype=np.uint8)
        _check(lib().plk_ipa_read(self.handle, _p64(a), _p64(b), _p64(g), _p8(z)))
        return a, b, g, z

    def close(self):
        if self.handle:
            lib().plk_ipa_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# Generator derivation (src/hash_to_curve.rs) and the point wire format (src/serialization.rs)
# ------------------------------------------------------------------------------------------------
def blake_hash_usize_to_curve(curve: int, seed_start: int, n: int = 1) -> np.ndarray:
    """[blake_hash_usize_to_curve::<C>(seed) for seed in seed_start .. seed_start + n] (hash_to_curve.rs:53-76) as
    (n, 2, L) affine Montgomery limbs -- pedersen_g is seeds 0..degree (circuit_builder.rs:1127)."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    out = np.zeros((n, 2, Lb), dtype=np.uint64)
    _check(lib().plk_blake_hash_usize_to_curve(curve, seed_start, n, _p64(out)))
    return out


def blake_hash_base_field_to_curve(curve: int, seeds) -> np.ndarray:
    """blake_hash_base_field_to_curve (hash_to_curve.rs:57-76) for (n, L) Montgomery seeds."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    s = _u64(seeds).reshape(-1, Lb)
    out = np.zeros((s.shape[0], 2, Lb), dtype=np.uint64)
    _check(lib().plk_blake_hash_base_field_to_curve(curve, _p64(s), s.shape[0], _p64(out)))
    return out


def points_to_bytes(curve: int, points_xy, zero=None) -> np.ndarray:
    """AffinePoint::write (serialization.rs:32-44) per point: (n, 1 + 8 L) bytes."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    g = _u64(points_xy).reshape(-1, 2, Lb)
    n = g.shape[0]
    z = _zero_flags(zero, n)
    out = np.zeros((n, 1 + 8 * Lb), dtype=np.uint8)
    _check(lib().plk_points_compress(curve, _p64(g), _p8(z), n, _p8(out)))
    return out


def points_from_bytes(curve: int, data):
    """AffinePoint::read (serialization.rs:46-72) per point: ((n, 2, L) limbs, zero flags).  Raises ValueError for
    "Out of range" / "Invalid x coordinate" (io::Error in the reference)."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    d = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1, 1 + 8 * Lb)
    n = d.shape[0]
    out = np.zeros((n, 2, Lb), dtype=np.uint64)
    z = np.zeros(n, dtype=np.uint8)
    st = np.zeros(n, dtype=np.uint8)
    _check(lib().plk_points_decompress(curve, _p8(d), n, _p64(out), _p8(z), _p8(st)))
    return out, z


# ------------------------------------------------------------------------------------------------
# Polynomial helpers that are thin wrappers of the transforms (src/polynomial.rs:135-151, 209-227)
# ------------------------------------------------------------------------------------------------
def eval_domain(coeffs, fft_precomputation: "FftPrecomputation") -> np.ndarray:
    """Polynomial::eval_domain (polynomial.rs:135-143): zero-pad to the domain size and transform."""
    return fft_with_precomputation(coeffs, fft_precomputation)


def from_evaluations(values, fft_precomputation: "FftPrecomputation") -> np.ndarray:
    """Polynomial::from_evaluations (polynomial.rs:146-151)."""
    return ifft_with_precomputation_power_of_2(values, fft_precomputation)


def polynomial_mul(field: int, a, b) -> np.ndarray:
    """Polynomial::mul (polynomial.rs:209-227): the un-trimmed IFFT output, 2^log2_ceil(deg a + deg b + 1) coefficients
    ((1, L) zeros when either operand is the zero polynomial)."""
    L = FIELD_LIMBS[field]
    a, b = _u64(a).reshape(-1, L), _u64(b).reshape(-1, L)
    cap = 1
    while cap < max(1, a.shape[0] + b.shape[0]):
        cap <<= 1
    out = np.zeros((cap, L), dtype=np.uint64)
    n = C.c_size_t()
    _check(lib().plk_poly_mul(field, _p64(a), a.shape[0], _p64(b), b.shape[0], _p64(out), cap, C.byref(n)))
    return out[:n.value].copy()


def fft_subgroup(precomputation: "FftPrecomputation") -> np.ndarray:
    """The plan's evaluation domain w^k, k < size, natural order (Circuit::subgroup_n / subgroup_8n, src/plonk.rs:47-51)."""
    L = FIELD_LIMBS[precomputation.field]
    out = np.zeros((precomputation.size(), L), dtype=np.uint64)
    _check(lib().plk_fft_subgroup(precomputation.handle, _p64(out)))
    return out


def _vanishing_args(field, degree, wires_8n, constants_8n, sigma_8n, k_is, alpha, beta, gamma, inner_zeta, inner_a):
    L = FIELD_LIMBS[field]
    m = 8 * degree
    w = _u64(wires_8n).reshape(9, m, L)
    c = _u64(constants_8n).reshape(6, m, L)
    s = _u64(sigma_8n).reshape(6, m, L)
    small = [_u64(v).reshape(-1, L) for v in (k_is, alpha, beta, gamma, inner_zeta, inner_a)]
    assert small[0].shape[0] == 6
    return w, c, s, small


def vanishing_points(field: int, degree: int, wires_8n, constants_8n, sigma_8n, z_8n, subgroup_8n, k_is, alpha, beta, gamma, inner_zeta,
                     inner_a) -> np.ndarray:
    """The par_iter body of Circuit::vanishing_poly (src/plonk.rs:393-452) over all 8n points: (8n, L) values."""
    L = FIELD_LIMBS[field]
    w, c, s, small = _vanishing_args(field, degree, wires_8n, constants_8n, sigma_8n, k_is, alpha, beta, gamma, inner_zeta, inner_a)
    z = _u64(z_8n).reshape(8 * degree, L)
    x = _u64(subgroup_8n).reshape(8 * degree, L)
    out = np.zeros((8 * degree, L), dtype=np.uint64)
    _check(lib().plk_vanishing_points(field, degree, _p64(w), _p64(c), _p64(s), _p64(z), _p64(x), *[_p64(v) for v in small], _p64(out)))
    return out


def vanishing_poly(fft_precomputation_8n: "FftPrecomputation", degree: int, wires_8n, constants_8n, sigma_8n, plonk_z_coeffs, k_is, alpha, beta,
                   gamma, inner_zeta, inner_a) -> np.ndarray:
    """Circuit::vanishing_poly (src/plonk.rs:375-456): the 8n coefficients of the vanishing polynomial."""
    field = fft_precomputation_8n.field
    L = FIELD_LIMBS[field]
    w, c, s, small = _vanishing_args(field, degree, wires_8n, constants_8n, sigma_8n, k_is, alpha, beta, gamma, inner_zeta, inner_a)
    zc = _u64(plonk_z_coeffs).reshape(degree, L)
    out = np.zeros((8 * degree, L), dtype=np.uint64)
    _check(lib().plk_vanishing_poly(fft_precomputation_8n.handle, degree, _p64(w), _p64(c), _p64(s), _p64(zc), *[_p64(v) for v in small], _p64(out)))
    return out


def permutation_polynomial(field: int, subgroup, wire_values, sigma_values, k_is, beta, gamma, num_routed: int = 6,
                           sigma_stride: int = 8) -> np.ndarray:
    """permutation_polynomial(degree, subgroup, witness, sigma_values, beta, gamma) (src/plonk_util.rs:233-262).
    subgroup (n, L); wire_values (n, NUM_WIRES, L) = Witness::wire_values; sigma_values (num_routed, sigma_stride * n, L)
    (the 8n-point evaluations the reference indexes with 8 * (i - 1)); k_is (num_routed, L); beta, gamma (L,)."""
    L = FIELD_LIMBS[field]
    sub = _u64(subgroup).reshape(-1, L)
    n = sub.shape[0]
    w = _u64(wire_values)
    w = w.reshape(n, -1, L)
    sig = _u64(sigma_values).reshape(num_routed, -1, L)
    k = _u64(k_is).reshape(num_routed, L)
    b, g = _u64(beta).reshape(L), _u64(gamma).reshape(L)
    out = np.zeros((n, L), dtype=np.uint64)
    _check(lib().plk_permutation_polynomial(field, n, num_routed, _p64(sub), _p64(w), w.shape[1], _p64(sig), sig.shape[1], sigma_stride,
                                            _p64(k), _p64(b), _p64(g), _p64(out)))
    return out
