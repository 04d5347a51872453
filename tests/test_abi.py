"""The C-ABI library loads on a CPU-only box and exports every symbol include/plonky_b200.h declares;
compute calls fail loudly (never fall back to a CPU path) when no CUDA device is present."""
import ctypes as C
import os

import numpy as np
import pytest

import __graft_entry__ as ge
import plonky_b200 as pk


def test_library_exports_every_declared_symbol():
    L = pk.lib()
    assert L.plk_abi_version() == 1
    syms = ge.exported_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(L, s), s


def test_static_queries_work_without_gpu():
    L = pk.lib()
    assert [L.plk_field_limbs(i) for i in range(5)] == [4, 4, 4, 6, 0]
    assert [L.plk_curve_base_field(i) for i in range(3)] == [0, 1, 3]
    assert [L.plk_curve_scalar_field(i) for i in range(3)] == [1, 0, 2]
    assert L.plk_status_string(3) == b"Not a power of two"
    assert L.plk_msm_partial_limbs(0) == 16 and L.plk_msm_partial_limbs(2) == 24


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point returns an error status (surfaced as an
    exception by the python mirror); nothing is computed on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises((pk.CudaError, ValueError, pk.PlonkyPanic)):
        pk.fft_precompute(pk.TWEEDLEDEE_BASE, 8)
    with pytest.raises((pk.CudaError, ValueError, pk.PlonkyPanic)):
        pk.field_op(pk.TWEEDLEDEE_BASE, "mul", np.ones((2, 4), dtype=np.uint64), np.ones((2, 4), dtype=np.uint64))
    with pytest.raises((pk.CudaError, ValueError, pk.PlonkyPanic)):
        pk.msm_precompute_affine(pk.TWEEDLEDEE, np.ones((2, 2, 4), dtype=np.uint64), 8)
    # the rows added next to the path: IPA rounds, generator derivation, point codec, polynomial helpers
    one = np.ones((2, 4), dtype=np.uint64)
    with pytest.raises((pk.CudaError, ValueError, pk.PlonkyPanic)):
        pk.HaloIpaRounds(pk.TWEEDLEDEE, one, one, np.ones((2, 2, 4), dtype=np.uint64))
    with pytest.raises((pk.CudaError, ValueError, pk.PlonkyPanic)):
        pk.blake_hash_usize_to_curve(pk.TWEEDLEDEE, 0, 4)
    with pytest.raises((pk.CudaError, ValueError, pk.PlonkyPanic)):
        pk.points_to_bytes(pk.TWEEDLEDEE, np.ones((2, 2, 4), dtype=np.uint64))
    with pytest.raises((pk.CudaError, ValueError, pk.PlonkyPanic)):
        pk.polynomial_mul(pk.TWEEDLEDEE_BASE, one, one)
    with pytest.raises((pk.CudaError, ValueError, pk.PlonkyPanic)):
        pk.permutation_polynomial(pk.TWEEDLEDEE_BASE, one, np.ones((2, 9, 4), dtype=np.uint64), np.ones((6, 16, 4), dtype=np.uint64),
                                  np.ones((6, 4), dtype=np.uint64), one[0], one[0])


def test_host_side_checks_of_the_new_rows():
    """Argument checks that mirror the reference's asserts run before any device work."""
    one = np.ones((4, 4), dtype=np.uint64)
    with pytest.raises(pk.PlonkyPanic):          # debug_assert_eq!(halo_b.len(), n), halo.rs:68
        pk.HaloIpaRounds(pk.TWEEDLEDEE, one, one[:2], np.ones((4, 2, 4), dtype=np.uint64))
    with pytest.raises(ValueError):              # unknown curve id
        pk.blake_hash_usize_to_curve(7, 0, 1)
    assert pk.lib().plk_point_compressed_bytes(pk.TWEEDLEDEE) == 33 and pk.lib().plk_point_compressed_bytes(pk.BLS12_377) == 49
    assert pk.lib().plk_point_compressed_bytes(9) == 0


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "plonky_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "plonky_oracle" not in text and "ref_port" not in text.replace("oracle/ref_port.cpp gen_points_t", ""), fn
