"""Multi-GPU composition of the MSM: one process per GPU, torch.distributed (NCCL over NVLink) as plumbing.

The MSM is a sum over independent terms, so it shards with no data-path collective except the final
partial-sum reduce: rank r owns generators / scalars [r n/G, (r+1) n/G), runs the
full single-GPU pipeline on its shard and produces one un-normalised partial (XYZZ, 4 L limbs = 128 B
for the Tweedle curves).  Elliptic-curve addition is not an NCCL reduction op, so the G partials are
all-gathered (G x 128 B) and every rank adds them and normalises -- identical results on all ranks.

Device buffers are torch tensors (int64 views of u64 limbs); the kernels run on torch's current stream.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import lib, _check, FIELD_LIMBS, CURVE_BASE_FIELD, MsmPrecomputation

__all__ = ["msm_precompute_affine_dev", "points_generate_dev", "msm_execute_dev", "msm_execute_sharded", "fft_dev"]


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def points_generate_dev(curve: int, seed: int, n: int) -> torch.Tensor:
    """(n, 2, L) int64 tensor on the current device holding P_i = [splitmix64(seed + i)] G."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    pts = torch.empty((n, 2, Lb), dtype=torch.int64, device="cuda")
    _check(lib().plk_points_generate_dev(curve, seed, n, C.c_void_p(pts.data_ptr()), _stream_ptr()))
    return pts


def msm_precompute_affine_dev(curve: int, points_xy: torch.Tensor, w: int) -> MsmPrecomputation:
    """msm_precompute from device-resident affine generators (no identity points)."""
    assert points_xy.is_cuda and points_xy.dtype == torch.int64 and points_xy.is_contiguous()
    n = points_xy.shape[0]
    torch.cuda.current_stream().synchronize()
    h = C.c_void_p()
    _check(lib().plk_msm_precompute_affine_dev(curve, C.c_void_p(points_xy.data_ptr()), n, w, C.byref(h)))
    return MsmPrecomputation(curve, h, n, w, affine=True)


def msm_execute_dev(pre: MsmPrecomputation, scalars: torch.Tensor, out_xyz: torch.Tensor, out_zero: torch.Tensor):
    """Asynchronous msm_execute on device buffers: scalars (n, 4) int64, out_xyz (3, L) int64, out_zero (8,) uint8."""
    _check(lib().plk_msm_execute_dev(pre.handle, C.c_void_p(scalars.data_ptr()), scalars.shape[0],
                                     C.c_void_p(out_xyz.data_ptr()), C.c_void_p(out_zero.data_ptr()), _stream_ptr()))


def msm_execute_sharded(pre: MsmPrecomputation, scalars: torch.Tensor, partial: torch.Tensor, gathered: torch.Tensor,
                        out_xyz: torch.Tensor, out_zero: torch.Tensor, group=None):
    """This rank's shard -> partial; all-gather; combine.  With world_size 1 it is msm_execute_dev."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return msm_execute_dev(pre, scalars, out_xyz, out_zero)
    L = lib()
    _check(L.plk_msm_execute_partial_dev(pre.handle, C.c_void_p(scalars.data_ptr()), scalars.shape[0],
                                         C.c_void_p(partial.data_ptr()), _stream_ptr()))
    dist.all_gather_into_tensor(gathered, partial, group=group)
    _check(L.plk_msm_combine_partials_dev(pre.curve, C.c_void_p(gathered.data_ptr()), world,
                                          C.c_void_p(out_xyz.data_ptr()), C.c_void_p(out_zero.data_ptr()), _stream_ptr()))


def fft_dev(pre, d_in: torch.Tensor, d_out: torch.Tensor, inverse: bool = False, coset: bool = False):
    """Asynchronous transform(s) on device buffers: d_in (k, n_in, L) or (n_in, L); d_out (k, size, L) or (size, L)."""
    k = d_in.shape[0] if d_in.dim() == 3 else 1
    n_in = d_in.shape[-2]
    flags = (1 if inverse else 0) | (2 if coset else 0)
    _check(lib().plk_fft_dev(pre.handle, C.c_void_p(d_in.data_ptr()), n_in, k, flags, C.c_void_p(d_out.data_ptr()), _stream_ptr()))
