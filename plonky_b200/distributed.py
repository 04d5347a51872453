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
from .sharding import shard_range, partial_layout

__all__ = ["msm_precompute_affine_dev", "points_generate_dev", "pedersen_generators_dev", "msm_execute_dev", "msm_execute_sharded", "msm_execute_batch_dev", "fft_dev",
           "msm_parallel_dev", "exchange_partials", "ShardedMsm", "DistributedNtt"]


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def points_generate_dev(curve: int, seed: int, n: int) -> torch.Tensor:
    """(n, 2, L) int64 tensor on the current device holding P_i = [splitmix64(seed + i)] G."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    pts = torch.empty((n, 2, Lb), dtype=torch.int64, device="cuda")
    _check(lib().plk_points_generate_dev(curve, seed, n, C.c_void_p(pts.data_ptr()), _stream_ptr()))
    return pts


def pedersen_generators_dev(curve: int, start: int, n: int) -> torch.Tensor:
    """(n, 2, L) int64 tensor: blake_hash_usize_to_curve(start + i), i < n -- the reference's pedersen_g
    (src/circuit_builder.rs:1127) derived on the device (src/hash_to_curve.rs:53-76)."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    pts = torch.empty((n, 2, Lb), dtype=torch.int64, device="cuda")
    _check(lib().plk_blake_hash_usize_to_curve_dev(curve, start, n, C.c_void_p(pts.data_ptr()), _stream_ptr()))
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


def msm_execute_batch_dev(pre: MsmPrecomputation, scalars_k: torch.Tensor, out_xyz: torch.Tensor, out_zero: torch.Tensor):
    """k executes against one table: scalars_k (k, n, 4) int64, out_xyz (k, 3, L) int64, out_zero (>= k,) uint8 --
    the device-resident body of commit_polynomials (src/plonk_util.rs:215-231)."""
    k, n = scalars_k.shape[0], scalars_k.shape[1]
    _check(lib().plk_msm_execute_batch_dev(pre.handle, C.c_void_p(scalars_k.data_ptr()), n, k, C.c_void_p(out_xyz.data_ptr()),
                                           C.c_void_p(out_zero.data_ptr()), _stream_ptr()))


def msm_parallel_dev(curve: int, scalars: torch.Tensor, points_xy: torch.Tensor, out_xyz: torch.Tensor, out_zero: torch.Tensor):
    """msm_parallel (curve_msm.rs:54-61) on device buffers, table-free: scalars (n, 4), points_xy (n, 2, L) int64."""
    _check(lib().plk_msm_parallel_dev(curve, C.c_void_p(scalars.data_ptr()), C.c_void_p(points_xy.data_ptr()), scalars.shape[0],
                                      C.c_void_p(out_xyz.data_ptr()), C.c_void_p(out_zero.data_ptr()), _stream_ptr()))


def msm_execute_partial_dev(pre: MsmPrecomputation, scalars: torch.Tensor, partial: torch.Tensor):
    """This shard's un-normalised sum (XYZZ, 4 L limbs) into `partial`."""
    _check(lib().plk_msm_execute_partial_dev(pre.handle, C.c_void_p(scalars.data_ptr()), scalars.shape[0],
                                             C.c_void_p(partial.data_ptr()), _stream_ptr()))


def msm_combine_partials_dev(curve: int, gathered: torch.Tensor, count: int, out_xyz: torch.Tensor, out_zero: torch.Tensor):
    """Sum of `count` gathered partials (layout: sharding.partial_layout) -> one normalised point."""
    _check(lib().plk_msm_combine_partials_dev(curve, C.c_void_p(gathered.data_ptr()), count,
                                              C.c_void_p(out_xyz.data_ptr()), C.c_void_p(out_zero.data_ptr()), _stream_ptr()))


def exchange_partials(partial: torch.Tensor, gathered: torch.Tensor, group=None) -> torch.Tensor:
    """The MSM's only collective: all-gather of the per-rank partials into the layout of sharding.partial_layout
    (rank r's limbs at offset r * limbs).  Backend-agnostic (NCCL on device tensors, gloo in the CPU tests)."""
    world = dist.get_world_size(group)
    off, total = partial_layout(world, partial.numel())
    assert gathered.numel() == total and off(world - 1) + partial.numel() == total
    dist.all_gather_into_tensor(gathered, partial, group=group)
    return gathered


def msm_execute_sharded(pre: MsmPrecomputation, scalars: torch.Tensor, partial: torch.Tensor, gathered: torch.Tensor,
                        out_xyz: torch.Tensor, out_zero: torch.Tensor, group=None):
    """This rank's shard -> partial; all-gather; combine.  With world_size 1 it is msm_execute_dev."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return msm_execute_dev(pre, scalars, out_xyz, out_zero)
    msm_execute_partial_dev(pre, scalars, partial)
    exchange_partials(partial, gathered, group)
    msm_combine_partials_dev(pre.curve, gathered, world, out_xyz, out_zero)


class ShardedMsm:
    """An n-term fixed-base MSM split over the ranks of `group`: rank r owns the terms
    sharding.shard_range(n, world, r), builds the table of its generators only and keeps the exchange buffers.

      sm = ShardedMsm(curve, n_total, make_points=lambda lo, hi: <(hi - lo, 2, L) device tensor>)
      sm.execute(local_scalars) -> (out_xyz, out_zero) device tensors, identical on every rank
    """

    def __init__(self, curve: int, n_total: int, make_points, w: int = 11, group=None, world: int = None, rank: int = None):
        self.curve, self.group = curve, group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.lo, self.hi = shard_range(n_total, self.world, self.rank)
        Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
        self.points = make_points(self.lo, self.hi)
        self.table = msm_precompute_affine_dev(curve, self.points, w)
        limbs = 4 * Lb
        _, total = partial_layout(self.world, limbs)
        self.partial = torch.zeros(limbs, dtype=torch.int64, device="cuda")
        self.gathered = torch.zeros(total, dtype=torch.int64, device="cuda")
        self.out_xyz = torch.zeros((3, Lb), dtype=torch.int64, device="cuda")
        self.out_zero = torch.zeros(8, dtype=torch.uint8, device="cuda")

    def __len__(self):
        return self.hi - self.lo

    def execute(self, local_scalars: torch.Tensor):
        assert local_scalars.shape[0] == self.hi - self.lo
        msm_execute_sharded(self.table, local_scalars, self.partial, self.gathered, self.out_xyz, self.out_zero, self.group)
        return self.out_xyz, self.out_zero


def fft_dev(pre, d_in: torch.Tensor, d_out: torch.Tensor, inverse: bool = False, coset: bool = False):
    """Asynchronous transform(s) on device buffers: d_in (k, n_in, L) or (n_in, L); d_out (k, size, L) or (size, L)."""
    k = d_in.shape[0] if d_in.dim() == 3 else 1
    n_in = d_in.shape[-2]
    flags = (1 if inverse else 0) | (2 if coset else 0)
    _check(lib().plk_fft_dev(pre.handle, C.c_void_p(d_in.data_ptr()), n_in, k, flags, C.c_void_p(d_out.data_ptr()), _stream_ptr()))


class DistributedNtt:
    """Domain-split NTT of size N = 2^log_n over `world` GPUs: four-step with ONE all-to-all
    (SURVEY.md section 8(e)).  N = R1 * M with R1 = 2^log_r1 (<= 256) the digit transformed after the exchange.

      input  (rank r): rows j_1 in [r R1/G, (r+1) R1/G) of the matrix x[j_1 + R1 j'], shape (R1/G, M, L)
      output (rank s): X[k' + M k_1] at [k_1][kl], k' = s M/G + kl, shape (R1, M/G, L)

    With world == 1 the input is the R1 x M "decimated" view of x and the output is X in natural order.
    `input_rows(x)` / `gather_natural(out)` convert between these layouts and natural order (tests, G = 1)."""

    def __init__(self, field: int, log_n: int, world: int = None, rank: int = None, group=None, log_r1: int = None, p2p=None):
        from . import fft_precompute, FIELD_LIMBS as FL
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        g_log = self.world.bit_length() - 1
        assert 1 << g_log == self.world, "world size must be a power of two"
        if log_r1 is None:
            log_r1 = min(8, log_n - g_log - 3) if log_n - g_log - 3 >= g_log else min(8, log_n // 2)
        assert g_log <= log_r1 <= 8 and log_n - log_r1 >= g_log, "transform too small for this world size"
        self.field, self.log_n, self.log_r1, self.log_m = field, log_n, log_r1, log_n - log_r1
        self.L = FL[field]
        self.rows = (1 << log_r1) // self.world
        self.cols = (1 << self.log_m) // self.world
        self.plan_m = fft_precompute(field, 1 << self.log_m)
        self.plan_n = fft_precompute(field, 1 << log_n)
        shape_in = (self.rows, 1 << self.log_m, self.L)
        self.work = torch.empty(shape_in, dtype=torch.int64, device="cuda")
        self.send = torch.empty(shape_in, dtype=torch.int64, device="cuda")
        self.recv = torch.empty((1 << log_r1, self.cols, self.L), dtype=torch.int64, device="cuda")
        self.exchange_kind = "NCCL all_to_all_single" if self.world > 1 else "local copy"
        self._p2p = None
        self.copy_out = True       # P2P mode: copy the result from the IPC buffer into self.recv (a torch tensor); the timed
                                   # loop of bench.py switches it off and reads the IPC buffer's pointer (result_ptr) instead
        if self.world > 1 and self.world <= 8 and p2p is not False and dist.is_initialized() and dist.get_world_size(group) == self.world:
            err = None
            try:
                self._setup_p2p()
            except Exception as e:                      # noqa: BLE001 -- e.g. IPC not permitted in this container
                err = e
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)          # all ranks take the same path
            if int(ok.item()) == 1:
                self.exchange_kind = "fused into phase A: NVLink peer stores (CUDA IPC) + one 1-element all-reduce as the barrier"
            else:
                self._p2p = None
                self.p2p_error = repr(err) if err else "a peer failed to map the buffers"
                if p2p is True:
                    raise RuntimeError(f"peer-to-peer exchange unavailable: {self.p2p_error}")

    def _setup_p2p(self):
        """Two receive buffers per rank (alternating per transform), cudaMalloc'ed by the library and mapped into every
        peer process with CUDA IPC; phase A then stores straight into them over NVLink."""
        L = lib()
        nbytes = self.recv.numel() * 8
        own, handles = [], []
        for _ in range(2):
            ptr = C.c_void_p()
            h = (C.c_uint8 * 64)()
            _check(L.plk_ipc_alloc(nbytes, C.byref(ptr), h))
            own.append(ptr.value)
            handles.append(bytes(h))
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handles, group=self.group)
        tables, opened = [], []
        for b in range(2):
            arr = (C.c_void_p * self.world)()
            for r in range(self.world):
                if r == self.rank:
                    arr[r] = own[b]
                else:
                    ptr = C.c_void_p()
                    hb = (C.c_uint8 * 64).from_buffer_copy(gathered[r][b])
                    _check(L.plk_ipc_open(hb, C.byref(ptr)))
                    arr[r] = ptr.value
                    opened.append(ptr.value)
            tables.append(arr)
        self._p2p = {"own": own, "tables": tables, "opened": opened, "turn": 0, "flag": torch.zeros(1, dtype=torch.int32, device="cuda")}

    def close(self):
        if self._p2p:
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier(group=self.group)
            L = lib()
            for ptr in self._p2p["opened"]:
                L.plk_ipc_close(C.c_void_p(ptr))
            if dist.is_initialized():
                dist.barrier(group=self.group)
            for ptr in self._p2p["own"]:
                L.plk_ipc_free(C.c_void_p(ptr))
            self._p2p = None

    def phase_a(self, local_rows: torch.Tensor, inverse: bool = False, rank: int = None, send: torch.Tensor = None):
        r = self.rank if rank is None else rank
        send = self.send if send is None else send
        assert local_rows.shape == self.work.shape and local_rows.is_contiguous()
        _check(lib().plk_fft_dist_phase_a(self.plan_m.handle, self.plan_n.handle, C.c_void_p(local_rows.data_ptr()), self.rows,
                                          r * self.rows, self.world, 1 if inverse else 0, C.c_void_p(self.work.data_ptr()),
                                          C.c_void_p(send.data_ptr()), _stream_ptr()))
        return send

    def phase_b(self, recv: torch.Tensor, inverse: bool = False):
        _check(lib().plk_fft_dist_phase_b(self.plan_n.handle, C.c_void_p(recv.data_ptr()), self.log_r1,
                                          self.log_m - (self.world.bit_length() - 1), 1 if inverse else 0, _stream_ptr()))
        return recv

    def forward(self, local_rows: torch.Tensor, inverse: bool = False) -> torch.Tensor:
        """Phase A, all-to-all (NCCL), phase B.  Returns this rank's (R1, M/G, L) slice of the output."""
        if self._p2p:
            # exchange fused into phase A: stores go straight into the peers' receive buffers over NVLink
            st = self._p2p
            b = st["turn"]
            st["turn"] ^= 1
            _check(lib().plk_fft_dist_phase_a_p2p(self.plan_m.handle, self.plan_n.handle, C.c_void_p(local_rows.data_ptr()), self.rows,
                                                  self.rank * self.rows, self.world, 1 if inverse else 0, C.c_void_p(self.work.data_ptr()),
                                                  st["tables"][b], _stream_ptr()))
            dist.all_reduce(st["flag"], group=self.group)          # every rank's phase A has been issued and completed before phase B
            mine = C.c_void_p(st["own"][b])
            _check(lib().plk_fft_dist_phase_b(self.plan_n.handle, mine, self.log_r1, self.log_m - (self.world.bit_length() - 1),
                                              1 if inverse else 0, _stream_ptr()))
            self.result_ptr = st["own"][b]
            if self.copy_out:
                _check(lib().plk_copy_dev(C.c_void_p(self.recv.data_ptr()), mine, self.recv.numel() * 8, _stream_ptr()))
            return self.recv
        self.phase_a(local_rows, inverse)
        if self.world == 1:
            self.recv.copy_(self.send.view(self.recv.shape))
        else:
            dist.all_to_all_single(self.recv.view(-1), self.send.view(-1), group=self.group)
        return self.phase_b(self.recv, inverse)

    def check_against_single_gpu(self, local_rows: torch.Tensor, inverse: bool = False) -> bool:
        """Untimed self-check (bench.py): gather every rank's rows, run the single-GPU transform of the whole vector on
        this rank and compare this rank's output slice word for word.  -> AND over the ranks."""
        keep, self.copy_out = self.copy_out, True
        out = self.forward(local_rows, inverse).clone()
        self.copy_out = keep
        R1, M = 1 << self.log_r1, 1 << self.log_m
        if self.world > 1:
            full = torch.empty((R1, M, self.L), dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(full.view(-1), local_rows.contiguous().view(-1), group=self.group)
        else:
            full = local_rows
        x = full.permute(1, 0, 2).contiguous().view(R1 * M, self.L)           # x[j_1 + R1 j'] = rows[j_1][j']
        want = torch.empty_like(x)
        fft_dev(self.plan_n, x, want, inverse=inverse)
        mine = want.view(R1, M, self.L)[:, self.rank * self.cols:(self.rank + 1) * self.cols].contiguous()
        ok = torch.tensor([1 if torch.equal(out, mine) else 0], dtype=torch.int64, device="cuda")
        if self.world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        return bool(ok.item())

    # ---- layout helpers (host-side, for tests and for callers that start from natural order) ----
    def input_rows(self, x_natural: torch.Tensor, rank: int = None) -> torch.Tensor:
        """natural-order x (N, L) -> this rank's rows: x[j_1 + R1 j'] for its j_1 block."""
        r = self.rank if rank is None else rank
        R1 = 1 << self.log_r1
        m = x_natural.view(1 << self.log_m, R1, self.L).transpose(0, 1)        # [j_1][j']
        return m[r * self.rows:(r + 1) * self.rows].contiguous()

    def natural_from_outputs(self, outs) -> torch.Tensor:
        """per-rank outputs [(R1, M/G, L)] -> natural-order X (N, L): X[k' + M k_1] = outs[s][k_1][kl]."""
        return torch.cat(list(outs), dim=1).reshape(-1, self.L)
