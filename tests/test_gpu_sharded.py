"""The multi-GPU MSM path on ONE GPU: an n-term MSM is split into G shards exactly like plonky_b200.distributed.ShardedMsm
does across ranks (sharding.shard_range), every shard goes through plk_msm_execute_partial_dev against its own table,
the partials are concatenated in the all-gather layout (sharding.partial_layout) and reduced by
plk_msm_combine_partials_dev.  The result must equal the C++ restatement of the reference
(curve_msm.rs:102-157, oracle/ref_port.cpp) and the single-table plk_msm_execute on the whole set."""
import numpy as np
import pytest
import torch

import plonky_oracle as po
import plonky_b200 as pk
import ref_port as rp
from plonky_b200 import distributed as pkd
from plonky_b200.sharding import shard_range, partial_layout
from helpers import mont_array, rand_scalars, limbs_to_ints

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def sharded_msm(c, xy, scalars, G, w=11):
    """-> (out (3, L) uint64, zero flag) through partial + combine, one table per shard"""
    Lb = c.base.limbs
    n = xy.shape[0]
    limbs = 4 * Lb
    off, total = partial_layout(G, limbs)
    gathered = torch.zeros(total, dtype=torch.int64, device="cuda")
    tables = []
    for r in range(G):
        lo, hi = shard_range(n, G, r)
        t = pkd.msm_precompute_affine_dev(c.cid, _dev(xy[lo:hi]), w)
        tables.append(t)
        part = gathered[off(r):off(r) + limbs]
        pkd.msm_execute_partial_dev(t, _dev(scalars[lo:hi]), part)
    out = torch.zeros((3, Lb), dtype=torch.int64, device="cuda")
    oz = torch.zeros(8, dtype=torch.uint8, device="cuda")
    pkd.msm_combine_partials_dev(c.cid, gathered, G, out, oz)
    torch.cuda.synchronize()
    return out.cpu().numpy().view(np.uint64), bool(oz[0].item())


@pytest.mark.parametrize("name", ["Tweedledee", "Tweedledum", "Bls12377"])
@pytest.mark.parametrize("G", [2, 4, 8])
def test_sharded_msm_matches_port_and_single_table(name, G):
    c = po.CURVES[name]
    n = 1 << 12
    xy = rp.gen_points(c.cid, 21 + G, n)
    scalars = mont_array(c.scalar, rand_scalars(c.scalar, 33 + G, n))
    want_xy, want_zero = rp.MsmTable(c.cid, xy, None, 11).execute(scalars, parallel=True)
    got, gz = sharded_msm(c, xy, scalars, G)
    assert gz == want_zero and np.array_equal(got[:2], want_xy)
    one, oz = pk.msm_execute(pk.msm_precompute_affine(c.cid, xy, 11), scalars)
    assert gz == oz and np.array_equal(got, one)


def test_sharded_msm_large_shards_smem_sort_path():
    """2 shards of 2^14 terms: the shared-memory counting sort + 15-bit windows (the geometry the 8-GPU runs use)."""
    c = po.TWEEDLEDEE
    n = 1 << 15
    xy = pk.points_generate(c.cid, 99, n)
    scalars = mont_array(c.scalar, rand_scalars(c.scalar, 5, n))
    want_xy, want_zero = rp.MsmTable(c.cid, xy, None, 11).execute(scalars, parallel=True)
    got, gz = sharded_msm(c, xy, scalars, 2)
    assert gz == want_zero and np.array_equal(got[:2], want_xy)


@pytest.mark.parametrize("name", ["Tweedledee", "Bls12377"])
def test_sharded_msm_identity_and_cancelling_shards(name):
    """shard 0: all-zero scalars (identity partial); shards 1 and 2: same points, opposite scalars (they cancel);
    shard 3: ordinary.  Then everything cancels: the combined result is the identity."""
    c = po.CURVES[name]
    q = c.scalar.p
    m = 1 << 9
    base = rp.gen_points(c.cid, 77, 2 * m)
    xy = np.concatenate([base[:m], base[m:], base[m:], base[:m]])
    s = rand_scalars(c.scalar, 3, m)
    t = rand_scalars(c.scalar, 4, m)
    canon = [0] * m + s + [(q - v) % q for v in s] + t
    scalars = mont_array(c.scalar, canon)
    want_xy, want_zero = rp.MsmTable(c.cid, xy, None, 11).execute(scalars, parallel=True)
    got, gz = sharded_msm(c, xy, scalars, 4)
    assert not gz and gz == want_zero and np.array_equal(got[:2], want_xy)
    # the surviving shard alone gives the same point
    alone, az = pk.msm_execute(pk.msm_precompute_affine(c.cid, base[:m], 11), mont_array(c.scalar, t))
    assert np.array_equal(alone, got)
    # all four partials cancel pairwise -> identity
    canon2 = s + [(q - v) % q for v in s] + t + [(q - v) % q for v in t]
    xy2 = np.concatenate([base[m:], base[m:], base[:m], base[:m]])
    got2, gz2 = sharded_msm(c, xy2, mont_array(c.scalar, canon2), 4)
    assert gz2 and not got2.any()


def test_msm_parallel_dev_matches_table_path():
    """the table-free device path bench.py uses as its independent cross-check"""
    c = po.TWEEDLEDEE
    n = 1 << 13
    xy = pk.points_generate(c.cid, 5, n)
    scalars = mont_array(c.scalar, rand_scalars(c.scalar, 6, n))
    out = torch.zeros((3, 4), dtype=torch.int64, device="cuda")
    oz = torch.zeros(8, dtype=torch.uint8, device="cuda")
    pkd.msm_parallel_dev(c.cid, _dev(scalars), _dev(xy), out, oz)
    torch.cuda.synchronize()
    want_xy, want_zero = rp.MsmTable(c.cid, xy, None, 11).execute(scalars, parallel=True)
    assert bool(oz[0].item()) == want_zero and np.array_equal(out.cpu().numpy().view(np.uint64)[:2], want_xy)
