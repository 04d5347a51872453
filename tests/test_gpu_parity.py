"""GPU parity: the CUDA path (through the C ABI, plonky_b200 -> libplonky_b200.so) against the oracles
on identical seeded inputs.  Bit-exact: field vectors compared limb for limb (Montgomery form),
points compared on the normalised affine (x, y) + zero flag."""
import numpy as np
import pytest

import plonky_oracle as po
import ref_port as rp
import plonky_b200 as pk
from helpers import (kats, ints_to_limbs, limbs_to_ints, mont_array, canon_list, points_to_array,
                     array_to_point, rand_scalars, splitmix_hash)

pytestmark = pytest.mark.gpu
K = kats()


def proj_from_affine(curve, pts):
    """affine list -> (n,3,L) projective limbs with z = ONE (AffinePoint::to_projective, curve.rs:97-105)."""
    f = curve.base
    xy, zero = points_to_array(curve, pts)
    n = len(pts)
    xyz = np.zeros((n, 3, f.limbs), dtype=np.uint64)
    xyz[:, :2] = xy
    one = ints_to_limbs([f.R], f.limbs)[0]
    for i in range(n):
        if not zero[i]:
            xyz[i, 2] = one
    return xyz, zero


def result_point(curve, out, oz):
    f = curve.base
    if oz:
        assert not out.any()
        return None
    x, y, z = limbs_to_ints(out)
    assert z == f.R, "result must be normalised (z = ONE)"
    return (f.from_mont(x), f.from_mont(y))


# ---------------------------------------------------------------- field arithmetic on device
@pytest.mark.parametrize("name", list(po.FIELDS))
def test_field_ops(name):
    f = po.FIELDS[name]
    vals = po.field_test_inputs(f.p, 32)[::3] + rand_scalars(f, 3, 64)
    a = [x for x in vals for _ in vals]
    b = [y for _ in vals for y in vals]
    A, B = mont_array(f, a), mont_array(f, b)
    assert canon_list(f, pk.field_op(f.fid, "add", A, B)) == [(x + y) % f.p for x, y in zip(a, b)]
    assert canon_list(f, pk.field_op(f.fid, "sub", A, B)) == [(x - y) % f.p for x, y in zip(a, b)]
    assert canon_list(f, pk.field_op(f.fid, "mul", A, B)) == [(x * y) % f.p for x, y in zip(a, b)]
    V = mont_array(f, vals)
    assert canon_list(f, pk.field_op(f.fid, "square", V)) == [x * x % f.p for x in vals]
    assert canon_list(f, pk.field_op(f.fid, "neg", V)) == [(-x) % f.p for x in vals]
    assert canon_list(f, pk.field_op(f.fid, "double", V)) == [2 * x % f.p for x in vals]
    nz = [x for x in vals if x]
    assert canon_list(f, pk.field_op(f.fid, "inverse", mont_array(f, nz))) == [f.inv(x) for x in nz]
    assert canon_list(f, pk.field_op(f.fid, "inverse_gcd", mont_array(f, nz))) == [f.inv(x) for x in nz]
    raw = ints_to_limbs(vals, f.limbs)
    m = pk.field_op(f.fid, "from_canonical", raw)
    assert limbs_to_ints(m) == [f.to_mont(x) for x in vals]
    assert limbs_to_ints(pk.field_op(f.fid, "to_canonical", m)) == vals
    # the C++ restatement of the reference agrees limb for limb
    assert np.array_equal(pk.field_op(f.fid, "mul", A, B), rp.field_op(f.fid, "mul", A, B))
    with pytest.raises(pk.PlonkyPanic):
        pk.field_op(f.fid, "inverse", mont_array(f, [1, 0]))


@pytest.mark.parametrize("name", list(po.FIELDS))
def test_batch_inverse(name):
    f = po.FIELDS[name]
    x = [v for v in rand_scalars(f, 17, 1000) if v]
    got = pk.batch_multiplicative_inverse(f.fid, mont_array(f, x))
    assert canon_list(f, got) == [f.inv(v) for v in x]
    assert np.array_equal(got, rp.batch_inverse(f.fid, mont_array(f, x)))
    with pytest.raises(pk.PlonkyPanic):
        pk.batch_multiplicative_inverse(f.fid, mont_array(f, [3, 0, 5]))
    assert pk.batch_multiplicative_inverse(f.fid, np.zeros((0, f.limbs), dtype=np.uint64)).shape[0] == 0


# ---------------------------------------------------------------- NTT
def test_fft_and_ifft_reference_case():
    """fft.rs:164-185 on the device."""
    e = K["fft_and_ifft"]
    f = po.FIELDS[e["field"]]
    coeffs = [(i * e["mul"]) % e["mod"] for i in range(e["degree"])]
    pre = pk.fft_precompute(f.fid, e["degree"])
    assert pre.size() == 256
    points = pk.fft_with_precomputation(mont_array(f, coeffs), pre)
    assert canon_list(f, points) == po.dft_naive(f, coeffs + [0] * 56)
    back = pk.ifft_with_precomputation_power_of_2(points, pre)
    assert canon_list(f, back) == coeffs + [0] * 56


@pytest.mark.parametrize("name", list(po.FIELDS))
@pytest.mark.parametrize("logn", [0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 12, 13, 16])
def test_fft_matches_oracle(name, logn):
    f = po.FIELDS[name]
    if name == "Bls12377Base" and logn > 10:
        pytest.skip("6-limb field: small sizes only")
    n = 1 << logn
    x = rand_scalars(f, 2000 + logn, n)
    X = mont_array(f, x)
    pre = pk.fft_precompute(f.fid, n)
    got = pk.fft_with_precomputation_power_of_2(X, pre)
    if n <= 4096:
        assert canon_list(f, got) == po.ntt(f, x)
    want = rp.fft(f.fid, X)                       # C++ restatement of fft.rs (validated against big ints)
    assert np.array_equal(got, want)
    inv = pk.ifft_with_precomputation_power_of_2(got, pre)
    assert np.array_equal(inv, X)
    assert np.array_equal(pk.ifft_with_precomputation_power_of_2(X, pre), rp.fft(f.fid, X, inverse=True))


def test_fft_error_contract():
    f = po.TWEEDLEDEE_BASE
    pre = pk.fft_precompute(f.fid, 16)
    with pytest.raises(pk.PlonkyPanic):           # log2_strict, util.rs:16-19
        pk.fft_with_precomputation_power_of_2(mont_array(f, [1, 2, 3]), pre)
    with pytest.raises(pk.PlonkyPanic):           # size mismatch, fft.rs:107-111
        pk.fft_with_precomputation_power_of_2(mont_array(f, [1] * 8), pre)
    with pytest.raises(pk.PlonkyPanic):           # two-adicity, field.rs:430
        pk.fft_precompute(po.TWEEDLEDUM_BASE.fid, 1 << 34)
    with pytest.raises(ValueError):
        pk.fft_precompute(17, 8)


@pytest.mark.parametrize("name", ["TweedledeeBase", "TweedledumBase"])
def test_padded_and_batched(name):
    f = po.FIELDS[name]
    n_in, size, k = 300, 512, 5
    rows = [rand_scalars(f, 40 + i, n_in) for i in range(k)]
    pre = pk.fft_precompute(f.fid, n_in)
    assert pre.size() == size
    got = pk.fft_batch(np.stack([mont_array(f, r) for r in rows]), pre)
    for i in range(k):
        assert canon_list(f, got[i]) == po.fft_padded(f, rows[i])
    back = pk.fft_batch(got, pre, inverse=True)
    for i in range(k):
        assert canon_list(f, back[i]) == rows[i] + [0] * (size - n_in)
    # zero-pad LDE n -> 8n (polynomials_to_values_padded, plonk_util.rs:179-190)
    pre8 = pk.fft_precompute(f.fid, 8 * 256)
    c = rand_scalars(f, 77, 256)
    got8 = pk.fft_batch(mont_array(f, c)[None], pre8)[0]
    assert canon_list(f, got8) == po.ntt(f, c + [0] * (7 * 256))


@pytest.mark.parametrize("name", ["TweedledeeBase", "TweedledumBase", "Bls12377Scalar"])
@pytest.mark.parametrize("sizes", [(1, 1), (5, 8), (64, 512), (1000, 8192), (1 << 13, 1 << 16)])
def test_coset_lde_and_back(name, sizes):
    f = po.FIELDS[name]
    n_in, size = sizes
    c = rand_scalars(f, 99 + n_in, n_in)
    pre = pk.fft_precompute(f.fid, size)
    got = pk.coset_lde(mont_array(f, c), pre)
    if size <= 8192:
        assert canon_list(f, got) == po.coset_lde(f, c, size)
    else:                                          # evaluate at a few points of g*H directly
        w = f.primitive_root_of_unity(po.log2_strict(size))
        vals = canon_list(f, got)
        for k in (0, 1, 2, size // 2 + 3, size - 1):
            x = f.generator * pow(w, k, f.p) % f.p
            acc = 0
            for cc in reversed(c):
                acc = (acc * x + cc) % f.p
            assert vals[k] == acc
    back = pk.coset_ifft(got, pre)
    assert canon_list(f, back) == c + [0] * (size - n_in)
    # explicit shift
    s = 7
    got7 = pk.coset_lde(mont_array(f, c), pre, shift=mont_array(f, [s])[0])
    if size <= 512:
        assert canon_list(f, got7) == po.coset_lde(f, c, size, shift=s)
    assert canon_list(f, pk.coset_ifft(got7, pre, shift=mont_array(f, [s])[0])) == c + [0] * (size - n_in)


@pytest.mark.parametrize("name", ["TweedledeeBase", "TweedledumBase"])
def test_divide_by_z_h(name):
    """polynomial.rs:330-380: (a * Z_H) / Z_H == a, and equality with the big-int restatement."""
    f = po.FIELDS[name]
    n = 64
    a = rand_scalars(f, 5, 7 * n - 3)
    prod = [0] * (len(a) + n)
    for i, c in enumerate(a):
        prod[i + n] = (prod[i + n] + c) % f.p
        prod[i] = (prod[i] - c) % f.p
    pre = pk.fft_precompute(f.fid, 8 * n)
    got = canon_list(f, pk.divide_by_z_h(mont_array(f, prod), n, pre))
    assert got == po.divide_by_z_h(f, prod, n)
    assert got[:len(a)] == a and not any(got[len(a):])


# ---------------------------------------------------------------- MSM
def test_msm_reference_case():
    """curve_msm.rs:218-241 (test_msm) on the device."""
    e = K["test_msm"]
    c = po.CURVES[e["curve"]]
    G = c.gen
    gens = [G, c.double(G), c.add(G, c.double(G))]
    scalars = [c.scalar.from_limbs([int(v) for v in s]) for s in e["scalars_canonical"]]
    xyz, zero = proj_from_affine(c, gens)
    pre = pk.msm_precompute(c.cid, xyz, e["w"], zero)
    out, oz = pk.msm_execute(pre, mont_array(c.scalar, scalars))
    assert result_point(c, out, oz) == c.msm_naive(scalars, gens)
    with pytest.raises(pk.PlonkyPanic):           # assert_eq!, curve_msm.rs:67
        pk.msm_execute(pre, mont_array(c.scalar, scalars[:2]))


@pytest.mark.parametrize("name", list(po.CURVES))
def test_msm_edge_cases(name):
    """Edge set of SURVEY.md 8(d): scalar 0, 1, q-1, 2^k; repeated points, P next to -P, identity,
    all-equal points (G, 2G, 3G pattern); non-trivial z on input."""
    c = po.CURVES[name]
    q = c.scalar.p
    G = c.gen
    rng = po.SplitMix64(31 + c.cid)
    base = po.rand_points(c, rng, 8)
    pts = [G, c.double(G), c.mul(3, G), G, G, base[0], c.neg(base[0]), None, base[1], base[1], base[2], None] + base[3:]
    scalars = [0, 1, q - 1, 1 << 17, (1 << 254) % q, 5, 5, 12345, q - 2, 2, (1 << 128) - 1, 0] + rand_scalars(c.scalar, 8, len(base) - 3)
    assert len(pts) == len(scalars)
    want = c.msm_naive(scalars, pts)
    xyz, zero = proj_from_affine(c, pts)
    # scale some inputs to a non-trivial projective z: (x z, y z, z)
    f = c.base
    for i in (1, 5, 8):
        z = 0xABCDEF0123456789 + i
        x, y, _ = [f.from_mont(v) for v in limbs_to_ints(xyz[i])]
        xyz[i] = mont_array(f, [x * z % f.p, y * z % f.p, z])
    for w in (4, 11):
        pre = pk.msm_precompute(c.cid, xyz, w, zero)
        out, oz = pk.msm_execute_parallel(pre, mont_array(c.scalar, scalars))
        assert result_point(c, out, oz) == want
    out, oz = pk.msm_parallel(c.cid, mont_array(c.scalar, scalars), xyz, 8, zero)
    assert result_point(c, out, oz) == want
    # cancellation to the identity and the empty MSM
    pre = pk.msm_precompute(c.cid, xyz[5:7], 8, zero[5:7])
    out, oz = pk.msm_execute(pre, mont_array(c.scalar, [9, 9]))
    assert oz and result_point(c, out, oz) is None
    pre0 = pk.msm_precompute(c.cid, np.zeros((0, 3, f.limbs), dtype=np.uint64), 8)
    out, oz = pk.msm_execute(pre0, np.zeros((0, 4), dtype=np.uint64))
    assert oz


@pytest.mark.parametrize("name,n", [("Tweedledee", 4096), ("Tweedledum", 1500), ("Bls12377", 1024)])
def test_msm_config1_matches_reference_port(name, n):
    """BASELINE config 1: 2^12 random points/scalars; GPU == C++ restatement of msm_execute_parallel
    (w = 11, the prover's window, circuit_builder.rs:1131) == big-int Pippenger."""
    c = po.CURVES[name]
    xy = rp.gen_points(c.cid, 0x504C4B59 + 1, n)          # [k_i] G, validated against big ints in the CPU suite
    scalars = rand_scalars(c.scalar, 0x504C4B59 + 2, n)
    S = mont_array(c.scalar, scalars)
    ref_table = rp.MsmTable(c.cid, xy, None, 11)
    ref_out, ref_zero = ref_table.execute(S, parallel=True)
    pre = pk.msm_precompute_affine(c.cid, xy, 11)
    out, oz = pk.pedersen_hash(S, pre)
    assert oz == ref_zero
    assert np.array_equal(out[:2], ref_out)
    # closed form: sum s_i [k_i] G = [sum s_i k_i] G
    ksum = sum(s * splitmix_hash(0x504C4B59 + 1 + i) for i, s in enumerate(scalars)) % c.scalar.p
    assert result_point(c, out, oz) == c.mul(ksum, c.gen)
    # device generator == oracle generator
    assert np.array_equal(pk.points_generate(c.cid, 0x504C4B59 + 1, 64), xy[:64])


def test_msm_batch_and_linearity():
    c = po.TWEEDLEDEE
    n, k = 777, 3
    xy = rp.gen_points(c.cid, 5, n)
    pre = pk.msm_precompute_affine(c.cid, xy, 11)
    rows = [rand_scalars(c.scalar, 60 + i, n) for i in range(k)]
    out, oz = pk.msm_execute_batch(pre, np.stack([mont_array(c.scalar, r) for r in rows]))
    pts = []
    for i in range(k):
        single, sz = pk.msm_execute(pre, mont_array(c.scalar, rows[i]))
        assert np.array_equal(single, out[i]) and sz == bool(oz[i])
        pts.append(result_point(c, out[i], oz[i]))
    # linearity: msm(a + b) == msm(a) + msm(b)
    ab = [(x + y) % c.scalar.p for x, y in zip(rows[0], rows[1])]
    s_out, s_z = pk.msm_execute(pre, mont_array(c.scalar, ab))
    assert result_point(c, s_out, s_z) == c.add(pts[0], pts[1])


def test_msm_batch_dev_concurrent_streams():
    """k executes forked onto the table's internal streams (per-stream scratch) give the same points as k
    sequential executes, and executes issued from two torch streams against one table do not interfere."""
    import torch
    from plonky_b200 import distributed as pkd
    c = po.TWEEDLEDEE
    n, k = 5000, 7
    xy = torch.from_numpy(rp.gen_points(c.cid, 21, n).view(np.int64)).cuda()
    pre = pkd.msm_precompute_affine_dev(c.cid, xy, 11)
    rows = np.stack([mont_array(c.scalar, rand_scalars(c.scalar, 300 + i, n)) for i in range(k)])
    S = torch.from_numpy(rows.view(np.int64)).cuda()
    out = torch.zeros((k, 3, 4), dtype=torch.int64, device="cuda")
    oz = torch.zeros(16, dtype=torch.uint8, device="cuda")
    pkd.msm_execute_batch_dev(pre, S, out, oz)
    torch.cuda.synchronize()
    single = torch.zeros((k, 3, 4), dtype=torch.int64, device="cuda")
    sz = torch.zeros((k, 8), dtype=torch.uint8, device="cuda")
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for i in range(k):
        with torch.cuda.stream(streams[i % 2]):
            pkd.msm_execute_dev(pre, S[i], single[i], sz[i])
    torch.cuda.synchronize()
    assert torch.equal(out, single)
    for i in range(k):
        want, wz = pk.msm_execute(pre, rows[i])
        assert np.array_equal(out[i].cpu().numpy().view(np.uint64), want)


@pytest.mark.parametrize("name,n,k", [("Tweedledee", 1000, 19), ("Tweedledum", 4097, 5), ("Bls12377", 300, 4), ("Tweedledee", 1 << 14, 9)])
def test_msm_batch_merged_pipeline(name, n, k):
    """Batches through run_batch (csrc/msm.cu): the default fork/join over side streams here, and -- re-run by
    test_gpu_optin_paths.py with PLK_MSM_BATCH_MERGE -- the merged pipeline (up to 16 short vectors as ONE sort / accumulate /
    reduce with a bucket set per vector; k = 19 is one merged execute of 16 and one of 3): every vector's point must equal its own msm_execute,
    including an all-zero vector (identity flag), a vector of equal scalars (every term of a window in one bucket: the
    big-bucket path inside one set) and a vector with a single non-zero scalar."""
    c = po.CURVES[name]
    xy = rp.gen_points(c.cid, 17, n)
    pre = pk.msm_precompute_affine(c.cid, xy, 11)
    rows = [rand_scalars(c.scalar, 900 + i, n) for i in range(k)]
    rows[1] = [0] * n
    rows[2] = [c.scalar.p - 2] * n
    rows[3] = [0] * (n - 1) + [5]
    out, oz = pk.msm_execute_batch(pre, np.stack([mont_array(c.scalar, r) for r in rows]))
    assert bool(oz[1]) and not out[1].any()
    for i in range(k):
        single, sz = pk.msm_execute(pre, mont_array(c.scalar, rows[i]))
        assert np.array_equal(single, out[i]) and sz == bool(oz[i]), i
    # one vector against the big-int oracle as well (the single execute is itself checked elsewhere)
    assert result_point(c, out[3], oz[3]) == c.mul(5, array_to_point(c, xy[n - 1], 0))


def test_msm_skewed_scalars():
    """All scalars equal / tiny: every term lands in the same bucket (worst-case load balance) and
    repeated identical points force the doubling branch."""
    c = po.TWEEDLEDUM
    n = 3000
    xy = rp.gen_points(c.cid, 9, n)
    xy[1::2] = xy[0]                                   # half the points identical
    pre = pk.msm_precompute_affine(c.cid, xy, 11)
    for s in (1, 2, c.scalar.p - 1, 0x8000):
        out, oz = pk.msm_execute(pre, mont_array(c.scalar, [s] * n))
        ref_out, ref_zero = rp.MsmTable(c.cid, xy, None, 8).execute(mont_array(c.scalar, [s] * n))
        assert oz == ref_zero and np.array_equal(out[:2], ref_out)


def test_batch_to_affine():
    c = po.TWEEDLEDEE
    f = c.base
    rng = po.SplitMix64(4)
    pts = po.rand_points(c, rng, 20) + [None]
    xyz, zero = proj_from_affine(c, pts)
    for i in range(0, 20, 2):
        z = 1234567 + i
        x, y, _ = [f.from_mont(v) for v in limbs_to_ints(xyz[i])]
        xyz[i] = mont_array(f, [x * z % f.p, y * z % f.p, z])
    out, oz = pk.batch_to_affine(c.cid, xyz, zero)
    for i, P in enumerate(pts):
        assert array_to_point(c, out[i], oz[i]) == P


@pytest.mark.parametrize("name", list(po.CURVES))
def test_affine_summations(name):
    """curve_summations.rs:164-184 cases plus longer lists, several lists at once."""
    c = po.CURVES[name]
    G = c.gen
    G2 = c.double(G)
    rng = po.SplitMix64(3 + c.cid)
    base = po.rand_points(c, rng, 9)
    lists = [[G, G], [G, G2], [G, G, G], [], [G, c.neg(G)], [None, G], base * 30 + [c.neg(base[0])], [G] * 257]
    arrays, zeros, want = [], [], []
    for pts in lists:
        xy, z = points_to_array(c, pts)
        arrays.append(xy)
        zeros.append(z)
        acc = None
        for P in pts:
            acc = c.add(acc, P)
        want.append(acc)
    out, oz = pk.affine_multisummation_best(c.cid, arrays, zeros)
    for i in range(len(lists)):
        assert result_point(c, out[i], oz[i]) == want[i], i
    one, onez = pk.affine_summation_best(c.cid, arrays[6], zeros[6])
    assert result_point(c, one, onez) == want[6]


@pytest.mark.parametrize("name", list(po.CURVES))
def test_curve_mul(name):
    """test_g1_multiplication style (bls12_377_curve.rs:49-62): 10 * G == G + ... + G; random pairs vs big ints."""
    c = po.CURVES[name]
    q = c.scalar.p
    rng = po.SplitMix64(8 + c.cid)
    pts = po.rand_points(c, rng, 6) + [c.gen, c.gen, None]
    scalars = rand_scalars(c.scalar, 15, 6) + [10, q - 1, 12345]
    xyz, zero = proj_from_affine(c, pts)
    out, oz = pk.curve_mul(c.cid, xyz, mont_array(c.scalar, scalars), zero)
    for i, (s, P) in enumerate(zip(scalars, pts)):
        assert result_point(c, out[i], oz[i]) == c.mul(s, P)
    ten = None
    for _ in range(10):
        ten = c.add(ten, c.gen)
    assert result_point(c, out[6], oz[6]) == ten


@pytest.mark.parametrize("logn,world,inverse", [(12, 1, False), (16, 1, False), (16, 2, False), (18, 4, True), (20, 8, False), (13, 2, True)])
def test_domain_split_ntt_emulated_ranks(logn, world, inverse):
    """The four-step domain-split transform with the all-to-all emulated on ONE GPU: phase A per emulated rank,
    the exchange done with tensor copies, phase B per rank -- equals the single-GPU transform bit for bit."""
    import torch
    from plonky_b200.distributed import DistributedNtt, fft_dev
    f = po.TWEEDLEDEE_BASE
    n = 1 << logn
    x = torch.from_numpy(mont_array(f, rand_scalars(f, 600 + logn, min(n, 2048)) * (n // min(n, 2048))).view(np.int64)).cuda()
    x = (x.cpu().numpy().view(np.uint64))                       # vary the repeated blocks a little
    x[:, 0] ^= np.arange(n, dtype=np.uint64) & np.uint64(0xFFFF)
    X_in = torch.from_numpy(x.view(np.int64)).cuda()
    plan = pk.fft_precompute(f.fid, n)
    want = torch.empty_like(X_in)
    fft_dev(plan, X_in, want, inverse=inverse)
    ranks = [DistributedNtt(f.fid, logn, world=world, rank=r) for r in range(world)]
    sends = [ranks[r].phase_a(ranks[r].input_rows(X_in), inverse=inverse).clone() for r in range(world)]
    outs = []
    for s in range(world):
        d = ranks[s]
        # all-to-all: rank s receives block s of every rank's send buffer, ordered by source rank
        blocks = [sends[r].view(world, d.rows, d.cols, d.L)[s] for r in range(world)]
        recv = torch.cat(blocks, dim=0).contiguous()           # [j_1 global][kl]
        outs.append(d.phase_b(recv, inverse=inverse).clone())
    got = ranks[0].natural_from_outputs(outs)
    torch.cuda.synchronize()
    assert torch.equal(got, want)


def test_reentrancy_from_host_threads():
    """The reference calls this layer from rayon workers (plonk_util.rs:173-189, halo.rs:119-123): concurrent
    host threads against ONE table and ONE plan must each get the right answer."""
    import threading
    c = po.TWEEDLEDEE
    f = po.TWEEDLEDUM_BASE
    n = 3000
    xy = rp.gen_points(c.cid, 33, n)
    pre = pk.msm_precompute_affine(c.cid, xy, 11)
    plan = pk.fft_precompute(f.fid, 4096)
    jobs = []
    for i in range(6):
        S = mont_array(c.scalar, rand_scalars(c.scalar, 900 + i, n))
        X = mont_array(f, rand_scalars(f, 950 + i, 4096))
        jobs.append((S, X))
    want = [(pk.msm_execute(pre, S), pk.fft_with_precomputation_power_of_2(X, plan)) for S, X in jobs]
    got = [None] * len(jobs)

    def work(i):
        S, X = jobs[i]
        for _ in range(3):
            got[i] = (pk.msm_execute(pre, S), pk.fft_with_precomputation_power_of_2(X, plan))
    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    for i in range(len(jobs)):
        assert np.array_equal(got[i][0][0], want[i][0][0]) and got[i][0][1] == want[i][0][1]
        assert np.array_equal(got[i][1], want[i][1])


@pytest.mark.parametrize("name,n", [("Tweedledee", 5000), ("Tweedledum", 37), ("Bls12377", 1200), ("Tweedledee", 1 << 15)])
def test_msm_parallel_variable_base(name, n):
    """msm_parallel(scalars, generators, w) (curve_msm.rs:54-61) through the table-free variable-base path:
    per-window buckets + Horner over the windows; equals the closed form [sum s_i k_i] G and the port."""
    c = po.CURVES[name]
    seed = 4242
    xy = rp.gen_points(c.cid, seed, n)
    f = c.base
    xyz = np.zeros((n, 3, f.limbs), dtype=np.uint64)
    xyz[:, :2] = xy
    xyz[:, 2] = ints_to_limbs([f.R], f.limbs)[0]
    scalars = rand_scalars(c.scalar, 77, n)
    scalars[0], scalars[1 % n] = 0, c.scalar.p - 1
    S = mont_array(c.scalar, scalars)
    out, oz = pk.msm_parallel(c.cid, S, xyz, 8)
    ksum = sum(s * splitmix_hash(seed + i) for i, s in enumerate(scalars)) % c.scalar.p
    assert result_point(c, out, oz) == c.mul(ksum, c.gen)
    if n <= 5000:
        ref_out, ref_zero = rp.MsmTable(c.cid, xy, None, 8).execute(S, parallel=True)
        assert oz == ref_zero and np.array_equal(out[:2], ref_out)
