"""Edge cases of the hot path on the GPU (empty / tiny / ragged inputs, extreme values, collisions), each against
the big-integer oracle.  Mirrors the corner cases the reference's own tests poke at (zero points, G + G,
G + (-G), empty sums, non power-of-two lengths) and adds the ones a bucket method is sensitive to."""
import random

import numpy as np
import pytest

import plonky_oracle as po
import plonky_b200 as pk
from helpers import ints_to_limbs, limbs_to_ints, mont_array, canon_list, points_to_array, rand_scalars, array_to_point

pytestmark = pytest.mark.gpu


def proj(curve, pts):
    f = curve.base
    xy, zero = points_to_array(curve, pts)
    xyz = np.zeros((len(pts), 3, f.limbs), dtype=np.uint64)
    xyz[:, :2] = xy
    one = ints_to_limbs([f.R], f.limbs)[0]
    for i in range(len(pts)):
        if not zero[i]:
            xyz[i, 2] = one
    return xyz, zero


def point_of(curve, out, oz):
    f = curve.base
    if oz:
        assert not np.asarray(out).any()
        return None
    x, y, z = limbs_to_ints(out)
    assert z == f.R
    return (f.from_mont(x), f.from_mont(y))


@pytest.mark.parametrize("name", list(po.CURVES))
def test_msm_tiny_and_adversarial(name):
    """Random small MSMs whose scalars and points are drawn from adversarial pools: 0, 1, 2, q-1, q-2, powers of
    two and digit-boundary values (2^15, 2^16 - 1, 2^16, ...: carries of the signed recoding); points G, 2G, -G,
    the identity and repeats, so that doubling / cancellation branches fire inside buckets."""
    c = po.CURVES[name]
    q = c.scalar.p
    G = c.gen
    rnd = random.Random(1234 + c.cid)
    rng = po.SplitMix64(5 + c.cid)
    pool_pts = [G, c.double(G), c.neg(G), None, c.mul(7, G)] + po.rand_points(c, rng, 3)
    pool_s = [0, 1, 2, q - 1, q - 2, 1 << 15, (1 << 15) + 1, (1 << 16) - 1, 1 << 16, (1 << 16) + 1, (1 << 31) - 1, 1 << 32,
              (1 << 254) % q, (1 << 200) - 1, 0x8000800080008000800080008000 % q, 0x7FFF7FFF7FFF7FFF7FFF7FFF % q]
    for trial in range(14):
        n = rnd.choice([1, 2, 3, 5, 8, 17, 40])
        pts = [rnd.choice(pool_pts) for _ in range(n)]
        scalars = [rnd.choice(pool_s) if rnd.random() < 0.7 else rnd.randrange(q) for _ in range(n)]
        want = c.msm_naive(scalars, pts)
        xyz, zero = proj(c, pts)
        S = mont_array(c.scalar, scalars)
        pre = pk.msm_precompute(c.cid, xyz, 4 + trial % 9, zero)
        out, oz = pk.msm_execute(pre, S)
        assert point_of(c, out, oz) == want, (trial, n)
        out, oz = pk.msm_parallel(c.cid, S, xyz, 8, zero)
        assert point_of(c, out, oz) == want, ("variable", trial, n)


def test_msm_all_identity_and_all_zero_scalars():
    c = po.TWEEDLEDEE
    n = 300
    xyz, zero = proj(c, [None] * n)
    pre = pk.msm_precompute(c.cid, xyz, 11, zero)
    out, oz = pk.msm_execute(pre, mont_array(c.scalar, list(range(1, n + 1))))
    assert oz and not out.any()
    rng = po.SplitMix64(3)
    xyz, zero = proj(c, po.rand_points(c, rng, 5) * 60)
    pre = pk.msm_precompute(c.cid, xyz, 11, zero)
    out, oz = pk.msm_execute(pre, np.zeros((n, 4), dtype=np.uint64))
    assert oz and not out.any()


@pytest.mark.parametrize("name", ["TweedledeeBase", "TweedledumBase", "Bls12377Scalar", "Bls12377Base"])
def test_ntt_extreme_values(name):
    f = po.FIELDS[name]
    for n in (1, 2, 4, 64, 512):
        pre = pk.fft_precompute(f.fid, n)
        for vals in ([0] * n, [f.p - 1] * n, [1] + [0] * (n - 1), [0] * (n - 1) + [f.p - 1], list(range(n))):
            got = pk.fft_with_precomputation_power_of_2(mont_array(f, vals), pre)
            assert canon_list(f, got) == po.ntt(f, vals)
            assert canon_list(f, pk.ifft_with_precomputation_power_of_2(got, pre)) == vals


def test_ragged_and_padded_lengths():
    """fft_with_precomputation (fft.rs:61-80) for every input length 1..33 against zero-padded big-int DFTs;
    fft() builds its own plan (fft.rs:42-45)."""
    f = po.TWEEDLEDUM_BASE
    for n_in in list(range(1, 34)) + [255, 257]:
        c = [(7 * i + 3) % f.p for i in range(n_in)]
        assert canon_list(f, pk.fft(f.fid, mont_array(f, c))) == po.fft_padded(f, c)
    pre = pk.fft_precompute(f.fid, 64)
    with pytest.raises(pk.PlonkyPanic):                 # 17 coefficients pad to 32, not to the plan's 64
        pk.fft_with_precomputation(mont_array(f, [1] * 17), pre)
    got = pk.fft_batch(mont_array(f, [5])[None], pre)[0]   # a constant polynomial evaluates to itself everywhere
    assert canon_list(f, got) == [5] * 64


def test_coset_lde_without_padding_and_zero_polynomial():
    f = po.TWEEDLEDEE_BASE
    size = 256
    pre = pk.fft_precompute(f.fid, size)
    c = [(i * i + 1) % f.p for i in range(size)]
    assert canon_list(f, pk.coset_lde(mont_array(f, c), pre)) == po.coset_lde(f, c, size)
    z = pk.coset_lde(np.zeros((10, 4), dtype=np.uint64), pre)
    assert not z.any()
    assert not pk.divide_by_z_h(np.zeros((size, 4), dtype=np.uint64), 32, pre).any()


@pytest.mark.parametrize("ratio", [2, 4, 8])
def test_divide_by_z_h_ratios(ratio):
    """Z_H = X^n - 1 on a coset of size ratio * n: the denominators take `ratio` distinct values (polynomial.rs:351-361)."""
    f = po.TWEEDLEDUM_BASE
    n = 32
    size = ratio * n
    a = [(3 * i + 1) % f.p for i in range(size - n - 2)]
    prod = [0] * (len(a) + n)
    for i, cf in enumerate(a):
        prod[i + n] = (prod[i + n] + cf) % f.p
        prod[i] = (prod[i] - cf) % f.p
    pre = pk.fft_precompute(f.fid, size)
    got = canon_list(f, pk.divide_by_z_h(mont_array(f, prod), n, pre))
    assert got[:len(a)] == a and not any(got[len(a):])
    assert got == po.divide_by_z_h(f, prod, n)


def test_batch_to_affine_unflagged_zero_z():
    """ProjectivePoint with z = 0 but zero = false cannot be produced by the reference's constructors; the device
    maps it to the identity instead of dividing by zero."""
    c = po.TWEEDLEDEE
    xyz = np.zeros((2, 3, 4), dtype=np.uint64)
    xyz[0] = proj(c, [c.gen])[0][0]
    out, oz = pk.batch_to_affine(c.cid, xyz, np.array([0, 1], dtype=np.uint8))
    assert not oz[0] and oz[1]
    f = c.base
    assert (f.from_mont(limbs_to_ints(out[0])[0]), f.from_mont(limbs_to_ints(out[0])[1])) == c.gen


def test_invalid_arguments_return_errors_not_crashes():
    with pytest.raises(ValueError):
        pk.msm_precompute(9, np.zeros((1, 3, 4), dtype=np.uint64), 8)
    with pytest.raises(ValueError):
        pk.msm_precompute(pk.TWEEDLEDEE, np.zeros((1, 3, 4), dtype=np.uint64), 0)
    with pytest.raises(ValueError):
        pk.field_op(7, "add", np.zeros((1, 4), dtype=np.uint64), np.zeros((1, 4), dtype=np.uint64))
    c = po.TWEEDLEDEE
    pre = pk.msm_precompute(c.cid, proj(c, [c.gen, c.gen])[0], 8)
    with pytest.raises(pk.PlonkyPanic):
        pk.msm_execute(pre, np.zeros((3, 4), dtype=np.uint64))
    with pytest.raises(pk.PlonkyPanic):
        pk.msm_parallel(c.cid, np.zeros((3, 4), dtype=np.uint64), proj(c, [c.gen, c.gen])[0], 8)


@pytest.mark.parametrize("name", list(po.FIELDS))
def test_polynomial_mul(name):
    """Polynomial::mul (polynomial.rs:209-227) against the schoolbook product, including zero operands, trailing zero
    coefficients (degree < length) and the un-trimmed power-of-two output length."""
    f = po.FIELDS[name]
    cases = [([1], [1]), ([0], [5, 6]), ([], [1]), ([3, 0, 0], [0, 0, 7, 0]), (rand_scalars(f, 1, 5), rand_scalars(f, 2, 7)),
             (rand_scalars(f, 3, 200), rand_scalars(f, 4, 313)), ([f.p - 1] * 33, [f.p - 1] * 32), (rand_scalars(f, 5, 1024) + [0, 0], [2])]
    for a, b in cases:
        got = pk.polynomial_mul(f.fid, mont_array(f, a) if a else np.zeros((0, f.limbs), dtype=np.uint64),
                                mont_array(f, b) if b else np.zeros((0, f.limbs), dtype=np.uint64))
        assert canon_list(f, got) == po.poly_mul(f, a, b)


@pytest.mark.parametrize("name,n", [("TweedledeeBase", 1), ("TweedledeeBase", 2), ("TweedledumBase", 700), ("TweedledumBase", 4096), ("Bls12377Scalar", 1500)])
def test_permutation_polynomial(name, n):
    """permutation_polynomial (plonk_util.rs:233-262): Z on the subgroup, 6 routed of 9 wires, sigma read with stride 8,
    against the sequential big-integer restatement; a zero denominator panics like the reference's division."""
    f = po.FIELDS[name]
    routed, wires, stride = 6, 9, 8
    sub = rand_scalars(f, 41, n)
    w = [rand_scalars(f, 100 + i, wires) for i in range(n)]
    sig = [rand_scalars(f, 5000 + j, stride * n) for j in range(routed)]
    k_is = rand_scalars(f, 77, routed)
    beta, gamma = rand_scalars(f, 78, 2)
    want = po.permutation_polynomial(f, sub, w, sig, k_is, beta, gamma)
    W = np.stack([mont_array(f, row) for row in w]) if n else np.zeros((0, wires, f.limbs), dtype=np.uint64)
    S = np.stack([mont_array(f, row) for row in sig])
    got = pk.permutation_polynomial(f.fid, mont_array(f, sub), W, S, mont_array(f, k_is), mont_array(f, [beta])[0], mont_array(f, [gamma])[0])
    assert canon_list(f, got) == want
    if n >= 700:
        # force w + beta sigma + gamma == 0 at gate 5, wire 2
        w[5][2] = (-(beta * sig[2][stride * 5] + gamma)) % f.p
        W = np.stack([mont_array(f, row) for row in w])
        with pytest.raises(pk.PlonkyPanic):
            pk.permutation_polynomial(f.fid, mont_array(f, sub), W, S, mont_array(f, k_is), mont_array(f, [beta])[0], mont_array(f, [gamma])[0])


def test_msm_generator_of_order_two():
    """BLS12-377 G1 has an even cofactor: T = (-1, 0) lies on y^2 = x^3 + 1 and 2 T = O.  Every higher power of T in the
    fixed-base table is the identity; the reference handles arbitrary curve points (curve.rs:234-260 doubles through
    y = 0 to ZERO)."""
    c = po.BLS12_377
    f = c.base
    T = (f.p - 1, 0)
    assert c.is_on_curve(T)
    gens = [c.gen, T, c.double(c.gen), T]
    xy, zero = points_to_array(c, gens)
    for seed in (1, 2):
        s = rand_scalars(c.scalar, seed, 4)
        s[1] |= 1                       # odd multiple of T = T
        s[3] &= ~1                      # even multiple of T = O
        want = c.add(c.add(c.mul(s[0], gens[0]), c.mul(s[2], gens[2])), T)
        for w in (11, 4):
            out, oz = pk.msm_execute(pk.msm_precompute_affine(c.cid, xy, w, zero), mont_array(c.scalar, s))
            assert not oz and array_to_point(c, out[:2], 0) == want
        out, oz = pk.msm_parallel(c.cid, mont_array(c.scalar, s), np.concatenate([xy, np.broadcast_to(mont_array(f, [1])[0], (4, 1, f.limbs))], axis=1), 4)
        assert not oz and array_to_point(c, out[:2], 0) == want
