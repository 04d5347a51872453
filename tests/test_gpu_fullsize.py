"""BASELINE.json full sizes on the GPU, checked through size-independent properties (the oracles
cannot finish 2^20 / 2^24 in seconds):
  MSM 2^20  generators P_i = [k_i] G with known k_i  =>  sum s_i P_i == [sum s_i k_i mod q] G (one
            big-int scalar multiplication), plus linearity and a zero-scalar / identity mix;
  NTT 2^24  INTT(NTT(x)) == x bit for bit, linearity NTT(x + y) == NTT(x) + NTT(y), spot values
            X[k] = sum_j x_j w^(jk) for sparse inputs, coset LDE == coset evaluation at sample points."""
import numpy as np
import pytest

import plonky_oracle as po
import plonky_b200 as pk
from helpers import mont_array, canon_list, limbs_to_ints, ints_to_limbs, splitmix_hash

pytestmark = pytest.mark.gpu


def rand_limbs(n, seed, limbs=4):
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, 1 << 63, size=(n, limbs), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, limbs), dtype=np.uint64)
    a[:, limbs - 1] >>= np.uint64(2)          # < 2^254 < p for the Tweedle fields
    return a


def point_of(curve, out, oz):
    f = curve.base
    if oz:
        return None
    x, y, z = limbs_to_ints(out)
    assert z == f.R
    return (f.from_mont(x), f.from_mont(y))


def limbs_to_pyints(a):
    """(n, 4) uint64 -> python ints, vectorised per limb."""
    a = np.asarray(a, dtype=np.uint64)
    cols = [a[:, j].astype(object) for j in range(a.shape[1])]
    return [int(c0) + (int(c1) << 64) + (int(c2) << 128) + (int(c3) << 192) for c0, c1, c2, c3 in zip(*cols)]


@pytest.mark.parametrize("name,logn", [("Tweedledee", 20), ("Tweedledum", 18), ("Bls12377", 17)])
def test_msm_full_size_closed_form(name, logn):
    c = po.CURVES[name]
    q = c.scalar.p
    n = 1 << logn
    seed = 0x504C4B59 + 1
    xy = pk.points_generate(c.cid, seed, n)
    pre = pk.msm_precompute_affine(c.cid, xy, 11)
    S = rand_limbs(n, 5)
    if name == "Bls12377":
        S[:, 3] >>= np.uint64(2)                 # 253-bit field: keep the raw pattern below r
    S[::1000] = 0                                # some zero scalars
    out, oz = pk.msm_execute_parallel(pre, S)
    rinv = pow(c.scalar.R, -1, q)
    svals = limbs_to_pyints(S)                   # Montgomery patterns -> canonical = s * R^-1
    ksum = 0
    for i, s in enumerate(svals):
        ksum += s * splitmix_hash(seed + i)
    ksum = ksum * rinv % q
    assert point_of(c, out, oz) == c.mul(ksum, c.gen)
    # linearity on the full size: msm(2 s) == 2 msm(s)
    S2 = mont_array(c.scalar, [2])                # multiply every scalar by 2 on device
    two = np.broadcast_to(S2, S.shape).copy()
    Sd = pk.field_op(c.scalar.fid, "mul", S, two)
    out2, oz2 = pk.msm_execute(pre, Sd)
    assert point_of(c, out2, oz2) == c.double(point_of(c, out, oz))


def test_ntt_full_size_properties():
    f = po.TWEEDLEDEE_BASE
    logn = 24
    n = 1 << logn
    pre = pk.fft_precompute(f.fid, n)
    x = rand_limbs(n, 11)
    X = pk.fft_with_precomputation_power_of_2(x, pre)
    back = pk.ifft_with_precomputation_power_of_2(X, pre)
    assert np.array_equal(back, x)                                     # round trip, bit exact
    # sparse input: X[k] = sum over the few non-zero j
    sp = np.zeros((n, 4), dtype=np.uint64)
    idx = [0, 1, 12345, n // 2 + 7, n - 1]
    vals = [3, 5, 7, 11, 13]
    for j, v in zip(idx, vals):
        sp[j] = mont_array(f, [v])[0]
    SP = pk.fft_with_precomputation_power_of_2(sp, pre)
    w = f.primitive_root_of_unity(logn)
    for k in (0, 1, 2, 255, 256, 65535, 65536, n // 3, n - 1):
        want = sum(v * pow(w, j * k, f.p) for j, v in zip(idx, vals)) % f.p
        assert canon_list(f, SP[k:k + 1]) == [want]
    # linearity: NTT(x + sp) == NTT(x) + NTT(sp)
    xs = pk.field_op(f.fid, "add", x, sp)
    XS = pk.fft_with_precomputation_power_of_2(xs, pre)
    assert np.array_equal(XS, pk.field_op(f.fid, "add", X, SP))


def test_coset_lde_full_size():
    """config 3: 2^21 coefficients -> 2^24 evaluations on g*H (fused shift + zero-pad)."""
    f = po.TWEEDLEDEE_BASE
    n_in, size = 1 << 21, 1 << 24
    pre = pk.fft_precompute(f.fid, size)
    c = np.zeros((n_in, 4), dtype=np.uint64)
    idx = [0, 1, 2, 1000, n_in - 1]
    vals = [1, 2, 3, 4, 5]
    for j, v in zip(idx, vals):
        c[j] = mont_array(f, [v])[0]
    E = pk.coset_lde(c, pre)
    w = f.primitive_root_of_unity(24)
    g = f.generator
    for k in (0, 1, 77, size // 2, size - 1):
        xk = g * pow(w, k, f.p) % f.p
        want = sum(v * pow(xk, j, f.p) for j, v in zip(idx, vals)) % f.p
        assert canon_list(f, E[k:k + 1]) == [want]
    back = pk.coset_ifft(E, pre)
    assert np.array_equal(back[:n_in], c) and not back[n_in:].any()
    # dense random coefficients: LDE followed by the inverse gives the padded input back
    cr = rand_limbs(n_in, 21)
    back = pk.coset_ifft(pk.coset_lde(cr, pre), pre)
    assert np.array_equal(back[:n_in], cr) and not back[n_in:].any()


# ---- dense comparisons at BASELINE's full sizes against the C++ restatement of the reference ----------------------
def test_ntt_full_size_dense_vs_port():
    """NTT 2^24 forward and inverse, every output word against fft_with_precomputation_power_of_2 /
    ifft_with_precomputation_power_of_2 (fft.rs:82-156) restated in oracle/ref_port.cpp."""
    import ref_port as rp
    f = po.TWEEDLEDEE_BASE
    n = 1 << 24
    x = rand_limbs(n, 2024)
    pre = pk.fft_precompute(f.fid, n)
    port = rp.FftPlan(f.fid, n)
    got = pk.fft_with_precomputation_power_of_2(x, pre)
    assert np.array_equal(got, port.run(x))
    got_inv = pk.ifft_with_precomputation_power_of_2(x, pre)
    assert np.array_equal(got_inv, port.run(x, inverse=True))


@pytest.mark.parametrize("name,logn", [("Tweedledee", 20), ("Bls12377", 19)])
def test_msm_full_size_pedersen_g_vs_port(name, logn):
    """bench.py's actual input -- pedersen_g = blake_hash_usize_to_curve(i) (circuit_builder.rs:1127), 2^20 terms
    (BLS12-377: 2^19, the per-GPU shard of BASELINE config 4) -- against msm_execute_parallel with w = 11
    (curve_msm.rs:102-157) restated in oracle/ref_port.cpp."""
    import ref_port as rp
    c = po.CURVES[name]
    n = 1 << logn
    g = pk.blake_hash_usize_to_curve(c.cid, 0, n)
    S = rand_limbs(n, 77)
    if name == "Bls12377":
        S[:, 3] >>= np.uint64(2)
    out, oz = pk.msm_execute_parallel(pk.msm_precompute_affine(c.cid, g, 11), S)
    want_xy, want_zero = rp.MsmTable(c.cid, g, None, 11).execute(S, parallel=True)
    assert oz == want_zero and np.array_equal(out[:2], want_xy)


def test_ntt_four_pass_size_properties():
    """2^25 over TweedledumBase (the prover's field for Circuit<Tweedledee>): the smallest size that needs FOUR passes
    (digits 7, 6, 6, 6).  Round trip bit for bit, spot values of a sparse input against big integers, linearity."""
    f = po.TWEEDLEDUM_BASE
    logn = 25
    n = 1 << logn
    pre = pk.fft_precompute(f.fid, n)
    x = rand_limbs(n, 25)
    X = pk.fft_with_precomputation_power_of_2(x, pre)
    assert np.array_equal(pk.ifft_with_precomputation_power_of_2(X, pre), x)
    sp = np.zeros((n, 4), dtype=np.uint64)
    idx = [0, 1, 127, 128, 8191, 8192, n // 2 + 3, n - 1]          # digit borders of the 7 / 6 / 6 / 6 split
    vals = [3, 5, 7, 11, 13, 17, 19, 23]
    for j, v in zip(idx, vals):
        sp[j] = mont_array(f, [v])[0]
    SP = pk.fft_with_precomputation_power_of_2(sp, pre)
    w = f.primitive_root_of_unity(logn)
    for k in (0, 1, 63, 64, 4095, 4096, 262143, 262144, n // 3, n - 1):
        want = sum(v * pow(w, j * k, f.p) for j, v in zip(idx, vals)) % f.p
        assert canon_list(f, SP[k:k + 1]) == [want]
    XS = pk.fft_with_precomputation_power_of_2(pk.field_op(f.fid, "add", x, sp), pre)
    assert np.array_equal(XS, pk.field_op(f.fid, "add", X, SP))
