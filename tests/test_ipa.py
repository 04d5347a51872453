"""Halo inner-product-argument rounds (src/halo.rs:63-124): the C++ restatement against the big-integer oracle on
CPU, and the device-resident rounds (plk_ipa_*) against the restatement on the GPU.  Bit-exact: scalars limb for limb,
points on the normalised affine (x, y) + zero flag."""
import numpy as np
import pytest

import plonky_oracle as po
import ref_port as rp
from helpers import limbs_to_ints, mont_array, canon_list, points_to_array, array_to_point, rand_scalars

CURVES = [po.TWEEDLEDEE, po.TWEEDLEDUM, po.BLS12_377]


def make_inputs(curve, n, seed, special=False):
    sf = curve.scalar
    a = rand_scalars(sf, seed, n)
    b = rand_scalars(sf, seed + 1, n)
    ks = [1 + (k % 97) for k in rand_scalars(sf, seed + 2, n)]
    g = [curve.mul(k, curve.gen) for k in ks]
    if special and n >= 8:
        g[1] = None                       # identity (AffinePoint::ZERO)
        g[n // 2 + 2] = g[2]              # G_lo_i == G_hi_i: doubling inside the fold
        g[n // 2 + 3] = curve.neg(g[3])   # G_lo_i == -G_hi_i
        a[0] = 0
        a[n - 1] = sf.p - 1
        b[n // 2] = 0
    return a, b, g


def challenge(curve, seed):
    sf = curve.scalar
    u = rand_scalars(sf, 1000 + seed, 1)[0] or 1
    return u, sf.inv(u)


def pack(curve, a, b, g):
    sf = curve.scalar
    xy, zero = points_to_array(curve, g)
    return mont_array(sf, a), mont_array(sf, b), xy, zero


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_port_matches_bigint_oracle(curve):
    """ref_port.cpp's restatement of halo.rs:87-123 (msm_parallel with w = 8 / w = 4, as the reference calls it)
    against plain big-integer arithmetic."""
    n = 8
    a, b, g = make_inputs(curve, n, 77, special=True)
    A, B, G, Z = pack(curve, a, b, g)
    (l, lz), (r, rz), ipl, ipr = rp.ipa_round_lr(curve.cid, A, B, G, Z)
    wl, wr, wipl, wipr = po.halo_round_lr(curve, a, b, g)
    assert array_to_point(curve, l, lz) == wl and array_to_point(curve, r, rz) == wr
    assert canon_list(curve.scalar, ipl.reshape(1, -1)) == [wipl] and canon_list(curve.scalar, ipr.reshape(1, -1)) == [wipr]
    u, u_inv = challenge(curve, 5)
    U, UI = mont_array(curve.scalar, [u])[0], mont_array(curve.scalar, [u_inv])[0]
    na, nb, ng, nz = rp.ipa_fold(curve.cid, A, B, G, Z, U, UI)
    wa, wb, wg = po.halo_fold(curve, a, b, g, u, u_inv)
    assert canon_list(curve.scalar, na) == wa and canon_list(curve.scalar, nb) == wb
    assert [array_to_point(curve, ng[i], nz[i]) for i in range(n // 2)] == wg


def test_port_rejects_non_power_of_two():
    curve = po.TWEEDLEDEE
    a, b, g = make_inputs(curve, 6, 3)
    with pytest.raises(AssertionError):
        rp.ipa_round_lr(curve.cid, *pack(curve, a, b, g))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("curve,n", [(po.TWEEDLEDEE, 64), (po.TWEEDLEDUM, 32), (po.BLS12_377, 16)], ids=lambda v: getattr(v, "name", str(v)))
def test_device_rounds_match_restatement(curve, n):
    """All log2(n) rounds on the device, each compared with the restatement run on the same state."""
    import plonky_b200 as pk
    a, b, g = make_inputs(curve, n, 11, special=True)
    A, B, G, Z = pack(curve, a, b, g)
    st = pk.HaloIpaRounds(curve.cid, A, B, G, Z)
    rnd = 0
    while len(st) > 1:
        (l, lz), (r, rz), ipl, ipr = st.round_lr()
        (wl, wlz), (wr, wrz), wipl, wipr = rp.ipa_round_lr(curve.cid, A, B, G, Z)
        assert lz == wlz and rz == wrz
        assert np.array_equal(l[:2], wl) and np.array_equal(r[:2], wr)
        if not lz:
            assert limbs_to_ints(l[2:3])[0] == curve.base.R            # normalised: z = ONE
        assert np.array_equal(ipl, wipl) and np.array_equal(ipr, wipr)
        u, u_inv = challenge(curve, rnd)
        U, UI = mont_array(curve.scalar, [u])[0], mont_array(curve.scalar, [u_inv])[0]
        st.fold(U, UI)
        A, B, G, Z = rp.ipa_fold(curve.cid, A, B, G, Z, U, UI)
        ga, gb, gg, gz = st.read()
        assert len(st) == A.shape[0]
        assert np.array_equal(ga, A) and np.array_equal(gb, B)
        assert np.array_equal(gz, Z) and np.array_equal(gg, G)
        rnd += 1
    assert rnd == po.log2_strict(n)


@pytest.mark.gpu
@pytest.mark.parametrize("curve,n", [(po.TWEEDLEDEE, 64), (po.TWEEDLEDUM, 16), (po.BLS12_377, 8)], ids=lambda v: getattr(v, "name", str(v)))
def test_table_mode_rounds_match_restatement(curve, n):
    """plk_ipa_new_with_table: G is never folded (coefficients over the original generators + fixed-base MSMs);
    every L_j, R_j, inner product, folded a / b and the final halo_g must equal the folding restatement's."""
    import plonky_b200 as pk
    a, b, g = make_inputs(curve, n, 19, special=True)
    A, B, G, Z = pack(curve, a, b, g)
    pre = pk.msm_precompute_affine(curve.cid, G, 11, zero=Z)
    st = pk.HaloIpaRounds(curve.cid, A, B, precomputation=pre)
    rnd = 0
    while len(st) > 1:
        (l, lz), (r, rz), ipl, ipr = st.round_lr()
        (wl, wlz), (wr, wrz), wipl, wipr = rp.ipa_round_lr(curve.cid, A, B, G, Z)
        assert lz == wlz and rz == wrz
        assert np.array_equal(l[:2], wl) and np.array_equal(r[:2], wr)
        assert np.array_equal(ipl, wipl) and np.array_equal(ipr, wipr)
        u, u_inv = challenge(curve, 50 + rnd)
        U, UI = mont_array(curve.scalar, [u])[0], mont_array(curve.scalar, [u_inv])[0]
        st.fold(U, UI)
        A, B, G, Z = rp.ipa_fold(curve.cid, A, B, G, Z, U, UI)
        ga, gb, _, _ = st.read(with_g=False)
        assert np.array_equal(ga, A) and np.array_equal(gb, B)
        if len(st) > 1:
            with pytest.raises(ValueError):          # the folded generators are only materialised at length 1
                st.read()
        rnd += 1
    ga, gb, gg, gz = st.read()
    assert np.array_equal(ga, A) and np.array_equal(gb, B)
    assert np.array_equal(gz, Z) and np.array_equal(gg, G)          # halo_g[0].to_affine(), halo.rs:126-127


@pytest.mark.gpu
def test_table_mode_matches_folding_mode_2p10():
    """Both device paths on the same inputs at n = 2^10 (Tweedledee, reference generator set)."""
    import plonky_b200 as pk
    curve = po.TWEEDLEDEE
    sf = curve.scalar
    n = 1 << 10
    A = mont_array(sf, rand_scalars(sf, 61, n))
    B = mont_array(sf, rand_scalars(sf, 62, n))
    G = pk.blake_hash_usize_to_curve(curve.cid, 0, n)           # pedersen_g, circuit_builder.rs:1127
    pre = pk.msm_precompute_affine(curve.cid, G, 11)
    s1 = pk.HaloIpaRounds(curve.cid, A, B, G)
    s2 = pk.HaloIpaRounds(curve.cid, A, B, precomputation=pre)
    rnd = 0
    while len(s1) > 1:
        r1, r2 = s1.round_lr(), s2.round_lr()
        for x, y in zip(r1, r2):
            if isinstance(x, tuple):
                assert x[1] == y[1] and np.array_equal(x[0], y[0])
            else:
                assert np.array_equal(x, y)
        u, u_inv = challenge(curve, 70 + rnd)
        U, UI = mont_array(sf, [u])[0], mont_array(sf, [u_inv])[0]
        s1.fold(U, UI)
        s2.fold(U, UI)
        rnd += 1
    f1, f2 = s1.read(), s2.read()
    for x, y in zip(f1, f2):
        assert np.array_equal(x, y)
    with pytest.raises(pk.PlonkyPanic):               # table length != vector length
        pk.HaloIpaRounds(curve.cid, A[:512], B[:512], precomputation=pre)


@pytest.mark.gpu
def test_device_rounds_errors():
    import plonky_b200 as pk
    curve = po.TWEEDLEDEE
    a, b, g = make_inputs(curve, 6, 3)
    with pytest.raises(pk.PlonkyPanic):                   # log2_strict(degree), halo.rs:62
        pk.HaloIpaRounds(curve.cid, *pack(curve, a, b, g))
    a, b, g = make_inputs(curve, 1, 3)
    st = pk.HaloIpaRounds(curve.cid, *pack(curve, a, b, g))
    with pytest.raises(ValueError):
        st.round_lr()
    with pytest.raises(ValueError):
        st.fold(np.ones(4, dtype=np.uint64), np.ones(4, dtype=np.uint64))
    with pytest.raises(pk.PlonkyPanic):                   # debug_assert_eq!(halo_b.len(), n), halo.rs:68
        pk.HaloIpaRounds(curve.cid, mont_array(curve.scalar, [1, 2]), mont_array(curve.scalar, [1]), pack(curve, [1, 2], [1, 2], [curve.gen, curve.gen])[2])


@pytest.mark.gpu
def test_device_round_invariant_2p12():
    """Size-independent property at a prover-like size (n = 2^12, Tweedledee):
    <a', G'> = <a, G> + u^2 <a_lo, G_hi> + u^-2 <a_hi, G_lo>   and   <a', b'> = <a, b> + u^2 <a_lo, b_hi> + u^-2 <a_hi, b_lo>."""
    import plonky_b200 as pk
    curve = po.TWEEDLEDEE
    sf = curve.scalar
    n = 1 << 12
    A = mont_array(sf, rand_scalars(sf, 21, n))
    B = mont_array(sf, rand_scalars(sf, 22, n))
    G = pk.points_generate(curve.cid, 99, n)

    def msm(sc, pts):
        xyz = np.zeros((pts.shape[0], 3, 4), dtype=np.uint64)
        xyz[:, :2] = pts
        xyz[:, 2] = np.array(sf_one_base, dtype=np.uint64)
        out, oz = pk.msm_parallel(curve.cid, sc, xyz, 8)
        return array_to_point(curve, out[:2], oz)

    from helpers import ints_to_limbs
    sf_one_base = ints_to_limbs([curve.base.R], 4)[0]
    st = pk.HaloIpaRounds(curve.cid, A, B, G)
    before = msm(A, G)
    (l, lz), (r, rz), ipl, ipr = st.round_lr()
    u, u_inv = challenge(curve, 42)
    st.fold(mont_array(sf, [u])[0], mont_array(sf, [u_inv])[0])
    a2, b2, g2, z2 = st.read()
    assert not z2.any()
    after = msm(a2, g2)
    L, R = array_to_point(curve, l[:2], lz), array_to_point(curve, r[:2], rz)
    want = curve.add(before, curve.add(curve.mul(u * u % sf.p, L), curve.mul(u_inv * u_inv % sf.p, R)))
    assert after == want
    ca, cb = canon_list(sf, A), canon_list(sf, B)
    ip0 = sum(x * y for x, y in zip(ca, cb)) % sf.p
    ip1 = sum(x * y for x, y in zip(canon_list(sf, a2), canon_list(sf, b2))) % sf.p
    cl, cr = canon_list(sf, ipl.reshape(1, -1))[0], canon_list(sf, ipr.reshape(1, -1))[0]
    assert ip1 == (ip0 + u * u * cl + u_inv * u_inv * cr) % sf.p


@pytest.mark.gpu
def test_ipa_and_generators_from_concurrent_host_threads():
    """The reference runs this path from rayon workers: four host threads, each with its own IPA state against ONE
    shared table (plus a generator derivation), must reproduce the single-threaded results."""
    import threading
    import plonky_b200 as pk
    curve = po.TWEEDLEDEE
    sf = curve.scalar
    n = 256
    G = pk.blake_hash_usize_to_curve(curve.cid, 0, n)
    pre = pk.msm_precompute_affine(curve.cid, G, 11)
    jobs = [(mont_array(sf, rand_scalars(sf, 300 + t, n)), mont_array(sf, rand_scalars(sf, 400 + t, n))) for t in range(4)]

    def run(A, B):
        st = pk.HaloIpaRounds(curve.cid, A, B, precomputation=pre)
        trace = []
        rnd = 0
        while len(st) > 1:
            (l, lz), (r, rz), ipl, ipr = st.round_lr()
            trace += [l.copy(), r.copy(), ipl.copy(), ipr.copy()]
            u, u_inv = challenge(curve, 900 + rnd)
            st.fold(mont_array(sf, [u])[0], mont_array(sf, [u_inv])[0])
            rnd += 1
        trace += list(st.read())
        trace.append(pk.blake_hash_usize_to_curve(curve.cid, 1000, 8))
        return trace

    want = [run(A, B) for A, B in jobs]
    got = [None] * len(jobs)
    errors = []

    def work(i):
        try:
            got[i] = run(*jobs[i])
        except Exception as e:      # surfaced below
            errors.append(e)
    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for w, g in zip(want, got):
        assert len(w) == len(g)
        for x, y in zip(w, g):
            assert np.array_equal(x, y)
