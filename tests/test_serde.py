"""Wire formats: field / point ToBytes, the CBOR shapes of MsmPrecomputation and
FftPrecomputation (src/serialization.rs:17-148, :254-328; src/curve/curve_msm.rs:16-25; src/fft.rs:28-34), and the one-call
blinded commitment (src/poly_commit.rs:32-66).

CPU: the CBOR encoder / decoder against hand-assembled RFC 8949 bytes, and the expected documents built from the big-int
oracle.  GPU: the device exports must equal those documents byte for byte and import must give back working handles."""
import numpy as np
import pytest

import plonky_oracle as po
from helpers import mont_array, rand_scalars, points_to_array, array_to_point, limbs_to_ints

from plonky_b200 import serde


def oracle_msm_cbor(curve, gens, w):
    digits = -(-curve.scalar.bits // w)
    rows = []
    for P in gens:
        row, Q = [], P
        for _ in range(digits):
            row.append(po.point_to_bytes(curve, Q))
            for _ in range(w):
                Q = curve.double(Q)
        rows.append(row)
    return serde.cbor_dumps({"powers_per_generator": rows, "w": w})


def oracle_fft_cbor(field, degree):
    nb = 8 * field.limbs
    levels = [[v.to_bytes(nb, "little") for v in lvl] for lvl in po.fft_precompute(field, degree)]
    return serde.cbor_dumps({"subgroups_rev": levels})


def test_cbor_primitives_match_rfc8949():
    # RFC 8949 appendix A examples
    assert serde.cbor_dumps(0) == bytes.fromhex("00")
    assert serde.cbor_dumps(23) == bytes.fromhex("17")
    assert serde.cbor_dumps(24) == bytes.fromhex("1818")
    assert serde.cbor_dumps(1000) == bytes.fromhex("1903e8")
    assert serde.cbor_dumps(1000000) == bytes.fromhex("1a000f4240")
    assert serde.cbor_dumps(bytes.fromhex("01020304")) == bytes.fromhex("4401020304")
    assert serde.cbor_dumps("IETF") == bytes.fromhex("6449455446")
    assert serde.cbor_dumps([1, [2, 3], [4, 5]]) == bytes.fromhex("8301820203820405")
    assert serde.cbor_dumps({"a": 1, "b": [2, 3]}) == bytes.fromhex("a26161016162820203")
    assert serde.cbor_dumps(None) == bytes.fromhex("f6")
    for doc in (0, 24, 1 << 40, b"", b"\x00" * 300, "w", [], [[b"ab"], []], {"powers_per_generator": [[b"x" * 33]], "w": 11}, None):
        assert serde.cbor_loads(serde.cbor_dumps(doc)) == doc
    with pytest.raises(ValueError):
        serde.cbor_loads(bytes.fromhex("0000"))            # trailing bytes


def test_reference_document_shapes():
    """the exact leading bytes serde_cbor produces for the two structs"""
    c = po.TWEEDLEDEE
    doc = oracle_msm_cbor(c, [c.gen, c.double(c.gen)], 11)
    head = bytes([0xA2, 0x74]) + b"powers_per_generator" + bytes([0x82, 0x98, 24, 0x58, 33])
    assert doc[:len(head)] == head                          # map(2), text(20), array(2), array(24), bytes(33)
    assert doc[-3:] == bytes([0x61]) + b"w" + bytes([11])
    f = po.TWEEDLEDEE_BASE
    doc = oracle_fft_cbor(f, 8)
    head = bytes([0xA1, 0x6D]) + b"subgroups_rev" + bytes([0x84, 0x81, 0x58, 32]) + (1).to_bytes(32, "little")
    assert doc[:len(head)] == head                          # map(1), text(13), array(4), array(1), bytes(32) = ONE canonical
    lv = serde.cbor_loads(doc)["subgroups_rev"]
    assert [len(x) for x in lv] == [1, 2, 4, 8]
    assert int.from_bytes(lv[1][1], "little") == f.p - 1    # w_1 = -1


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["Tweedledee", "Tweedledum", "Bls12377"])
def test_field_and_point_bytes(name):
    import plonky_b200 as pk
    c = po.CURVES[name]
    for f in (c.base, c.scalar):
        vals = rand_scalars(f, 5, 6) + [0, 1, f.p - 1]
        enc = pk.field_to_bytes(f.fid, mont_array(f, vals))
        assert [int.from_bytes(enc[i].tobytes(), "little") for i in range(len(vals))] == vals
        assert np.array_equal(pk.field_from_bytes(f.fid, enc), mont_array(f, vals))
        bad = enc.copy()
        bad[2] = np.frombuffer(f.p.to_bytes(8 * f.limbs, "little"), dtype=np.uint8)        # == p: "Out of range"
        with pytest.raises(ValueError):
            pk.field_from_bytes(f.fid, bad)


@pytest.mark.gpu
@pytest.mark.parametrize("name,w", [("Tweedledee", 11), ("Tweedledum", 8), ("Bls12377", 5)])
def test_msm_precomputation_cbor(name, w):
    import plonky_b200 as pk
    c = po.CURVES[name]
    gens = [c.gen, c.mul(7, c.gen), None, c.neg(c.gen), po.blake_hash_usize_to_curve(c, 3)]
    xy, zero = points_to_array(c, gens)
    want = oracle_msm_cbor(c, gens, w)
    got = serde.msm_precomputation_to_cbor(c.cid, xy, w, zero)
    assert got == want
    pre = serde.msm_precomputation_from_cbor(c.cid, got, check_powers=True)
    assert len(pre) == len(gens) and pre.w == w
    assert np.array_equal(pre.generators, xy) and np.array_equal(pre.zero, zero)
    s = rand_scalars(c.scalar, 9, len(gens))
    out, oz = pk.msm_execute(pre, mont_array(c.scalar, s))
    acc = None
    for k, P in zip(s, gens):
        acc = c.add(acc, c.mul(k, P))
    assert (None if oz else array_to_point(c, out[:2], 0)) == acc
    tampered = bytearray(got)
    tampered[40] ^= 1                                       # inside the first point's x: no longer the generator's powers
    with pytest.raises(ValueError):
        serde.msm_precomputation_from_cbor(c.cid, bytes(tampered), check_powers=True)


@pytest.mark.gpu
@pytest.mark.parametrize("fname", ["TweedledeeBase", "TweedledumBase", "Bls12377Scalar", "Bls12377Base"])
@pytest.mark.parametrize("degree", [1, 2, 8, 200])
def test_fft_precomputation_cbor(fname, degree):
    import plonky_b200 as pk
    f = po.FIELDS[fname]
    want = oracle_fft_cbor(f, degree)
    got = serde.fft_precomputation_to_cbor(f.fid, degree)
    assert got == want
    pre = serde.fft_precomputation_from_cbor(f.fid, got)
    n = pre.size()
    assert n == 1 << max(0, (degree - 1).bit_length())
    x = mont_array(f, rand_scalars(f, 1, n))
    assert np.array_equal(pk.ifft_with_precomputation_power_of_2(pk.fft_with_precomputation_power_of_2(x, pre), pre), x)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["Tweedledee", "Bls12377"])
@pytest.mark.parametrize("blinding", [True, False])
def test_coeffs_vec_to_commitments(name, blinding):
    """poly_commit.rs:32-66: pedersen_hash(coeffs) + [blinding_factor] * H per polynomial, batch_to_affine over all of them --
    against the C++ restatement (msm_execute_parallel, CurveScalar * point) and big-int additions."""
    import plonky_b200 as pk
    import ref_port as rp
    c = po.CURVES[name]
    n, k = 256, 5
    g = pk.blake_hash_usize_to_curve(c.cid, 0, n + 1)
    gens, h = g[:n], g[n]                                   # pedersen_g = seeds 0..n, pedersen_h = seed n (circuit_builder.rs:1127-1128)
    pre = pk.msm_precompute_affine(c.cid, gens, 11)
    coeffs = np.stack([mont_array(c.scalar, rand_scalars(c.scalar, 40 + i, n)) for i in range(k)])
    coeffs[3] = 0                                            # the zero polynomial: the commitment is the blinding term alone
    bl = mont_array(c.scalar, rand_scalars(c.scalar, 50, k))
    out, oz = pk.coeffs_vec_to_commitments(coeffs, pre, h, bl if blinding else None)
    table = rp.MsmTable(c.cid, gens, None, 11)
    hP = array_to_point(c, h, 0)
    for i in range(k):
        xy, z = table.execute(coeffs[i], parallel=True)
        want = None if z else array_to_point(c, xy, 0)
        if blinding:
            want = c.add(want, c.mul(c.scalar.from_mont(limbs_to_ints(bl[i:i + 1])[0]), hP))
        assert (None if oz[i] else array_to_point(c, out[i], 0)) == want
    if not blinding:
        assert oz[3] and not out[3].any()


@pytest.mark.gpu
def test_new_entry_points_error_paths():
    """status codes instead of aborts on the round-2 entry points (the reference panics / returns io::Error here)"""
    import plonky_b200 as pk
    c = po.TWEEDLEDEE
    f = c.scalar
    g = pk.blake_hash_usize_to_curve(c.cid, 0, 9)
    pre = pk.msm_precompute_affine(c.cid, g[:8], 11)
    coeffs = np.stack([mont_array(f, rand_scalars(f, 1, 8))])
    # commit: scalars.len() != precomputation.len() is the reference's assert_eq! (curve_msm.rs:106)
    with pytest.raises(pk.PlonkyPanic):
        pk.coeffs_vec_to_commitments(coeffs[:, :7], pre, g[8])
    # commit with the blinding point at infinity: the blinding term vanishes
    out, oz = pk.coeffs_vec_to_commitments(coeffs, pre, np.zeros((2, 4), dtype=np.uint64), mont_array(f, [5]), blinding_point_zero=True)
    plain, pz = pk.msm_execute(pre, coeffs[0])
    assert not oz[0] and np.array_equal(out[0], plain[:2])
    # vanishing_poly: non power-of-two degree = log2_strict panic; unknown field = EINVAL
    with pytest.raises(pk.PlonkyPanic):
        pk.vanishing_points(f.fid, 6, np.zeros((9, 48, 4), np.uint64), np.zeros((6, 48, 4), np.uint64), np.zeros((6, 48, 4), np.uint64),
                            np.zeros((48, 4), np.uint64), np.zeros((48, 4), np.uint64), np.zeros((6, 4), np.uint64), *[np.zeros((1, 4), np.uint64)] * 5)
    with pytest.raises(ValueError):
        pk.lib()  # keep the library loaded
        pk._check(pk.lib().plk_vanishing_points(3, 8, *[None] * 12))
    # field bytes: wrong field id, and the all-ones pattern (>= p) is "Out of range"
    with pytest.raises(ValueError):
        pk.field_from_bytes(f.fid, np.full((1, 32), 0xFF, dtype=np.uint8))
    # degree-1 vanishing evaluation (8 points): L_1 == 1 everywhere on the trivial subgroup {1}
    deg = 1
    rows = lambda k, seed: np.stack([mont_array(f, rand_scalars(f, seed + j, 8)) for j in range(k)])
    pre8 = pk.fft_precompute(f.fid, 8)
    sub = pk.fft_subgroup(pre8)
    wires, consts, sigma, z = rows(9, 10), rows(6, 30), rows(6, 50), mont_array(f, rand_scalars(f, 70, 8))
    small = [mont_array(f, rand_scalars(f, 80 + j, 6 if j == 0 else 1)) for j in range(6)]
    got = pk.vanishing_points(f.fid, deg, wires, consts, sigma, z, sub, *small)
    C = lambda a: [f.from_mont(v) for v in limbs_to_ints(a)]
    want = po.vanishing_points(f, deg, [C(w) for w in wires], [C(w) for w in consts], [C(w) for w in sigma], C(z), C(sub), C(small[0]),
                               *[C(s)[0] for s in small[1:]])
    assert C(got) == want
