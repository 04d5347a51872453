"""The oracles (Python big-int + C++ restatement) against the reference's own known answers
(tests/golden/reference_kats.json, extracted from the reference sources) and against each other."""
import numpy as np
import pytest

import plonky_oracle as po
import ref_port as rp
from helpers import (kats, ints_to_limbs, limbs_to_ints, mont_array, canon_list, points_to_array,
                     array_to_point, rand_scalars, splitmix_hash)

K = kats()


@pytest.mark.parametrize("name", list(po.FIELDS))
def test_field_constants_match_reference(name):
    f = po.FIELDS[name]
    e = K["fields"][name]
    L = lambda key: f.from_limbs([int(v) for v in e[key]["limbs"]])
    assert L("ORDER") == f.p
    assert L("R") == f.R and L("R2") == pow(f.R, 2, f.p) and L("R3") == pow(f.R, 3, f.p)
    assert int(e["MU"]["value"]) == f.mu
    assert int(e["BITS"]["value"]) == f.bits == f.p.bit_length()
    assert int(e["TWO_ADICITY"]["value"]) == f.two_adicity
    assert (f.p - 1) % (1 << f.two_adicity) == 0 and ((f.p - 1) >> f.two_adicity) & 1
    for nm, v in (("TWO", 2), ("THREE", 3), ("FOUR", 4), ("FIVE", 5), ("NEG_ONE", f.p - 1)):
        assert L(nm) == f.to_mont(v)
    assert L("T") == f.to_mont(f.t)
    if "ORDER_X2" in e:
        assert L("ORDER_X2") == 2 * f.p
    if "MULTIPLICATIVE_SUBGROUP_GENERATOR" in e:
        assert L("MULTIPLICATIVE_SUBGROUP_GENERATOR") == f.to_mont(f.generator)
    # the generator is a quadratic non-residue, so g^T has full 2-power order
    assert pow(f.generator, (f.p - 1) // 2, f.p) == f.p - 1
    w = f.primitive_root_of_unity(f.two_adicity)
    assert pow(w, 1 << (f.two_adicity - 1), f.p) == f.p - 1


@pytest.mark.parametrize("name", list(po.CURVES))
def test_curve_constants_match_reference(name):
    c = po.CURVES[name]
    f = c.base
    e = K["curves"][name]

    def val(x):
        if isinstance(x, str):
            return {"ZERO": 0, "ONE": 1, "TWO": 2, "FIVE": 5, "NEG_ONE": f.p - 1}[x]
        return f.from_mont(f.from_limbs([int(v) for v in x["limbs"]]))
    assert val(e["A"]) == c.a and val(e["B"]) == c.b
    assert (val(e["gen_x"]), val(e["gen_y"])) == c.gen
    assert c.is_on_curve(c.gen)
    assert c.mul(c.scalar.p - 1, c.gen) == c.neg(c.gen)     # generator has the scalar-field order
    if "ZETA" in e:                                          # endomorphism test, tweedledee_curve.rs:65-74
        zeta = val(e["ZETA"])
        zs = c.scalar.from_mont(c.scalar.from_limbs([int(v) for v in e["ZETA_SCALAR"]["limbs"]]))
        P = c.mul(123456789, c.gen)
        assert c.mul(zs, P) == (zeta * P[0] % f.p, P[1])


def test_to_digits_kat():
    td = K["to_digits"]
    f = po.FIELDS[td["field"]]
    x = f.from_limbs([int(v) for v in td["x_canonical"]])
    want = [int(d) for d in td["digits"]]
    assert po.to_digits(x, td["w"], f.bits) == want
    assert rp.to_digits(2, mont_array(f, [x])[0], td["w"]) == want


def test_div2_kat():
    for case in K["div2"]["cases"]:
        got = rp.div2(np.array([int(v) for v in case["in"]], dtype=np.uint64))
        assert [int(v) for v in got] == [int(v) for v in case["out"]]


def test_reverse_bits_kat():
    rb = K["reverse_bits"]
    assert po.reverse_bits(rb["n"], rb["bits"]) == rb["out"]
    assert rp.reverse_bits(rb["n"], rb["bits"]) == rb["out"]
    assert po.reverse_index_bits(["a", "b"]) == ["a", "b"]
    assert po.reverse_index_bits(["a", "b", "c", "d"]) == rb["index_perm_4"]


@pytest.mark.parametrize("name", list(po.FIELDS))
def test_port_field_arithmetic_on_reference_test_inputs(name):
    """test_arithmetic! (field.rs:618-780): add/sub/neg/mul/square on the carry-stressing inputs,
    C++ limb arithmetic vs big ints."""
    f = po.FIELDS[name]
    vals = po.field_test_inputs(f.p, 64)
    if len(vals) > 120:
        vals = vals[::3]
    a = [x for x in vals for _ in vals]
    b = [y for _ in vals for y in vals]
    A, B = mont_array(f, a), mont_array(f, b)
    assert canon_list(f, rp.field_op(f.fid, "add", A, B)) == [(x + y) % f.p for x, y in zip(a, b)]
    assert canon_list(f, rp.field_op(f.fid, "sub", A, B)) == [(x - y) % f.p for x, y in zip(a, b)]
    assert canon_list(f, rp.field_op(f.fid, "mul", A, B)) == [(x * y) % f.p for x, y in zip(a, b)]
    V = mont_array(f, vals)
    assert canon_list(f, rp.field_op(f.fid, "square", V)) == [x * x % f.p for x in vals]
    assert canon_list(f, rp.field_op(f.fid, "neg", V)) == [(-x) % f.p for x in vals]
    assert canon_list(f, rp.field_op(f.fid, "double", V)) == [2 * x % f.p for x in vals]
    assert canon_list(f, rp.field_op(f.fid, "triple", V)) == [3 * x % f.p for x in vals]
    nz = [x for x in vals if x]
    assert canon_list(f, rp.field_op(f.fid, "inverse", mont_array(f, nz))) == [f.inv(x) for x in nz]
    assert canon_list(f, rp.batch_inverse(f.fid, mont_array(f, nz))) == [f.inv(x) for x in nz]
    with pytest.raises(ZeroDivisionError):
        rp.batch_inverse(f.fid, mont_array(f, [1, 0, 2]))
    # Montgomery round trip (bls12_377_base.rs:289-317): to_canonical(from_canonical(x)) == x
    raw = ints_to_limbs(vals, f.limbs)
    m = rp.field_op(f.fid, "from_canonical", raw)
    assert limbs_to_ints(m) == [f.to_mont(x) for x in vals]
    assert limbs_to_ints(rp.field_op(f.fid, "to_canonical", m)) == vals


@pytest.mark.parametrize("name", ["Bls12377Base", "Bls12377Scalar"])
def test_mont_mul_reference_vectors(name):
    f = po.FIELDS[name]
    e = K["mont_mul_inputs"][name]
    a = f.from_limbs([int(v) for v in e["a"]])
    b = f.from_limbs([int(v) for v in e["b"]])
    got = canon_list(f, rp.field_op(f.fid, "mul", mont_array(f, [a]), mont_array(f, [b])))
    assert got == [a * b % f.p]


@pytest.mark.parametrize("name", list(po.FIELDS))
def test_roots_of_unity(name):
    f = po.FIELDS[name]
    for k in (0, 1, 2, 5, 12, f.two_adicity):
        got = limbs_to_ints(rp.primitive_root_of_unity(f.fid, k)[None, :])[0]
        assert got == f.to_mont(f.primitive_root_of_unity(k))


def test_fft_and_ifft_reference_case():
    """fft.rs:164-185: degree 200 (padded to 256), c_i = i*1337 % 100 over Bls12377Scalar; FFT equals
    naive evaluation at all powers of the root; IFFT round-trips."""
    e = K["fft_and_ifft"]
    f = po.FIELDS[e["field"]]
    coeffs = [(i * e["mul"]) % e["mod"] for i in range(e["degree"])]
    padded = coeffs + [0] * (256 - len(coeffs))
    naive = po.dft_naive(f, padded)
    assert po.fft_pow2(f, padded) == naive
    assert po.ntt(f, padded) == naive
    assert po.fft_padded(f, coeffs) == naive
    got = canon_list(f, rp.fft(f.fid, mont_array(f, padded)))
    assert got == naive
    back = canon_list(f, rp.fft(f.fid, mont_array(f, naive), inverse=True))
    assert back == padded
    assert po.ifft_pow2(f, naive) == padded


@pytest.mark.parametrize("name", ["TweedledeeBase", "TweedledumBase"])
@pytest.mark.parametrize("logn", [0, 1, 2, 3, 7, 10])
def test_port_fft_matches_bigint(name, logn):
    f = po.FIELDS[name]
    n = 1 << logn
    x = rand_scalars(f, 1000 + logn, n)
    want = po.ntt(f, x)
    if n <= 128:
        assert want == po.dft_naive(f, x)
        assert want == po.fft_pow2(f, x)
    assert canon_list(f, rp.fft(f.fid, mont_array(f, x))) == want
    assert canon_list(f, rp.fft(f.fid, mont_array(f, want), inverse=True)) == x


def test_fft_rejects_non_power_of_two():
    f = po.TWEEDLEDEE_BASE
    with pytest.raises(AssertionError):
        rp.fft(f.fid, mont_array(f, [1, 2, 3]))
    with pytest.raises(AssertionError):
        po.log2_strict(12)


def test_divide_by_z_h_oracle():
    """polynomial.rs divide_by_z_h semantics: (a * Z_H) / Z_H == a."""
    f = po.TWEEDLEDEE_BASE
    n = 8
    a = rand_scalars(f, 9, 20)
    prod = [0] * (len(a) + n)
    for i, c in enumerate(a):            # a * (X^n - 1)
        prod[i + n] = (prod[i + n] + c) % f.p
        prod[i] = (prod[i] - c) % f.p
    q = po.divide_by_z_h(f, prod, n)
    assert q[:len(a)] == a and all(v == 0 for v in q[len(a):])


def test_test_msm_reference_case():
    """curve_msm.rs:218-241: generators G, 2G, 3G, fixed scalars, w = 5; msm_execute == naive sum."""
    e = K["test_msm"]
    c = po.CURVES[e["curve"]]
    G = c.gen
    gens = [G, c.double(G), c.add(G, c.double(G))]
    scalars = [c.scalar.from_limbs([int(v) for v in s]) for s in e["scalars_canonical"]]
    want = c.msm_naive(scalars, gens)
    table = po.msm_precompute(c, gens, e["w"])
    assert po.msm_execute(c, table, e["w"], scalars) == want
    xy, zero = points_to_array(c, gens)
    S = mont_array(c.scalar, scalars)
    t = rp.MsmTable(c.cid, xy, zero, e["w"])
    for par in (False, True):
        out, oz = t.execute(S, parallel=par)
        assert array_to_point(c, out, oz) == want
    with pytest.raises(AssertionError):
        t.execute(S[:2])


@pytest.mark.parametrize("name", list(po.CURVES))
def test_summation_reference_cases(name):
    """curve_summations.rs:164-184: {G,G}, {G,2G}, {G,G,G}, {} -- doubling and empty cases."""
    c = po.CURVES[name]
    G = c.gen
    G2 = c.double(G)
    G3 = c.add(G2, G)
    for pts, want in (([G, G], G2), ([G, G2], G3), ([G, G, G], G3), ([], None), ([G, c.neg(G)], None),
                      ([None, G], G), ([G, None, G2], G3)):
        xy, zero = points_to_array(c, pts)
        for mode in (0, 1, 2):
            out, oz = rp.affine_sum(c.cid, xy, zero, mode)
            assert array_to_point(c, out, oz) == want


@pytest.mark.parametrize("name", list(po.CURVES))
def test_port_scalar_mul_and_msm_match_bigint(name):
    c = po.CURVES[name]
    rng = po.SplitMix64(77 + c.cid)
    n = 150                                   # > 70 pair-adds in a chunk: exercises the batch-inversion tree
    base = po.rand_points(c, rng, 12)
    pts = [base[i % 12] if i % 12 else c.mul(i + 3, c.gen) for i in range(n)]
    scalars = rand_scalars(c.scalar, 5 + c.cid, n)
    scalars[0] = 0
    scalars[1] = 1
    scalars[2] = c.scalar.p - 1
    pts[5] = None
    pts[7] = c.neg(pts[6])
    scalars[7] = scalars[6]
    want = c.msm_pippenger(scalars, pts, 7)
    # pippenger oracle self-check against the naive sum on a prefix
    assert c.msm_pippenger(scalars[:16], pts[:16], 5) == c.msm_naive(scalars[:16], pts[:16])
    xy, zero = points_to_array(c, pts)
    S = mont_array(c.scalar, scalars)
    for w in (4, 8):
        t = rp.MsmTable(c.cid, xy, zero, w)
        for par in (False, True):
            out, oz = t.execute(S, parallel=par)
            assert array_to_point(c, out, oz) == want
    # single multiplication (curve_multiplication.rs) vs double-and-add
    out, oz = rp.curve_mul(c.cid, xy[3], False, S[3])
    assert array_to_point(c, out, oz) == c.mul(scalars[3], pts[3])


@pytest.mark.parametrize("name", list(po.CURVES))
def test_gen_points(name):
    c = po.CURVES[name]
    xy = rp.gen_points(c.cid, 42, 8)
    for i in range(8):
        assert array_to_point(c, xy[i], False) == c.mul(splitmix_hash(42 + i), c.gen)
