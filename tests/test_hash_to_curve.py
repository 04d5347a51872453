"""BLAKE3 generator derivation (src/hash_to_curve.rs:13-76) and the point wire format (src/serialization.rs:32-72).

CPU: the oracle's restatement of the BLAKE3 compression function and of blake_hash_usize_to_curve against golden
vectors produced with the independent `blake3` package (tools/gen_blake_golden.py -> tests/golden/blake_hash_to_curve.json).
GPU: the device kernels against the oracle and the golden vectors, bit for bit."""
import json
import os

import numpy as np
import pytest

import plonky_oracle as po
from helpers import GOLDEN, limbs_to_ints, points_to_array, array_to_point, rand_scalars, mont_array

CURVES = {c.name: c for c in (po.TWEEDLEDEE, po.TWEEDLEDUM, po.BLS12_377)}


def golden():
    with open(os.path.join(GOLDEN, "blake_hash_to_curve.json")) as f:
        return json.load(f)


def test_blake3_restatement_matches_package_vectors():
    for row in golden()["xof"]:
        data = bytes.fromhex(row["input"])
        assert po.blake3_xof_one_block(data, 64).hex() == row["out64"]
    # the published empty-input digest
    assert po.blake3_xof_one_block(b"", 32).hex() == "af1349b9f5f9a1a6a0404dea36dcc9499bcb25c9adc112b7cc9a93cae41f3262"


@pytest.mark.parametrize("name", list(CURVES))
def test_oracle_generators_match_golden(name):
    curve = CURVES[name]
    for row in golden()["generators"][name]:
        P = po.blake_hash_usize_to_curve(curve, row["seed"])
        assert P == (int(row["x"], 16), int(row["y"], 16))
        assert curve.is_on_curve(P)


@pytest.mark.parametrize("name", list(CURVES))
def test_port_generators_match_golden(name):
    """the C++ restatement (oracle/ref_port.cpp: the reference arm of bench.py derives pedersen_g with it) against the
    same golden vectors, every seed the fixture holds"""
    import ref_port as rp
    curve = CURVES[name]
    for row in golden()["generators"][name]:
        got = rp.blake_hash_usize_to_curve(curve.cid, row["seed"], 1)
        assert array_to_point(curve, got[0], 0) == (int(row["x"], 16), int(row["y"], 16))


def test_oracle_point_bytes_roundtrip():
    for curve in CURVES.values():
        pts = [po.blake_hash_usize_to_curve(curve, s) for s in range(6)] + [None, curve.gen, curve.neg(curve.gen)]
        for P in pts:
            enc = po.point_to_bytes(curve, P)
            assert len(enc) == 1 + 8 * curve.base.limbs
            assert po.point_from_bytes(curve, enc) == P
        with pytest.raises(ValueError):                       # "Out of range"
            po.point_from_bytes(curve, bytes([0]) + curve.base.p.to_bytes(8 * curve.base.limbs, "little"))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CURVES))
def test_device_generators_match_golden_and_oracle(name):
    import plonky_b200 as pk
    curve = CURVES[name]
    rows = golden()["generators"][name]
    n0 = 24
    got = pk.blake_hash_usize_to_curve(curve.cid, 0, n0)            # pedersen_g[0..24]
    for i in range(n0):
        assert rows[i]["seed"] == i
        assert array_to_point(curve, got[i], 0) == (int(rows[i]["x"], 16), int(rows[i]["y"], 16))
    for row in rows[n0:]:
        P = array_to_point(curve, pk.blake_hash_usize_to_curve(curve.cid, row["seed"], 1)[0], 0)
        assert P == (int(row["x"], 16), int(row["y"], 16))
    # more seeds against the oracle restatement (includes seeds that need several iterations)
    start, n = 1000, 200 if curve.base.limbs == 4 else 40
    got = pk.blake_hash_usize_to_curve(curve.cid, start, n)
    for i in range(n):
        assert array_to_point(curve, got[i], 0) == po.blake_hash_usize_to_curve(curve, start + i)
    # arbitrary base-field seeds (blake_hash_base_field_to_curve)
    seeds = rand_scalars(curve.base, 5, 16) + [0, curve.base.p - 1]
    got = pk.blake_hash_base_field_to_curve(curve.cid, mont_array(curve.base, seeds))
    for i, s in enumerate(seeds):
        assert array_to_point(curve, got[i], 0) == po.blake_hash_base_field_to_curve(curve, s)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CURVES))
def test_device_point_codec(name):
    import plonky_b200 as pk
    curve = CURVES[name]
    f = curve.base
    pts = [po.blake_hash_usize_to_curve(curve, s) for s in range(12)] + [None, curve.gen, curve.neg(curve.gen), None]
    xy, zero = points_to_array(curve, pts)
    enc = pk.points_to_bytes(curve.cid, xy, zero)
    assert enc.shape == (len(pts), 1 + 8 * f.limbs)
    for i, P in enumerate(pts):
        assert bytes(enc[i]) == po.point_to_bytes(curve, P)
    dec, dz = pk.points_from_bytes(curve.cid, enc)
    assert np.array_equal(dz, zero) and np.array_equal(dec, xy)
    # error cases of AffinePoint::read
    bad = enc.copy()
    bad[0, 1:] = np.frombuffer(f.p.to_bytes(8 * f.limbs, "little"), dtype=np.uint8)          # x = p: "Out of range"
    with pytest.raises(ValueError):
        pk.points_from_bytes(curve.cid, bad)
    x = 1
    while f.sqrt((x ** 3 + curve.a * x + curve.b) % f.p) is not None:
        x += 1
    bad = enc.copy()
    bad[1, 0] = 0
    bad[1, 1:] = np.frombuffer(x.to_bytes(8 * f.limbs, "little"), dtype=np.uint8)            # "Invalid x coordinate"
    with pytest.raises(ValueError):
        pk.points_from_bytes(curve.cid, bad)


@pytest.mark.gpu
@pytest.mark.parametrize("n,w", [(256, 11), (4096, 8)])
def test_msm_over_reference_generators(n, w):
    """pedersen_hash over the reference's own generator set (blake_hash_usize_to_curve(0..n), circuit_builder.rs:1127)
    against the restatement of msm_execute_parallel -- BASELINE config 1's size with the prover's window (11) and w = 8,
    edge scalars (0, 1, q - 1, 2^k) included."""
    import plonky_b200 as pk
    import ref_port as rp
    curve = po.TWEEDLEDEE
    q = curve.scalar.p
    g = pk.blake_hash_usize_to_curve(curve.cid, 0, n)
    vals = rand_scalars(curve.scalar, 31, n)
    vals[:6] = [0, 1, q - 1, 1 << 17, (1 << 254) % q, q - 2]
    scalars = mont_array(curve.scalar, vals)
    want_xy, want_zero = rp.MsmTable(curve.cid, g, None, w).execute(scalars, parallel=True)
    pre = pk.msm_precompute_affine(curve.cid, g, w)
    out, oz = pk.pedersen_hash(scalars, pre)
    assert oz == want_zero and np.array_equal(out[:2], want_xy)
