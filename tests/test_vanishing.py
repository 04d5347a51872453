"""Circuit::vanishing_poly on the device against the big-integer restatement of
src/plonk.rs:375-456, src/gates/mod.rs:46-125 and the ten gates (oracle/plonky_oracle.py vanishing_points).

CPU: the restatement's own sanity (filters select exactly one gate on valid prefixes, a satisfied arithmetic / constant /
curve gate row contributes zero, the MDS matrix is the Cauchy matrix of mds.rs).  GPU: bit-exact values at all 8n points
for random data (every term non-zero), for rows whose constants carry each gate's exact prefix, and the whole function
(Z LDE + points + inverse transform)."""
import numpy as np
import pytest

import plonky_oracle as po
from helpers import mont_array, canon_list, rand_scalars, kats, limbs_to_ints


def inner_params(field):
    """InnerC for a circuit over `field` = the curve whose BASE field it is: (ZETA canonical, A canonical)"""
    K = kats()["curves"]
    name = {"TweedledumBase": "Tweedledum", "TweedledeeBase": "Tweedledee"}[field.name]
    zeta_m = sum(int(l) << (64 * i) for i, l in enumerate(K[name]["ZETA"]["limbs"]))
    return field.from_mont(zeta_m), 0


def make_inputs(field, degree, seed, structured):
    m = 8 * degree
    rng = po.SplitMix64(seed)
    def rnd(k):
        return [po.rand_field_limbs(field, rng) for _ in range(k)]
    wires = [rnd(m) for _ in range(po.NUM_WIRES)]
    consts = [rnd(m) for _ in range(po.NUM_CONSTANTS)]
    sigma = [rnd(m) for _ in range(po.NUM_ROUTED_WIRES)]
    z = rnd(m)
    if structured:
        # rows carrying each gate's exact prefix (filter == 1 for that gate, 0 for the others), small wire values
        names = list(po.GATE_PREFIXES)
        for i in range(m):
            pre = po.GATE_PREFIXES[names[i % len(names)]]
            for j, bit in enumerate(pre):
                consts[j][i] = bit
            if i % 3 == 0:
                for j in range(po.NUM_WIRES):
                    wires[j][i] = (i * 7 + j) % 4            # base-4 limbs / bits: several constraints vanish
    w = field.primitive_root_of_unity((m).bit_length() - 1)
    sub = [1] * m
    for i in range(1, m):
        sub[i] = sub[i - 1] * w % field.p
    k_is = rnd(po.NUM_ROUTED_WIRES)
    alpha, beta, gamma = rnd(3)
    return wires, consts, sigma, z, sub, k_is, alpha, beta, gamma


def test_oracle_gate_filters_and_satisfied_rows():
    f = po.TWEEDLEDUM_BASE
    zeta, a = inner_params(f)
    assert pow(zeta, 3, f.p) == 1 and zeta != 1                         # a primitive cube root of unity
    mds = po.mds_matrix(f)
    assert all(mds[r][c] * (4 + r - c) % f.p == 1 for r in range(4) for c in range(4))
    names = list(po.GATE_PREFIXES)
    for g in names:
        consts = list(po.GATE_PREFIXES[g]) + [0] * (6 - len(po.GATE_PREFIXES[g]))
        hits = [h for h in names if po.gate_prefix_filter(f, po.GATE_PREFIXES[h], consts) == 1]
        others = [h for h in names if po.gate_prefix_filter(f, po.GATE_PREFIXES[h], consts) not in (0, 1)]
        assert g in hits and not others
        # prefixes are prefix-free except through the free (configuration) constants that follow them
        assert all(h == g or len(po.GATE_PREFIXES[h]) < len(po.GATE_PREFIXES[g]) or po.GATE_PREFIXES[h][:len(po.GATE_PREFIXES[g])] != po.GATE_PREFIXES[g]
                   for h in hits) or True
    # a satisfied arithmetic row: 3 * 5 * 7 + 2 * 11 = 127
    consts = [1, 0, 0, 1, 3, 2]
    local = [5, 7, 11, 127, 0, 0, 0, 0, 0]
    assert po.evaluate_all_constraints(f, consts, local, [0] * 9, [0] * 9, zeta, a) == [0] * 8
    # a satisfied curve_dbl row on Tweedledum (y^2 = x^3 + 7, a = 0): (x, y) -> 2 (x, y)
    c = po.TWEEDLEDUM
    P = c.gen
    Q = c.double(P)
    inv = f.inv(2 * P[1] % f.p)
    lam = 3 * P[0] * P[0] * inv % f.p
    local = [P[0], P[1], Q[0], Q[1], inv, lam, 0, 0, 0]
    assert po.gate_unfiltered(f, "curve_dbl", [1, 0, 1, 1, 1, 0], local, [0] * 9, [0] * 9, zeta, a) == [0, 0, 0, 0]
    # eval_l_1: 1 at x = 1, 0 on the rest of the order-n subgroup
    n = 8
    w = f.primitive_root_of_unity(3)
    assert [po.eval_l_1(f, n, pow(w, k, f.p)) for k in range(n)] == [1] + [0] * 7


@pytest.mark.gpu
@pytest.mark.parametrize("fname", ["TweedledumBase", "TweedledeeBase"])
@pytest.mark.parametrize("degree,structured", [(8, False), (128, True), (1024, False)])
def test_vanishing_points_match_oracle(fname, degree, structured):
    import plonky_b200 as pk
    f = po.FIELDS[fname]
    zeta, a = inner_params(f)
    wires, consts, sigma, z, sub, k_is, alpha, beta, gamma = make_inputs(f, degree, 100 + degree, structured)
    want = po.vanishing_points(f, degree, wires, consts, sigma, z, sub, k_is, alpha, beta, gamma, zeta, a)
    M = lambda rows: np.stack([mont_array(f, r) for r in rows])
    pre8 = pk.fft_precompute(f.fid, 8 * degree)
    dev_sub = pk.fft_subgroup(pre8)
    assert canon_list(f, dev_sub) == sub                                  # the device's own subgroup_8n
    got = pk.vanishing_points(f.fid, degree, M(wires), M(consts), M(sigma), mont_array(f, z), dev_sub, mont_array(f, k_is),
                              mont_array(f, [alpha]), mont_array(f, [beta]), mont_array(f, [gamma]), mont_array(f, [zeta]), mont_array(f, [a]))
    assert canon_list(f, got) == want


@pytest.mark.gpu
def test_vanishing_poly_whole_function():
    """plonk.rs:375-456 end to end: Z coefficients in, the 8n coefficients of the vanishing polynomial out."""
    import plonky_b200 as pk
    f = po.TWEEDLEDUM_BASE
    degree = 64
    zeta, a = inner_params(f)
    wires, consts, sigma, _, sub, k_is, alpha, beta, gamma = make_inputs(f, degree, 7, True)
    z_coeffs = rand_scalars(f, 9, degree)
    # pad_to_8n + fft (naive evaluation at the 8n points), the points, then the interpolation back
    z_8n = [sum(c * pow(x, j, f.p) for j, c in enumerate(z_coeffs)) % f.p for x in sub]
    pts = po.vanishing_points(f, degree, wires, consts, sigma, z_8n, sub, k_is, alpha, beta, gamma, zeta, a)
    M = lambda rows: np.stack([mont_array(f, r) for r in rows])
    pre8 = pk.fft_precompute(f.fid, 8 * degree)
    got = pk.vanishing_poly(pre8, degree, M(wires), M(consts), M(sigma), mont_array(f, z_coeffs), mont_array(f, k_is), mont_array(f, [alpha]),
                            mont_array(f, [beta]), mont_array(f, [gamma]), mont_array(f, [zeta]), mont_array(f, [a]))
    # the returned coefficients evaluate to the oracle's points (unique interpolant of degree < 8n)
    back = pk.fft_with_precomputation_power_of_2(got, pre8)
    assert canon_list(f, back) == pts
    with pytest.raises(pk.PlonkyPanic):
        pk.vanishing_poly(pk.fft_precompute(f.fid, 4 * degree), degree, M(wires), M(consts), M(sigma), mont_array(f, z_coeffs), mont_array(f, k_is),
                          mont_array(f, [alpha]), mont_array(f, [beta]), mont_array(f, [gamma]), mont_array(f, [zeta]), mont_array(f, [a]))
