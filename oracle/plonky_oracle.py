"""Big-integer oracle for plonky's MSM / NTT hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (plonky_b200/, the C-ABI
library) may import this module; it is the checker used by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg.

Everything here works on Python ints "mod p" and is therefore independent of any
limb-level arithmetic.  Each function cites the reference file:line whose
behaviour it restates (paths relative to /root/reference).  The reference is Rust
and cannot be built in this image (no cargo/rustc); parity is pinned on
  * the reference's own known-answer vectors (tests/golden/reference_kats.json,
    extracted by tools/extract_reference_kats.py), and
  * mathematical uniqueness of the outputs: an MSM result is the unique affine
    point sum(s_i * P_i); an NTT result is the unique natural-order DFT w.r.t.
    the root of unity defined at src/field/field.rs:429-435.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

MASK64 = (1 << 64) - 1


# --------------------------------------------------------------------------------------
# Fields (src/field/*.rs).  Elements cross the boundary in Montgomery form, little-endian
# u64 limbs, always fully reduced (src/field/monty.rs:38-107).
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Field:
    name: str
    fid: int            # field id used by the C ABI (include/plonky_b200.h)
    p: int              # ORDER
    limbs: int          # number of u64 limbs (4 or 6)
    bits: int           # Field::BITS
    two_adicity: int    # Field::TWO_ADICITY
    generator: int      # MULTIPLICATIVE_SUBGROUP_GENERATOR (canonical)

    @property
    def R(self) -> int:                      # src/field/monty.rs:26-27
        return (1 << (64 * self.limbs)) % self.p

    @property
    def Rbits(self) -> int:
        return 64 * self.limbs

    @property
    def mu(self) -> int:                     # -p^-1 mod 2^64, src/field/monty.rs:33
        return (-pow(self.p, -1, 1 << 64)) % (1 << 64)

    @property
    def t(self) -> int:                      # odd part of p-1 (Field::T)
        return (self.p - 1) >> self.two_adicity

    # canonical <-> Montgomery (the reference names these from_monty / to_monty the
    # "wrong" way round, src/field/monty.rs:169-177)
    def to_mont(self, x: int) -> int:
        return (x % self.p) * self.R % self.p

    def from_mont(self, xm: int) -> int:
        return xm * pow(self.R, -1, self.p) % self.p

    def to_limbs(self, x: int) -> List[int]:
        return [(x >> (64 * i)) & MASK64 for i in range(self.limbs)]

    def from_limbs(self, limbs: Sequence[int]) -> int:
        return sum(int(l) << (64 * i) for i, l in enumerate(limbs))

    def mont_limbs(self, x: int) -> List[int]:
        """canonical int -> Montgomery limbs (what the ABI carries)."""
        return self.to_limbs(self.to_mont(x))

    def inv(self, x: int) -> int:
        if x % self.p == 0:
            raise ZeroDivisionError("No inverse")   # src/field/field.rs:267
        return pow(x, -1, self.p)

    def primitive_root_of_unity(self, n_power: int) -> int:
        """src/field/field.rs:429-435: (g^T)^(2^(TWO_ADICITY - n_power))."""
        assert n_power <= self.two_adicity
        base_root = pow(self.generator, self.t, self.p)
        return pow(base_root, 1 << (self.two_adicity - n_power), self.p)

    def is_qr(self, x: int) -> bool:
        return x % self.p == 0 or pow(x, (self.p - 1) // 2, self.p) == 1

    def sqrt(self, x: int) -> Optional[int]:
        """Tonelli-Shanks, same structure as src/field/field.rs:440-473 (so the SAME root
        of the two is returned for a given input)."""
        p = self.p
        x %= p
        if x == 0:
            return 0
        if not self.is_qr(x):
            return None
        z = pow(self.generator, self.t, p)
        w = pow(x, (self.t - 1) // 2, p)
        xx = w * x % p
        b = xx * w % p
        v = self.two_adicity
        while b != 1:
            k = 0
            b2k = b
            while b2k != 1:
                b2k = b2k * b2k % p
                k += 1
            j = v - k - 1
            w = z
            for _ in range(j):
                w = w * w % p
            z = w * w % p
            b = b * z % p
            xx = xx * w % p
            v = k
        return xx


TWEEDLEDEE_BASE = Field(
    "TweedledeeBase", 0,
    0x40000000000000000000000000000000038AA127696286C9842CAFD400000001, 4, 255, 34, 5)
TWEEDLEDUM_BASE = Field(
    "TweedledumBase", 1,
    0x40000000000000000000000000000000038AA1276C3F59B9A14064E200000001, 4, 255, 33, 5)
BLS12_377_SCALAR = Field(
    "Bls12377Scalar", 2,
    0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001, 4, 253, 47, 11)
BLS12_377_BASE = Field(
    "Bls12377Base", 3,
    0x1AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001,
    6, 377, 46, 5)

FIELDS = {f.name: f for f in (TWEEDLEDEE_BASE, TWEEDLEDUM_BASE, BLS12_377_SCALAR, BLS12_377_BASE)}
FIELDS_BY_ID = {f.fid: f for f in FIELDS.values()}


# --------------------------------------------------------------------------------------
# Curves (src/curve/*_curve.rs): y^2 = x^3 + a x + b over base field, scalars in scalar field
# --------------------------------------------------------------------------------------
Affine = Optional[Tuple[int, int]]      # None == the zero point (AffinePoint::ZERO)


@dataclass(frozen=True)
class Curve:
    name: str
    cid: int
    base: Field
    scalar: Field
    a: int
    b: int
    gen: Tuple[int, int]

    def is_on_curve(self, P: Affine) -> bool:           # src/curve/curve.rs:92-95
        if P is None:
            return True
        x, y = P
        p = self.base.p
        return (y * y - (x * x * x + self.a * x + self.b)) % p == 0

    def neg(self, P: Affine) -> Affine:
        if P is None:
            return None
        return (P[0], (-P[1]) % self.base.p)

    def add(self, P: Affine, Q: Affine) -> Affine:
        """Affine group law with the same case split as src/curve/curve_adds.rs:92-128 /
        src/curve/curve_summations.rs:107-141 (zero, P == -Q, P == Q, general)."""
        if P is None:
            return Q
        if Q is None:
            return P
        p = self.base.p
        x1, y1 = P
        x2, y2 = Q
        if x1 == x2:
            if (y1 + y2) % p == 0:
                return None
            lam = (3 * x1 * x1 + self.a) * pow(2 * y1, -1, p) % p
        else:
            lam = (y1 - y2) * pow(x1 - x2, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        y3 = (lam * (x1 - x3) - y1) % p
        return (x3, y3)

    def double(self, P: Affine) -> Affine:
        return self.add(P, P)

    def mul(self, k: int, P: Affine) -> Affine:
        """Plain double-and-add; the value equals CurveScalar * ProjectivePoint
        (src/curve/curve_multiplication.rs:63-70) normalised by to_affine."""
        k %= self.scalar.p
        acc = None
        addend = P
        while k:
            if k & 1:
                acc = self.add(acc, addend)
            addend = self.double(addend)
            k >>= 1
        return acc

    def msm_naive(self, scalars: Sequence[int], points: Sequence[Affine]) -> Affine:
        """sum_i scalars[i] * points[i]; scalars canonical ints."""
        assert len(scalars) == len(points)          # src/curve/curve_msm.rs:67,106
        acc = None
        for s, P in zip(scalars, points):
            acc = self.add(acc, self.mul(s, P))
        return acc

    def msm_pippenger(self, scalars: Sequence[int], points: Sequence[Affine], c: int = 8) -> Affine:
        """An independent bucket MSM for medium sizes (still exact big-int arithmetic)."""
        assert len(scalars) == len(points)
        q = self.scalar.p
        nwin = (self.scalar.bits + c - 1) // c
        total = None
        for w in reversed(range(nwin)):
            for _ in range(c):
                total = self.double(total)
            buckets: List[Affine] = [None] * (1 << c)
            for s, P in zip(scalars, points):
                d = ((s % q) >> (w * c)) & ((1 << c) - 1)
                if d:
                    buckets[d] = self.add(buckets[d], P)
            run = None
            acc = None
            for d in range((1 << c) - 1, 0, -1):
                run = self.add(run, buckets[d])
                acc = self.add(acc, run)
            total = self.add(total, acc)
        return total


TWEEDLEDEE = Curve("Tweedledee", 0, TWEEDLEDEE_BASE, TWEEDLEDUM_BASE, 0, 5,
                   (TWEEDLEDEE_BASE.p - 1, 2))                       # tweedledee_curve.rs:7-19
TWEEDLEDUM = Curve("Tweedledum", 1, TWEEDLEDUM_BASE, TWEEDLEDEE_BASE, 0, 7,
                   (1, 16025420084666841664781398864595112061374725945655977686277658203168167360747))
BLS12_377 = Curve(
    "Bls12377", 2, BLS12_377_BASE, BLS12_377_SCALAR, 0, 1,
    (81937999373150964239938255573465948239988671502647976594219695644855304257327692006745978603320413799295628339695,
     241266749859715473739788878240585681733927191168601896383759122102112907357779751001206799952863815012735208165030))
CURVES = {c.name: c for c in (TWEEDLEDEE, TWEEDLEDUM, BLS12_377)}
CURVES_BY_ID = {c.cid: c for c in CURVES.values()}


# --------------------------------------------------------------------------------------
# MSM pieces (src/curve/curve_msm.rs)
# --------------------------------------------------------------------------------------
def to_digits(x_canonical: int, w: int, scalar_bits: int) -> List[int]:
    """src/curve/curve_msm.rs:159-180: ceil(BITS/w) little-endian w-bit digits of the
    canonical scalar, the top digit being short."""
    num_digits = (scalar_bits + w - 1) // w
    x = x_canonical & ((1 << scalar_bits) - 1)
    return [(x >> (i * w)) & ((1 << w) - 1) for i in range(num_digits)]


def msm_precompute(curve: Curve, generators: Sequence[Affine], w: int) -> List[List[Affine]]:
    """src/curve/curve_msm.rs:27-52: per generator the powers [(2^w)^j] G, j < digits."""
    digits = (curve.scalar.bits + w - 1) // w
    table = []
    for g in generators:
        row = [g]
        for _ in range(1, digits):
            nxt = row[-1]
            for _ in range(w):
                nxt = curve.double(nxt)
            row.append(nxt)
        table.append(row)
    return table


def msm_execute(curve: Curve, table: List[List[Affine]], w: int, scalars: Sequence[int]) -> Affine:
    """src/curve/curve_msm.rs:63-100 (Yao's method with one shared bucket set):
    u += sum of table entries whose digit == d; y += u, for d = 2^w-1 .. 1."""
    if len(table) != len(scalars):
        raise AssertionError("precomputation / scalars length mismatch")   # :67
    base = 1 << w
    occ: List[List[Tuple[int, int]]] = [[] for _ in range(base)]
    for i, s in enumerate(scalars):
        for j, d in enumerate(to_digits(s % curve.scalar.p, w, curve.scalar.bits)):
            occ[d].append((i, j))
    y = None
    u = None
    for d in range(base - 1, 0, -1):
        for (i, j) in occ[d]:
            u = curve.add(u, table[i][j])
        y = curve.add(y, u)
    return y


# --------------------------------------------------------------------------------------
# NTT pieces (src/fft.rs)
# --------------------------------------------------------------------------------------
def log2_strict(n: int) -> int:
    """src/util.rs:16-19: panics unless n is a power of two."""
    if n <= 0 or n & (n - 1):
        raise AssertionError("Not a power of two")
    return n.bit_length() - 1


def log2_ceil(n: int) -> int:
    """src/util.rs:11-13."""
    return 0 if n <= 1 else (n - 1).bit_length()


def reverse_bits(n: int, num_bits: int) -> int:
    """src/fft.rs:18-26."""
    r = 0
    for i in range(num_bits):
        r |= ((n >> i) & 1) << (num_bits - 1 - i)
    return r


def reverse_index_bits(arr: list) -> list:
    """src/fft.rs:8-16."""
    n = len(arr)
    k = log2_strict(n)
    return [arr[reverse_bits(i, k)] for i in range(n)]


def fft_precompute(field: Field, degree: int) -> List[List[int]]:
    """src/fft.rs:47-59: for each layer i <= log2_ceil(degree) the order-2^i subgroup,
    index-bit-reversed."""
    out = []
    for i in range(log2_ceil(degree) + 1):
        g = field.primitive_root_of_unity(i)
        sub = [1] * (1 << i)
        for k in range(1, 1 << i):
            sub[k] = sub[k - 1] * g % field.p
        out.append(reverse_index_bits(sub))
    return out


def fft_pow2(field: Field, coeffs: Sequence[int], pre: Optional[List[List[int]]] = None) -> List[int]:
    """src/fft.rs:103-156, layer by layer exactly as the reference: bit-reverse, log n layers of
    (even + tw*odd, even - tw*odd) with tw = subgroups_rev[i][2k], bit-reverse."""
    n = len(coeffs)
    k = log2_strict(n)
    p = field.p
    if pre is None:
        pre = fft_precompute(field, n)
    ev = reverse_index_bits([c % p for c in coeffs])
    half = n >> 1
    for i in range(1, k + 1):
        ppp = 1 << i
        pairs = 1 << (i - 1)
        new = [0] * n
        for pair in range(half):
            poly = pair // pairs
            within = pair % pairs
            c0 = poly * ppp + within
            c1 = c0 + pairs
            tw = pre[i][within * 2]
            prod = tw * ev[c1] % p
            new[2 * pair] = (ev[c0] + prod) % p
            new[2 * pair + 1] = (ev[c0] - prod) % p
        ev = new
    return reverse_index_bits(ev)


def dft_naive(field: Field, coeffs: Sequence[int]) -> List[int]:
    """out[k] = sum_j c_j w^(jk), Horner per point (src/fft.rs:197-232, evaluate_naive)."""
    n = len(coeffs)
    k = log2_strict(n)
    p = field.p
    w = field.primitive_root_of_unity(k)
    out = []
    x = 1
    for _ in range(n):
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * x + c) % p
        out.append(acc)
        x = x * w % p
    return out


def ntt(field: Field, coeffs: Sequence[int], root: Optional[int] = None) -> List[int]:
    """Fast exact natural-order DFT (recursive radix-2), used for big golden vectors; equal
    by uniqueness to fft_pow2 / dft_naive."""
    n = len(coeffs)
    k = log2_strict(n)
    p = field.p
    if root is None:
        root = field.primitive_root_of_unity(k)
    a = [c % p for c in coeffs]
    if n == 1:
        return a
    a = reverse_index_bits(a)
    m = 1
    while m < n:
        wm = pow(root, n // (2 * m), p)
        for s in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                u = a[s + j]
                v = a[s + j + m] * w % p
                a[s + j] = (u + v) % p
                a[s + j + m] = (u - v) % p
                w = w * wm % p
        m *= 2
    return a


def fft_padded(field: Field, coeffs: Sequence[int]) -> List[int]:
    """src/fft.rs:61-80 fft_with_precomputation: zero-pad to the next power of two."""
    n = 1 << log2_ceil(len(coeffs))
    return ntt(field, list(coeffs) + [0] * (n - len(coeffs)))


def ifft_pow2(field: Field, points: Sequence[int]) -> List[int]:
    """src/fft.rs:82-101: forward transform, then swap i <-> n-i and scale by n^-1."""
    n = len(points)
    p = field.p
    n_inv = field.inv(n % p)
    r = ntt(field, points)
    out = [0] * n
    for i in range(n):
        out[i] = r[(n - i) % n] * n_inv % p
    return out


def coset_lde(field: Field, coeffs: Sequence[int], out_len: int, shift: Optional[int] = None) -> List[int]:
    """Evaluate the polynomial on shift*H, |H| = out_len: c_i * g^i then zero-pad then FFT
    (src/polynomial.rs:336-347 with g = MULTIPLICATIVE_SUBGROUP_GENERATOR; shift=1 gives
    src/plonk_util.rs:179-190 polynomials_to_values_padded)."""
    p = field.p
    g = field.generator if shift is None else shift
    assert len(coeffs) <= out_len
    gp = 1
    scaled = []
    for c in coeffs:
        scaled.append(c * gp % p)
        gp = gp * g % p
    return ntt(field, scaled + [0] * (out_len - len(scaled)))


def divide_by_z_h(field: Field, a: Sequence[int], n: int) -> List[int]:
    """src/polynomial.rs:330-380 (assumes Z_H | a).  Returns the coefficient vector of the
    same length as the padded transform, like the reference."""
    p = field.p
    a = list(a)
    while a and a[-1] % p == 0:
        a.pop()
    if not a:
        return []
    g = field.generator
    d = len(a) - 1
    size = 1 << log2_ceil(d + 1)
    ev = coset_lde(field, a, size)
    root = field.primitive_root_of_unity(log2_ceil(d + 1))
    den_g = pow(g, n, p)
    root_n = pow(root, n, p)
    rp = 1
    out = []
    for i in range(size):
        if i:
            rp = rp * root_n % p
        out.append(ev[i] * field.inv((den_g * rp - 1) % p) % p)
    coeffs = ifft_pow2(field, out)
    g_inv = field.inv(g)
    gp = 1
    for i in range(size):
        coeffs[i] = coeffs[i] * gp % p
        gp = gp * g_inv % p
    return coeffs


# --------------------------------------------------------------------------------------
# Deterministic synthetic inputs; shared by tests and bench.
# --------------------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------------
# src/halo.rs:63-124 -- one round of the Halo inner-product argument on canonical ints / affine points
# (without the blinding and U' terms, which are single scalar multiplications added by the caller)
# ---------------------------------------------------------------------------------------------
def halo_round_lr(curve: Curve, a: Sequence[int], b: Sequence[int], g: Sequence[Affine]):
    """(<a_lo, G_hi>, <a_hi, G_lo>, <a_lo, b_hi>, <a_hi, b_lo>)  -- halo.rs:87-93"""
    n = len(a)
    log2_strict(n)
    assert len(b) == n and len(g) == n                      # debug_assert_eq!, halo.rs:67-69
    m = n // 2
    q = curve.scalar.p
    l = curve.msm_naive(a[:m], g[m:])
    r = curve.msm_naive(a[m:], g[:m])
    ip_l = sum(x * y for x, y in zip(a[:m], b[m:])) % q     # Field::inner_product, field.rs:214-221
    ip_r = sum(x * y for x, y in zip(a[m:], b[:m])) % q
    return l, r, ip_l, ip_r


def halo_fold(curve: Curve, a: Sequence[int], b: Sequence[int], g: Sequence[Affine], u: int, u_inv: int):
    """halo.rs:117-123: (u^-1 a_hi + u a_lo, u^-1 b_lo + u b_hi, [u^-1] G_lo + [u] G_hi)"""
    n = len(a)
    m = n // 2
    q = curve.scalar.p
    na = [(u_inv * a[m + i] + u * a[i]) % q for i in range(m)]
    nb = [(u_inv * b[i] + u * b[m + i]) % q for i in range(m)]
    ng = [curve.add(curve.mul(u_inv, g[i]), curve.mul(u, g[m + i])) for i in range(m)]
    return na, nb, ng


# ---------------------------------------------------------------------------------------------
# src/hash_to_curve.rs:13-76 -- BLAKE3-based generator derivation (pedersen_g / pedersen_h / U of
# src/circuit_builder.rs:1127-1129).  The reference depends on the blake3 crate (Cargo.toml: blake3 = "0.3.3"),
# absent from /root/reference: this restates the published BLAKE3 compression function for the only case the
# path needs -- inputs of at most one 64-byte block, at most 64 bytes of extended output -- and is pinned
# against the independent `blake3` Python package (tests/test_hash_to_curve.py, tools/gen_blake_golden.py).
# ---------------------------------------------------------------------------------------------
BLAKE3_IV = (0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19)
BLAKE3_MSG_PERMUTATION = (2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
BLAKE3_CHUNK_START, BLAKE3_CHUNK_END, BLAKE3_ROOT = 1, 2, 8


def _rotr32(x: int, n: int) -> int:
    return ((x >> n) | (x << (32 - n))) & 0xFFFFFFFF


def blake3_compress(cv: Sequence[int], block_words: Sequence[int], counter: int, block_len: int, flags: int) -> List[int]:
    """The BLAKE3 compression function: 7 rounds of the quarter-round G over a 4x4 state, full 16-word output."""
    v = list(cv) + list(BLAKE3_IV[:4]) + [counter & 0xFFFFFFFF, (counter >> 32) & 0xFFFFFFFF, block_len, flags]
    m = list(block_words)

    def g(a, b, c, d, mx, my):
        v[a] = (v[a] + v[b] + mx) & 0xFFFFFFFF
        v[d] = _rotr32(v[d] ^ v[a], 16)
        v[c] = (v[c] + v[d]) & 0xFFFFFFFF
        v[b] = _rotr32(v[b] ^ v[c], 12)
        v[a] = (v[a] + v[b] + my) & 0xFFFFFFFF
        v[d] = _rotr32(v[d] ^ v[a], 8)
        v[c] = (v[c] + v[d]) & 0xFFFFFFFF
        v[b] = _rotr32(v[b] ^ v[c], 7)

    for rnd in range(7):
        g(0, 4, 8, 12, m[0], m[1])
        g(1, 5, 9, 13, m[2], m[3])
        g(2, 6, 10, 14, m[4], m[5])
        g(3, 7, 11, 15, m[6], m[7])
        g(0, 5, 10, 15, m[8], m[9])
        g(1, 6, 11, 12, m[10], m[11])
        g(2, 7, 8, 13, m[12], m[13])
        g(3, 4, 9, 14, m[14], m[15])
        if rnd < 6:
            m = [m[BLAKE3_MSG_PERMUTATION[i]] for i in range(16)]
    for i in range(8):
        v[i] ^= v[i + 8]
        v[i + 8] ^= cv[i]
    return v


def blake3_xof_one_block(data: bytes, out_len: int) -> bytes:
    """blake3::Hasher::new().update(data).finalize_xof().fill(out) for len(data) <= 64 and out_len <= 64."""
    assert len(data) <= 64 and out_len <= 64
    block = data + bytes(64 - len(data))
    words = [int.from_bytes(block[4 * i:4 * i + 4], "little") for i in range(16)]
    out = blake3_compress(BLAKE3_IV, words, 0, len(data), BLAKE3_CHUNK_START | BLAKE3_CHUNK_END | BLAKE3_ROOT)
    return b"".join(w.to_bytes(4, "little") for w in out)[:out_len]


def blake_field(field: Field, it: int, seed: int) -> Tuple[int, bool]:
    """blake_field(iter, seed) (hash_to_curve.rs:13-51): (x, y_neg)."""
    nbytes = 8 * field.limbs
    j = 0
    while True:
        data = seed.to_bytes(nbytes, "little") + bytes([it, j])
        h = bytearray(blake3_xof_one_block(data, nbytes + 1))
        h[nbytes - 1] >>= 8 * nbytes - field.bits
        x = int.from_bytes(h[:nbytes], "little")
        if x < field.p:                                    # from_canonical_u8_vec: Err("Out of range") otherwise
            return x, (h[nbytes] & 1) == 1
        j += 1


def blake_hash_base_field_to_curve(curve: Curve, seed: int) -> Affine:
    """hash_to_curve.rs:57-76 (MapToGroup)."""
    f = curve.base
    i = 0
    while True:
        x, y_neg = blake_field(f, i, seed)
        y = f.sqrt((x * x * x + curve.a * x + curve.b) % f.p)      # Field::square_root, field.rs:440-473
        if y is not None:
            if y_neg:
                y = (-y) % f.p
            return (x, y)
        i += 1


def blake_hash_usize_to_curve(curve: Curve, seed: int) -> Affine:
    return blake_hash_base_field_to_curve(curve, seed)     # from_canonical_usize(seed), hash_to_curve.rs:53-55


# src/serialization.rs:32-72 -- AffinePoint ToBytes / FromBytes
def point_to_bytes(curve: Curve, P: Affine) -> bytes:
    nbytes = 8 * curve.base.limbs
    if P is None:
        return bytes([1]) + bytes(nbytes)                  # zero = 1, y = ZERO is even, x = ZERO
    return bytes([(P[1] & 1) << 1]) + P[0].to_bytes(nbytes, "little")


def point_from_bytes(curve: Curve, data: bytes) -> Affine:
    f = curve.base
    mask = data[0]
    if mask & 1:
        return None
    x = int.from_bytes(data[1:1 + 8 * f.limbs], "little")
    if x >= f.p:
        raise ValueError("Out of range")
    y = f.sqrt((x * x * x + curve.a * x + curve.b) % f.p)
    if y is None:
        raise ValueError("Invalid x coordinate")
    return (x, y) if (y & 1) == ((mask & 2) >> 1) else (x, (-y) % f.p)


def poly_mul(field: Field, a: Sequence[int], b: Sequence[int]) -> List[int]:
    """Polynomial::mul (src/polynomial.rs:209-227) on canonical ints: schoolbook product, returned with the length the
    reference's FFT path produces (2^log2_ceil(deg a + deg b + 1), or [0] for a zero operand)."""
    def trim(x):
        x = [v % field.p for v in x]
        while x and x[-1] == 0:
            x.pop()
        return x
    a, b = trim(a), trim(b)
    if not a or not b:
        return [0]
    size = len(a) + len(b) - 1
    out = [0] * (1 << log2_ceil(size))
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] = (out[i + j] + x * y) % field.p
    return out


def permutation_polynomial(field: Field, subgroup: Sequence[int], wire_values, sigma_values, k_is: Sequence[int], beta: int, gamma: int,
                           num_routed: int = 6, sigma_stride: int = 8) -> List[int]:
    """src/plonk_util.rs:233-262 on canonical ints: wire_values[i][j], sigma_values[j][sigma_stride * i]."""
    p = field.p
    z = [1]
    for i in range(1, len(subgroup)):
        x = subgroup[i - 1]
        num = den = 1
        for j in range(num_routed):
            w = wire_values[i - 1][j]
            num = num * (w + beta * k_is[j] * x + gamma) % p
            den = den * (w + beta * sigma_values[j][sigma_stride * (i - 1)] + gamma) % p
        if den == 0:
            raise ZeroDivisionError("No inverse")
        z.append(z[-1] * num * pow(den, -1, p) % p)
    return z


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & MASK64

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)


def rand_field_limbs(field: Field, rng: SplitMix64) -> int:
    """Rejection sampler with the shape of src/bigint/bigint_arithmetic.rs:98-117: N u64
    draws, top limb shifted right by the modulus' leading zeros, retry until < p.  The
    returned integer is used directly as the (Montgomery) limb pattern."""
    strip = 64 - (field.p >> (64 * (field.limbs - 1))).bit_length()
    while True:
        v = 0
        for i in range(field.limbs):
            l = rng.next_u64()
            if i == field.limbs - 1:
                l >>= strip
            v |= l << (64 * i)
        if v < field.p:
            return v


def rand_points(curve: Curve, rng: SplitMix64, n: int) -> List[Affine]:
    """x from the PRNG, keep if x^3+ax+b is a QR, y from Tonelli-Shanks, sign from a PRNG bit."""
    f = curve.base
    out: List[Affine] = []
    while len(out) < n:
        x = rand_field_limbs(f, rng)
        rhs = (x * x * x + curve.a * x + curve.b) % f.p
        y = f.sqrt(rhs)
        if y is None or y == 0:
            continue
        if rng.next_u64() & 1:
            y = f.p - y
        out.append((x, y))
    return out


def field_test_inputs(modulus: int, word_bits: int = 64) -> List[int]:
    """The carry-stressing value set of src/field/field.rs:498-547."""
    modwords = -(-modulus.bit_length() // word_bits)
    smalls = list(range(10))
    word_max = (1 << word_bits) - 1
    bigs = [word_max - x for x in smalls]
    one_words = smalls + bigs
    multiple = [x << (word_bits * i) for i in range(1, modwords) for x in one_words]
    basic = one_words + multiple
    maxval = (1 << (modwords * word_bits)) - 1
    diff_max = [maxval - x for x in basic if maxval - x < modulus]
    diff_mod = [modulus - x for x in basic if x < modulus and x != 0]
    basics = [x for x in basic if x < modulus]
    return basics + diff_max + diff_mod


# ---------------------------------------------------------------------------------------------
# src/plonk.rs:375-456 vanishing_poly (the pointwise part) with src/gates/mod.rs:46-125 evaluate_all_constraints and the
# ten gates' evaluate_unfiltered (src/gates/*.rs), src/plonk_util.rs:7-33 (eval_zero_poly, eval_l_1, reduce_with_powers),
# src/mds.rs:63-77 (Cauchy MDS matrix).  Canonical integers mod p throughout.
# ---------------------------------------------------------------------------------------------
NUM_WIRES, NUM_ROUTED_WIRES, NUM_CONSTANTS, GRID_WIDTH, RESCUE_SPONGE_WIDTH = 9, 6, 6, 65, 4
NUM_ADVICE_WIRES = NUM_WIRES - NUM_ROUTED_WIRES
GATE_PREFIXES = {                       # src/gates/*.rs `const PREFIX`
    "curve_add": (1, 0, 1, 0, 1), "curve_dbl": (1, 0, 1, 1, 1), "curve_endo": (1, 1), "base_4_sum": (1, 0, 0, 0),
    "public_input": (1, 0, 1, 0, 0, 1), "buffer": (1, 0, 1, 0, 0, 0), "constant": (1, 0, 1, 1, 0), "arithmetic": (1, 0, 0, 1),
    "rescue_a": (0, 0), "rescue_b": (0, 1),
}
GATE_ORDER = ("curve_add", "curve_dbl", "curve_endo", "base_4_sum", "public_input", "buffer", "constant", "arithmetic",
              "rescue_a", "rescue_b")               # the order of evaluate_all_constraints, gates/mod.rs:52-113


def mds_matrix(field: Field, n: int = RESCUE_SPONGE_WIDTH) -> List[List[int]]:
    """mds.rs:63-77: Cauchy matrix 1 / (x_r - y_c), x_r = n + r, y_c = c."""
    return [[field.inv((n + r - c) % field.p) for c in range(n)] for r in range(n)]


def eval_l_1(field: Field, n: int, x: int) -> int:
    """plonk_util.rs:14-24"""
    p = field.p
    if x == 1:
        return 1
    return (pow(x, n, p) - 1) * field.inv(n * (x - 1) % p) % p


def reduce_with_powers(field: Field, terms: Sequence[int], alpha: int) -> int:
    """plonk_util.rs:27-33"""
    s = 0
    for t in reversed(terms):
        s = (s * alpha + t) % field.p
    return s


def gate_prefix_filter(field: Field, prefix, consts) -> int:
    """gates/mod.rs:281-293"""
    prod = 1
    for i, bit in enumerate(prefix):
        prod = prod * (consts[i] if bit else (1 - consts[i])) % field.p
    return prod


def gate_unfiltered(field: Field, name: str, consts, local, right, below, inner_zeta: int, inner_a: int) -> List[int]:
    p = field.p
    pl = len(GATE_PREFIXES[name])
    if name == "curve_add":               # gates/curve_add.rs:38-81
        x1, y1, so, sn, x2, y2, bit, inv, lam = local[:9]
        x4, y4 = right[0], right[1]
        x3 = (lam * lam - x1 - x2) % p
        y3 = (lam * (x1 - x4) - y1) % p
        nb = (1 - bit) % p
        return [((y1 - y2) * inv - lam) % p, (bit * x3 + nb * x1 - x4) % p, (bit * y3 + nb * y1 - y4) % p,
                (sn - (2 * so + bit)) % p, bit * nb % p, (inv * (x1 - x2) - 1) % p]
    if name == "curve_dbl":               # gates/curve_dbl.rs:31-60
        xo, yo, xn, yn, inv, lam = local[:6]
        return [((3 * xo * xo + inner_a) * inv - lam) % p, (lam * lam - 2 * xo - xn) % p, (lam * (xo - xn) - yo - yn) % p,
                (2 * yo * inv - 1) % p]
    if name == "curve_endo":              # gates/curve_endo.rs:37-85
        x1, y1, su_old, ss_old, x_in, y_in, b0, b1, inv = local[:9]
        x3, y3 = right[0], right[1]
        su_new, ss_new = below[2], below[3]
        mult = ((inner_zeta - 1) * b1 + 1) % p
        x2 = mult * x_in % p
        y2 = (2 * b0 - 1) * y_in % p
        lam = (y1 - y2) * inv % p
        signed_limb = (2 * b0 - 1) * mult % p
        return [(lam * lam - x1 - x2 - x3) % p, (lam * (x1 - x3) - y1 - y3) % p, (su_new - (4 * su_old + 2 * b1 + b0)) % p,
                (ss_new - (2 * ss_old + signed_limb)) % p, b0 * (b0 - 1) % p, b1 * (b1 - 1) % p, (inv * (x1 - x2) - 1) % p]
    if name == "base_4_sum":              # gates/base_4_sum.rs:36-62
        acc = local[0]
        limbs = local[2:9]
        for l in limbs:
            acc = (4 * acc + l) % p
        out = [(acc - local[1]) % p]
        for l in limbs:
            out.append(l * (l - 1) * (l - 2) * (l - 3) % p)
        return out
    if name == "public_input":            # gates/public_input.rs:32-43
        return [(local[NUM_ROUTED_WIRES + i] - right[i]) % p for i in range(NUM_ADVICE_WIRES)]
    if name == "buffer":
        return []
    if name == "constant":                # gates/constant.rs:28-37
        return [(consts[pl] - local[0]) % p]
    if name == "arithmetic":              # gates/arithmetic.rs:35-49
        return [(consts[pl] * local[0] * local[1] + consts[pl + 1] * local[2] - local[3]) % p]
    mds = mds_matrix(field)
    W = RESCUE_SPONGE_WIDTH
    if name == "rescue_a":                # gates/rescue_a.rs:37-64
        ins, roots, outs = local[:W], local[W:2 * W], right[:W]
        out = []
        for i in range(W):
            out.append((pow(roots[i], 5, p) - ins[i]) % p)
            out.append((consts[pl + i] + sum(mds[i][j] * roots[j] for j in range(W)) - outs[i]) % p)
        return out
    if name == "rescue_b":                # gates/rescue_b.rs:32-56
        exps = [pow(v, 5, p) for v in local[:W]]
        return [(consts[pl + i] + sum(mds[i][j] * exps[j] for j in range(W)) - right[i]) % p for i in range(W)]
    raise KeyError(name)


def evaluate_all_constraints(field: Field, consts, local, right, below, inner_zeta: int, inner_a: int) -> List[int]:
    """gates/mod.rs:46-125: index-wise sum over the gates of filter * constraint."""
    unified: List[int] = []
    for name in GATE_ORDER:
        f = gate_prefix_filter(field, GATE_PREFIXES[name], consts)
        cs = gate_unfiltered(field, name, consts, local, right, below, inner_zeta, inner_a)
        while len(unified) < len(cs):
            unified.append(0)
        for i, c in enumerate(cs):
            unified[i] = (unified[i] + f * c) % field.p
    return unified


def vanishing_points(field: Field, degree: int, wires_8n, constants_8n, sigma_8n, z_8n, subgroup_8n, k_is, alpha: int, beta: int,
                     gamma: int, inner_zeta: int, inner_a: int, indices=None) -> List[int]:
    """plonk.rs:393-452: the vanishing polynomial evaluated at the 8n points (before Polynomial::from_evaluations);
    `indices` restricts the evaluation to a sample of the points (the inputs are indexable by point)."""
    p = field.p
    m = 8 * degree
    out = []
    for i in (range(m) if indices is None else indices):
        x = subgroup_8n[i]
        consts = [constants_8n[j][i] for j in range(NUM_CONSTANTS)]
        ir, ib = (i + 8) % m, (i + 8 * GRID_WIDTH) % m
        local = [wires_8n[j][i] for j in range(NUM_WIRES)]
        right = [wires_8n[j][ir] for j in range(NUM_WIRES)]
        below = [wires_8n[j][ib] for j in range(NUM_WIRES)]
        terms = evaluate_all_constraints(field, consts, local, right, below, inner_zeta, inner_a)
        z_x, z_gz = z_8n[i], z_8n[ir]
        z1 = eval_l_1(field, degree, x) * (z_x - 1) % p
        fp_, gp_ = 1, 1
        for j in range(NUM_ROUTED_WIRES):
            fp_ = fp_ * (local[j] + beta * (k_is[j] * x % p) + gamma) % p
            gp_ = gp_ * (local[j] + beta * sigma_8n[j][i] + gamma) % p
        shift = (fp_ * z_x - gp_ * z_gz) % p
        out.append(reduce_with_powers(field, [z1, shift] + terms, alpha))
    return out
