"""Wire formats of the two precomputations.

The reference derives `Serialize / Deserialize` for `MsmPrecomputation<C>` (src/curve/curve_msm.rs:16-25) and
`FftPrecomputation<F>` (src/fft.rs:28-34) and stores both inside `VerificationKey` (src/verifier.rs:15-27), which
travels as CBOR (serde_cbor, src/serialization.rs:254-328).  With serde's derive a struct is a CBOR map keyed by the
field names in declaration order, a `Vec` a definite-length array, `usize` an unsigned integer, and field elements /
affine points are byte strings (`serialize_bytes` of ToBytes, src/serialization.rs:76-148):

    MsmPrecomputation  {"powers_per_generator": [[bytes(1 + 8L); DIGITS]; n], "w": uint}
                       powers_per_generator[i][j] = [(2^w)^j] G_i compressed (mask byte + canonical x), DIGITS = ceil(BITS / w)
    FftPrecomputation  {"subgroups_rev": [[bytes(8L); 2^i]; i = 0 ..= log2 n]}
                       subgroups_rev[i][k] = w_i^(reverse_bits(k, i)), w_i = primitive_root_of_unity(i)

The device keeps neither structure (its table uses its own window, its plan holds no n-entry table): export
recomputes the reference's contents from (generators, w) / (field, degree) with the device kernels, import keeps
`powers_per_generator[i][0]` (the generators) and `w`, resp. the degree, and rebuilds the device handle -- what
INTEGRATION.md's `GpuTable` / `GpuPlan` wrappers do to stay `Clone + Serialize + Deserialize + PartialEq`.
"""
from __future__ import annotations

import numpy as np

from . import (FIELD_LIMBS, CURVE_BASE_FIELD, CURVE_SCALAR_FIELD, MsmPrecomputation, FftPrecomputation, msm_precompute_affine,
               fft_precompute, fft_with_precomputation_power_of_2, field_to_bytes, field_from_bytes, field_op, points_to_bytes,
               points_from_bytes, curve_mul, PlonkyPanic)

__all__ = ["cbor_dumps", "cbor_loads", "msm_precomputation_to_cbor", "msm_precomputation_from_cbor", "fft_precomputation_to_cbor",
           "fft_precomputation_from_cbor", "SCALAR_BITS"]

SCALAR_BITS = {0: 255, 1: 255, 2: 253}        # C::ScalarField::BITS per curve id
_FIELD_ONE_CACHE = {}


# ---- the subset of RFC 8949 serde_cbor emits for these types: uint, bytes, text, array, map, null ------------------
def _head(major: int, n: int) -> bytes:
    if n < 24:
        return bytes([(major << 5) | n])
    for code, size in ((24, 1), (25, 2), (26, 4), (27, 8)):
        if n < (1 << (8 * size)):
            return bytes([(major << 5) | code]) + n.to_bytes(size, "big")
    raise ValueError("integer too large for CBOR")


def cbor_dumps(obj) -> bytes:
    out = bytearray()

    def enc(o):
        if o is None:
            out.append(0xF6)
        elif isinstance(o, bool):
            out.append(0xF5 if o else 0xF4)
        elif isinstance(o, int):
            if o < 0:
                raise ValueError("negative integers do not occur in these formats")
            out.extend(_head(0, o))
        elif isinstance(o, (bytes, bytearray, memoryview)):
            out.extend(_head(2, len(o)))
            out.extend(o)
        elif isinstance(o, str):
            b = o.encode()
            out.extend(_head(3, len(b)))
            out.extend(b)
        elif isinstance(o, (list, tuple)):
            out.extend(_head(4, len(o)))
            for x in o:
                enc(x)
        elif isinstance(o, dict):                    # insertion order = serde's declaration order
            out.extend(_head(5, len(o)))
            for k, v in o.items():
                enc(k)
                enc(v)
        else:
            raise TypeError(f"cannot encode {type(o)}")
    enc(obj)
    return bytes(out)


def cbor_loads(data: bytes):
    pos = 0

    def arg(info):
        nonlocal pos
        if info < 24:
            return info
        if info > 27:
            raise ValueError("indefinite lengths are not produced by serde_cbor for these types")
        size = 1 << (info - 24)
        v = int.from_bytes(data[pos:pos + size], "big")
        pos += size
        return v

    def dec():
        nonlocal pos
        b = data[pos]
        pos += 1
        major, info = b >> 5, b & 31
        if major == 0:
            return arg(info)
        if major == 2:
            n = arg(info)
            v = bytes(data[pos:pos + n])
            pos += n
            return v
        if major == 3:
            n = arg(info)
            v = bytes(data[pos:pos + n]).decode()
            pos += n
            return v
        if major == 4:
            return [dec() for _ in range(arg(info))]
        if major == 5:
            n = arg(info)
            d = {}
            for _ in range(n):
                k = dec()
                d[k] = dec()
            return d
        if b == 0xF6:
            return None
        if b in (0xF4, 0xF5):
            return b == 0xF5
        raise ValueError(f"unsupported CBOR item 0x{b:02x}")
    v = dec()
    if pos != len(data):
        raise ValueError("trailing bytes")
    return v


# ---- MsmPrecomputation -------------------------------------------------------------------------------------------
def _one(field: int) -> np.ndarray:
    """ONE in Montgomery form (R mod p): from_canonical(1) on the device."""
    if field not in _FIELD_ONE_CACHE:
        c = np.zeros((1, FIELD_LIMBS[field]), dtype=np.uint64)
        c[0, 0] = 1
        _FIELD_ONE_CACHE[field] = field_op(field, "from_canonical", c)[0]
    return _FIELD_ONE_CACHE[field]


def msm_precomputation_to_cbor(curve: int, generators_xy, w: int, zero=None) -> bytes:
    """serde_cbor::to_vec(&msm_precompute(generators, w)) (curve_msm.rs:27-52): every power [(2^w)^j] G_i is recomputed on
    the device (plk_curve_mul with the scalars 2^(w j)) and compressed with AffinePoint::write."""
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    sf = CURVE_SCALAR_FIELD[curve]
    g = np.ascontiguousarray(generators_xy, dtype=np.uint64).reshape(-1, 2, Lb)
    n = g.shape[0]
    z = np.zeros(n, dtype=np.uint8) if zero is None else np.ascontiguousarray(zero, dtype=np.uint8)
    digits = -(-SCALAR_BITS[curve] // w)                              # curve_msm.rs:40 ceil(BITS / w)
    canon = np.zeros((digits, 4), dtype=np.uint64)
    for j in range(digits):
        e = w * j
        canon[j, e // 64] = np.uint64(1) << np.uint64(e % 64)
    pw = field_op(sf, "from_canonical", canon)                        # 2^(w j) as scalar-field elements
    xyz = np.zeros((n, digits, 3, Lb), dtype=np.uint64)
    xyz[:, :, :2] = g[:, None]
    xyz[:, :, 2] = _one(CURVE_BASE_FIELD[curve])
    scal = np.broadcast_to(pw[None], (n, digits, 4))
    zz = np.repeat(z, digits)
    out, oz = curve_mul(curve, xyz.reshape(-1, 3, Lb), np.ascontiguousarray(scal).reshape(-1, 4), zz)
    enc = points_to_bytes(curve, out[:, :2], oz).reshape(n, digits, -1)
    return cbor_dumps({"powers_per_generator": [[enc[i, j].tobytes() for j in range(digits)] for i in range(n)], "w": int(w)})


def msm_precomputation_from_cbor(curve: int, data: bytes, check_powers: bool = False) -> MsmPrecomputation:
    """serde_cbor::from_slice::<MsmPrecomputation<C>>: the generators are powers_per_generator[i][0]; the device table is
    rebuilt from them.  check_powers=True also verifies every stored power against a recomputation."""
    d = cbor_loads(data)
    if not isinstance(d, dict) or list(d.keys()) != ["powers_per_generator", "w"]:
        raise ValueError("not an MsmPrecomputation")
    w = int(d["w"])
    ppg = d["powers_per_generator"]
    Lb = FIELD_LIMBS[CURVE_BASE_FIELD[curve]]
    stride = 1 + 8 * Lb
    first = np.frombuffer(b"".join(p[0] if p else bytes([1]) + bytes(stride - 1) for p in ppg), dtype=np.uint8).reshape(-1, stride) \
        if ppg else np.zeros((0, stride), dtype=np.uint8)
    # AffinePoint::read consumes only the mask byte of a zero point (serialization.rs:51-57); `write` always emits 1 + 8L
    g, z = points_from_bytes(curve, first) if len(ppg) else (np.zeros((0, 2, Lb), dtype=np.uint64), np.zeros(0, dtype=np.uint8))
    if check_powers and data != msm_precomputation_to_cbor(curve, g, w, z):
        raise ValueError("stored powers do not match the generators")
    pre = msm_precompute_affine(curve, g, w, z)
    pre.generators, pre.zero = g, z
    return pre


# ---- FftPrecomputation -------------------------------------------------------------------------------------------
def _bitrev_indices(bits: int) -> np.ndarray:
    idx = np.arange(1 << bits, dtype=np.uint64)
    out = np.zeros_like(idx)
    for i in range(bits):
        out |= ((idx >> np.uint64(i)) & np.uint64(1)) << np.uint64(bits - 1 - i)
    return out.astype(np.int64)


def fft_precomputation_to_cbor(field: int, degree: int) -> bytes:
    """serde_cbor::to_vec(&fft_precompute::<F>(degree)) (fft.rs:47-59).  The top-level subgroup w^k is the device transform
    of the unit vector e_1; level i is its stride-2^(L - i) subsequence, bit-reversed (fft.rs:8-26)."""
    L = FIELD_LIMBS[field]
    pre = fft_precompute(field, degree)
    try:
        n = pre.size()
        e1 = np.zeros((n, L), dtype=np.uint64)
        if n > 1:
            e1[1] = _one(field)
            sub = fft_with_precomputation_power_of_2(e1, pre)                # sub[k] = w^k
        else:
            sub = _one(field).reshape(1, L).copy()
    finally:
        pre.close()
    raw = field_to_bytes(field, sub)                                         # (n, 8L) canonical little-endian
    logn = n.bit_length() - 1
    levels = []
    for i in range(logn + 1):
        lvl = raw[::1 << (logn - i)][_bitrev_indices(i)]
        levels.append([lvl[k].tobytes() for k in range(1 << i)])
    return cbor_dumps({"subgroups_rev": levels})


def fft_precomputation_from_cbor(field: int, data: bytes, check: bool = True) -> FftPrecomputation:
    """serde_cbor::from_slice::<FftPrecomputation<F>>: the size is subgroups_rev.last().len() (fft.rs:36-40); the device plan
    is rebuilt for it.  check=True verifies that the stored top-level subgroup is the one this field's
    primitive_root_of_unity generates (element 1 of the bit-reversed list is w^(n/2) = -1, element 2 is w^(n/4), ...)."""
    d = cbor_loads(data)
    if not isinstance(d, dict) or list(d.keys()) != ["subgroups_rev"] or not d["subgroups_rev"]:
        raise ValueError("not an FftPrecomputation")
    levels = d["subgroups_rev"]
    n = len(levels[-1])
    if n & (n - 1) or len(levels) != n.bit_length():
        raise PlonkyPanic("Not a power of two")
    L = FIELD_LIMBS[field]
    for lvl in levels:                                                       # FromBytes of every element ("Out of range" check)
        field_from_bytes(field, np.frombuffer(b"".join(lvl), dtype=np.uint8).reshape(-1, 8 * L))
    if check and data != fft_precomputation_to_cbor(field, n):
        raise ValueError("stored subgroups do not match this field's roots of unity")
    return fft_precompute(field, n)
