#!/usr/bin/env python3
"""Golden vectors for the BLAKE3 generator derivation (src/hash_to_curve.rs:13-76).

The reference hashes with the blake3 crate, which is not vendored in /root/reference.  This script is run in the
BUILD container, where the independent `blake3` Python package (the official Rust implementation's binding, 1.0.8)
is importable: it checks oracle/plonky_oracle.py's restatement of the BLAKE3 compression function against it and
writes tests/golden/blake_hash_to_curve.json:
  * raw XOF outputs of the package for the exact byte strings blake_field hashes,
  * the first generators blake_hash_usize_to_curve(seed) of each curve computed with the PACKAGE's hash
    (so the fixture does not depend on the restatement it pins).
The tests read only the JSON; nothing at test time needs the package.
"""
import json
import os
import sys

import blake3

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import plonky_oracle as po  # noqa: E402


def pkg_xof(data: bytes, n: int) -> bytes:
    return blake3.blake3(data).digest(length=n)


def blake_field_pkg(field, it, seed):
    nbytes = 8 * field.limbs
    j = 0
    while True:
        h = bytearray(pkg_xof(seed.to_bytes(nbytes, "little") + bytes([it, j]), nbytes + 1))
        h[nbytes - 1] >>= 8 * nbytes - field.bits
        x = int.from_bytes(h[:nbytes], "little")
        if x < field.p:
            return x, (h[nbytes] & 1) == 1
        j += 1


def hash_to_curve_pkg(curve, seed):
    f = curve.base
    i = 0
    while True:
        x, y_neg = blake_field_pkg(f, i, seed)
        y = f.sqrt((x * x * x + curve.a * x + curve.b) % f.p)
        if y is not None:
            return (x, (-y) % f.p if y_neg else y), i
        i += 1


def main():
    out = {"blake3_package_version": blake3.__version__, "xof": [], "generators": {}}
    for data in [b"", b"abc", bytes(range(34)), bytes(range(50)), bytes(64), (1 << 255).to_bytes(32, "little") + b"\x03\x01"]:
        want = pkg_xof(data, 64)
        assert po.blake3_xof_one_block(data, 64) == want
        out["xof"].append({"input": data.hex(), "out64": want.hex()})
    for curve in (po.TWEEDLEDEE, po.TWEEDLEDUM, po.BLS12_377):
        seeds = list(range(24)) + [1 << 12, (1 << 16) + 1, (1 << 20) - 1, (1 << 32) + 5]
        rows = []
        for s in seeds:
            (x, y), iters = hash_to_curve_pkg(curve, s)
            assert po.blake_hash_usize_to_curve(curve, s) == (x, y)
            assert curve.is_on_curve((x, y))
            rows.append({"seed": s, "x": hex(x), "y": hex(y), "iterations": iters})
        out["generators"][curve.name] = rows
    path = os.path.join(ROOT, "tests", "golden", "blake_hash_to_curve.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
