#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors for the MSM/NTT hot path into
tests/golden/reference_kats.json.

Runs ONLY in the build container (needs /root/reference).  The JSON it writes is committed;
tests read the JSON, never the reference tree.  Everything extracted is a literal constant or a
literal test vector of the reference (file:line recorded per entry) -- no reference code is
executed (the reference is Rust; there is no Rust toolchain in this image).
"""
import json
import os
import re
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "reference_kats.json")


def read(path):
    with open(os.path.join(REF, path)) as f:
        return f.read()


def line_of(text, pos):
    return text.count("\n", 0, pos) + 1


def grab_array(text, name, path):
    """const NAME: [u64; N] = [ ... ];  or  const NAME: Self = Self { limbs: [ ... ] };"""
    m = re.search(r"const\s+" + name + r"\s*:\s*[^=]*=\s*(?:Self\s*\{\s*limbs\s*:\s*)?\[([^\]]*)\]", text, re.S)
    if not m:
        raise KeyError(f"{name} not found in {path}")
    body = m.group(1)
    if ";" in body and "," not in body:        # [0; 6]
        v, n = body.split(";")
        vals = [int(v.strip().replace("u64", ""))] * int(n.strip())
    else:
        vals = [int(x.strip().replace("u64", ""), 0) for x in body.split(",") if x.strip()]
    return {"limbs": [str(v) for v in vals], "src": f"{path}:{line_of(text, m.start())}"}


def grab_scalar(text, name, path):
    m = re.search(r"const\s+" + name + r"\s*:\s*\w+\s*=\s*(\d+)\s*;", text)
    if not m:
        raise KeyError(f"{name} not found in {path}")
    return {"value": m.group(1), "src": f"{path}:{line_of(text, m.start())}"}


def field_entry(path):
    t = read(path)
    e = {"file": path}
    for name in ("ORDER", "R", "R2", "R3"):
        e[name] = grab_array(t, name, path)
    e["MU"] = grab_scalar(t, "MU", path)
    for name in ("ORDER_X2", "TWO", "THREE", "FOUR", "FIVE", "NEG_ONE", "T", "MULTIPLICATIVE_SUBGROUP_GENERATOR"):
        try:
            e[name] = grab_array(t, name, path)
        except KeyError:
            pass
    for name in ("BITS", "TWO_ADICITY"):
        e[name] = grab_scalar(t, name, path)
    return e


def limb_literal_after(text, anchor, path, count=1):
    """Find `count` `limbs: [...]` literals following `anchor`."""
    pos = text.index(anchor)
    out = []
    for _ in range(count):
        m = re.compile(r"limbs\s*:\s*\[([^\]]*)\]", re.S).search(text, pos)
        vals = [int(x.strip()) for x in m.group(1).split(",") if x.strip()]
        out.append({"limbs": [str(v) for v in vals], "src": f"{path}:{line_of(text, m.start())}"})
        pos = m.end()
    return out


def main():
    if not os.path.isdir(REF):
        print("no /root/reference here; nothing to do", file=sys.stderr)
        return 1
    kats = {"_comment": "Extracted by tools/extract_reference_kats.py from 0xPolygonZero/plonky; literals only."}

    kats["fields"] = {
        "TweedledeeBase": field_entry("src/field/tweedledee_base.rs"),
        "TweedledumBase": field_entry("src/field/tweedledum_base.rs"),
        "Bls12377Scalar": field_entry("src/field/bls12_377_scalar.rs"),
        "Bls12377Base": field_entry("src/field/bls12_377_base.rs"),
    }

    # curve constants ---------------------------------------------------------------
    p = "src/curve/tweedledee_curve.rs"
    t = read(p)
    z = limb_literal_after(t, "const ZETA:", p, 1) + limb_literal_after(t, "const ZETA_SCALAR:", p, 1)
    kats["curves"] = {"Tweedledee": {"A": "ZERO", "B": "FIVE", "gen_x": "NEG_ONE", "gen_y": "TWO",
                                     "ZETA": z[0], "ZETA_SCALAR": z[1], "src": f"{p}:7-38"}}
    p = "src/curve/tweedledum_curve.rs"
    t = read(p)
    b = limb_literal_after(t, "const B:", p, 1)[0]
    gy = limb_literal_after(t, "x: TweedledumBase::ONE", p, 1)[0]
    z = limb_literal_after(t, "const ZETA:", p, 1) + limb_literal_after(t, "const ZETA_SCALAR:", p, 1)
    kats["curves"]["Tweedledum"] = {"A": "ZERO", "B": b, "gen_x": "ONE", "gen_y": gy,
                                    "ZETA": z[0], "ZETA_SCALAR": z[1], "src": f"{p}:7-53"}
    p = "src/curve/bls12_377_curve.rs"
    t = read(p)
    gx = limb_literal_after(t, "const BLS12_377_GENERATOR_X", p, 1)[0]
    gy = limb_literal_after(t, "const BLS12_377_GENERATOR_Y", p, 1)[0]
    kats["curves"]["Bls12377"] = {"A": "ZERO", "B": "ONE", "gen_x": gx, "gen_y": gy, "src": f"{p}:10-33"}

    # to_digits KAT (src/curve/curve_msm.rs test_to_digits) ----------------------------
    p = "src/curve/curve_msm.rs"
    t = read(p)
    i0 = t.index("fn test_to_digits")
    i1 = t.index("fn test_msm")
    body = t[i0:i1]
    bins = re.findall(r"0b([01]+)", body)
    x_canonical = [str(int(b, 2)) for b in bins[:4]]
    digits = [str(int(b, 2)) for b in bins[4:]]
    w = int(re.search(r"to_digits::<Bls12377>\(&x,\s*(\d+)\)", body).group(1))
    kats["to_digits"] = {"field": "Bls12377Scalar", "x_canonical": x_canonical, "w": w, "digits": digits,
                         "src": f"{p}:{line_of(t, i0)}"}
    # test_msm inputs (self-consistency test; expected value = naive sum)
    body = t[i1:]
    sc = re.findall(r"from_canonical\(\[([0-9, ]+)\]\)", body)
    kats["test_msm"] = {"curve": "Bls12377", "w": int(re.search(r"let w = (\d+);", body).group(1)),
                        "generators": ["G", "2G", "3G"],
                        "scalars_canonical": [[s.strip() for s in x.split(",")] for x in sc],
                        "src": f"{p}:{line_of(t, i1)}"}

    # div2 KAT (src/bigint/bigint_arithmetic.rs test_div2) -----------------------------
    p = "src/bigint/bigint_arithmetic.rs"
    t = read(p)
    i0 = t.index("fn test_div2")
    arrays = re.findall(r"\[([0-9,\s]+)\]", t[i0:])
    arrays = [[str(int(v)) for v in a.split(",") if v.strip()] for a in arrays]
    kats["div2"] = {"cases": [{"in": arrays[0], "out": arrays[1]}, {"in": arrays[2], "out": arrays[3]}],
                    "src": f"{p}:{line_of(t, i0)}"}

    # reverse_bits KATs (src/fft.rs test_reverse_bits) ---------------------------------
    p = "src/fft.rs"
    t = read(p)
    i0 = t.index("fn test_reverse_bits")
    m = re.search(r"reverse_bits\(0b([01]+),\s*(\d+)\),\s*0b([01]+)", t[i0:])
    kats["reverse_bits"] = {"n": int(m.group(1), 2), "bits": int(m.group(2)), "out": int(m.group(3), 2),
                            "index_perm_4": ["a", "c", "b", "d"], "src": f"{p}:{line_of(t, i0)}"}
    # fft_and_ifft deterministic input recipe
    i0 = t.index("fn fft_and_ifft")
    deg = int(re.search(r"let degree = (\d+);", t[i0:]).group(1))
    mm = re.search(r"from_canonical_usize\(i \* (\d+) % (\d+)\)", t[i0:])
    kats["fft_and_ifft"] = {"field": "Bls12377Scalar", "degree": deg, "mul": int(mm.group(1)),
                            "mod": int(mm.group(2)), "src": f"{p}:{line_of(t, i0)}"}

    # montgomery round-trip / multiply vectors (bls12_377_base.rs / _scalar.rs tests) ------
    for key, p in (("Bls12377Base", "src/field/bls12_377_base.rs"), ("Bls12377Scalar", "src/field/bls12_377_scalar.rs")):
        t = read(p)
        i0 = t.index("#[cfg(test)]")
        lits = re.findall(r"let (\w+) = \[([0-9, ]+)\];", t[i0:])
        kats.setdefault("mont_mul_inputs", {})[key] = {
            name: [v.strip() for v in vals.split(",")] for name, vals in lits[:4]}
        kats["mont_mul_inputs"][key]["src"] = f"{p}:{line_of(t, i0)}"

    # MSM / FFT tunables that define behaviour
    kats["constants"] = {
        "DIGITS_PER_CHUNK": int(re.search(r"DIGITS_PER_CHUNK: usize = (\d+)", read("src/curve/curve_msm.rs")).group(1)),
        "summation_threshold": int(re.search(r"pairwise_sums < (\d+)", read("src/curve/curve_summations.rs")).group(1)),
        "fft_chunk": int(re.search(r"par_chunks\((\d+)\)", read("src/fft.rs")).group(1)),
        "mul_window_bits": int(re.search(r"WINDOW_BITS: usize = (\d+)", read("src/curve/curve_multiplication.rs")).group(1)),
    }

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(kats, f, indent=1, sort_keys=True)
    print("wrote", os.path.normpath(OUT))
    return 0


if __name__ == "__main__":
    sys.exit(main())
