#!/usr/bin/env python3
"""Pretty-print the interesting numbers of a bench.py JSON line (stdin)."""
import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline", {})
    print(f"MSM: {d['value']:.4g} {d['unit']}  {d['ms_per_step']:.3f} ms/step  e2e {d.get('e2e',{}).get('value',0):.4g}  launches {d.get('gpu_launches')}")
    if r:
        print("  phases:", {k: round(v, 3) for k, v in r.get("phases_ms", {}).items()}, " frac", round(r.get("frac", 0), 5))
    n = d.get("ntt")
    if n and "inverse_ms" not in n:
        print(f"NTT (domain split): {n['value']:.4g} {n['unit']}  {n['ms_per_step']:.3f} ms  {n['config']['workload']}")
    elif n:
        print(f"NTT: {n['value']:.4g} {n['unit']}  fwd {n['ms_per_step']:.3f} ms  inv {n['inverse_ms']:.3f}  lde {n['coset_lde_ms']:.3f}  passes {[round(x,3) for x in n['roofline']['pass_ms']]}  frac {n['roofline']['frac']:.4f}  e2e {n['e2e']['ms_per_step']:.2f} ms")
    if d.get("clocks"):
        print("  clocks:", d["clocks"])
    if d.get("cpu_baseline"):
        print("  cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
