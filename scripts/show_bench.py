#!/usr/bin/env python
"""Pretty-print the interesting parts of a bench.py JSON line: python scripts/show_bench.py gpurun_out/x.json"""
import json
import sys

l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"N={l['n_gpus']} value {l['value'] / 1e6:.2f}M ex/s  {l['ms_per_step']:.4f} ms/step   e2e {l['e2e']['value'] / 1e6:.2f}M   launches/step {l['gpu_launches'] // l['steps']}  clocks {l['clocks']}")
print(" timing", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in l["timing"].items() if k != "note"})
for k in ("step_phases_ms", "dp_phases_ms"):
    if l.get(k):
        print(" " + k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in l[k].items() if a != "note"})
r = l.get("roofline")
if r:
    print(f" softmax: {r['ms']:.4f} ms  {r['achieved']:.0f} TF/s  frac {r['frac']:.3f}  traffic {r['traffic']}")
for h in l.get("hbm_kernels", []):
    print(f" hbm: {h['kernel'][:60]:60s} {h['ms']:.4f} ms {h['achieved']:.0f} GB/s frac {h['frac']:.3f}")
c3 = l.get("c3")
if c3:
    rr = c3.get("roofline") or {}
    print(f" c3: {c3['value'] / 1e6:.2f}M ex/s {c3['ms_per_step']:.3f} ms/step e2e {c3['e2e']['value'] / 1e6:.2f}M  softmax {rr.get('ms', 0):.3f} ms frac {rr.get('frac', 0):.3f}")
ix = l["index"]
print(f" index105k: {ix['value'] / 1e6:.2f}M q/s {ix['ms_per_batch']:.4f} ms  e2e {ix['e2e']['value'] / 1e6:.2f}M  search {ix['roofline']['ms']:.4f} ms frac {ix['roofline']['frac']:.3f} launches {ix['gpu_launches_per_batch']}")
print("   stages", {k: round(v, 4) for k, v in (ix.get("stage_ms") or {}).items()})
if ix.get("row_sharded"):
    rs = ix["row_sharded"]
    print(f"   row_sharded105k: {rs['value'] / 1e6:.2f}M q/s {rs['ms_per_batch']:.4f} ms")
for leg in ix.get("row_sharded_large", []):
    print(f" index {leg['rows']:.0e}: {leg['value'] / 1e3:.0f}k q/s {leg['ms_per_batch']:.3f} ms e2e {leg['e2e']['value'] / 1e3:.0f}k frac {leg['roofline']['frac']:.3f} parity {leg.get('parity_ok')} build {leg['shard_build_s']:.2f}s")
    print("   stages", {k: round(v, 4) for k, v in (leg.get("stage_ms") or {}).items()})
if "parity" in l:
    print(" parity_ok", l.get("parity_ok"), {k: (v.get("ok"), v.get("max_update_err_rel_to_max_update", v.get("dense_gradient_l2_rel_err")), v.get("error")) for k, v in l["parity"].items()})
if "cpu_baseline" in l:
    print(" cpu", l["cpu_baseline"]["value"], "index cpu", ix.get("cpu_baseline", {}).get("value"))
