"""Print the per-kernel breakdown of a bench.py JSON line (default gpurun_out/bench.json)."""
import json, sys
d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bench.json"))
fb = d.get("fwd_bwd") or {}
print(f"value {d['value']} {d['unit']}  {d['ms_per_step']} ms/step  e2e {d['e2e']['value']}  fwd+bwd {fb.get('value')} ({fb.get('ms_per_step')} ms)  launches {d['gpu_launches']}")
for k, v in d.get("kernels", {}).items():
    print(f"{k:32s} share {v['share_of_step']:.3f}  {v['us_per_launch']:8.1f} us x{v['launches_per_step']:5.1f} = {v['us_per_launch'] * v['launches_per_step']:8.1f}  {v['achieved']:8.1f} {v['unit']} frac {v['frac']:.3f}")
for k in ("roofline", "roofline_gat_scatter", "cpu_baseline", "clocks"):
    print(k, d.get(k))
