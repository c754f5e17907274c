"""Per-kernel summary of an `ncu --metrics ... --csv` capture of one forward (tools/prof_step.py): launches, device time,
DRAM / L2 bytes, tensor-pipe activity. Writes a JSON next to the CSV when given an output path."""
import csv, collections, json, sys
rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
idc, kn, mn, mv, mu = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launch = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) <= mv:
        continue
    d = launch.setdefault(r[idc], {"name": r[kn].split("(")[0].replace("void ", "")[:60]})
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    u = r[mu]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    d[r[mn]] = v * scale
agg = collections.OrderedDict()
for d in launch.values():
    a = agg.setdefault(d["name"], collections.defaultdict(float))
    a["launches"] += 1
    a["us"] += d.get("gpu__time_duration.sum", 0.0)
    a["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["l2_bytes"] += d.get("lts__t_bytes.sum", 0.0)
    a["tensor_active_x_us"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * d.get("gpu__time_duration.sum", 0.0)
    a["regs"] = max(a["regs"], d.get("launch__registers_per_thread", 0.0))
tot = sum(a["us"] for a in agg.values())
out = {"total_kernel_us": round(tot, 1), "n_launches": len(launch), "kernels": {}}
print(f"{len(launch)} launches, {tot:.1f} us of kernel time (serialised, cold caches)")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    k = dict(launches=int(a["launches"]), us_total=round(a["us"], 1), share=round(a["us"] / tot, 4), us_mean=round(a["us"] / a["launches"], 2),
             dram_bytes_per_launch=round(a["dram_bytes"] / a["launches"]), l2_bytes_per_launch=round(a["l2_bytes"] / a["launches"]),
             tensor_pipe_active_pct=round(a["tensor_active_x_us"] / a["us"], 1) if a["us"] else 0.0, registers=int(a["regs"]))
    out["kernels"][name] = k
    print(f"{name:60s} n={k['launches']:3d} {k['us_total']:8.1f} us {k['share']:.3f}  mean {k['us_mean']:7.1f} us  dram/launch {k['dram_bytes_per_launch'] / 1e6:7.2f} MB  "
          f"l2/launch {k['l2_bytes_per_launch'] / 1e6:8.2f} MB  tensor {k['tensor_pipe_active_pct']:4.1f}%  regs {k['registers']}")
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
