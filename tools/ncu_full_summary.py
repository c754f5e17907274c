"""Key metrics of `ncu --set full` reports (read here with `ncu -i <rep> --page raw --csv`) -> one JSON for profiles/."""
import csv, json, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STALL = "smsp__average_warps_issue_stalled_"
out = {}
for label, rep in (a.split("=", 1) for a in sys.argv[2:]):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    d = {"kernel": vals[hdr.index("Kernel Name")][:90], "report": rep}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            try:
                d[h + (" [" + u + "]" if u else "")] = float(v.replace(",", ""))
            except ValueError:
                pass
        elif h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
            try:
                d.setdefault("stall_per_issue", {})[h[len(STALL):-len("_per_issue_active.ratio")]] = round(float(v), 3)
            except ValueError:
                pass
    d["stall_per_issue"] = dict(sorted(d.get("stall_per_issue", {}).items(), key=lambda kv: -kv[1])[:6])
    out[label] = d
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1)[:3000])
