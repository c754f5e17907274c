"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, mean, total, share)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[start + 1:]:
    if len(r) > vi:
        try:
            d[r[ki].split("(")[0][:64]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
tot = sum(sum(v) for v in d.values())
print(f"{sum(len(v) for v in d.values())} launches, {tot / 1e3:.1f} us of kernel time")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:64s} n={len(v):4d} mean={sum(v) / len(v) / 1e3:8.1f} us total={sum(v) / 1e3:9.1f} us share={sum(v) / tot:.3f}")
