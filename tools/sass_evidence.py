"""Per-kernel static evidence from the built library: registers / spill (cuobjdump --dump-resource-usage) and the count of
tensor-core / TMA / TMEM SASS instructions (cuobjdump -sass). Writes profiles/sass_evidence_r<round>.txt (argv[1], default 2). Runs without a GPU."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cvpr2023-vlsat_b200", "libvlsat_b200.so")
pat = {"UTC*MMA (tcgen05.mma)": r"\bUTC[A-Z]*MMA", "UTMALDG (TMA load)": r"\bUTMALDG", "UTMASTG (TMA store)": r"\bUTMASTG",
       "LDTM/STTM (tcgen05.ld/st)": r"\b(LDTM|STTM)", "UTCBAR (tcgen05.commit)": r"\bUTCBAR", "SYNCS (mbarrier)": r"\bSYNCS",
       "ATOMG/RED (global atomics)": r"\b(ATOMG|RED)\b", "LDL/STL (local memory)": r"\b(LDL|STL)\b"}
def demangle(n):
    d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    if d.endswith(")"):                      # drop the trailing parameter list, keep template arguments such as (vlsat::tc::Kind)1
        depth = 0
        for i in range(len(d) - 1, -1, -1):
            depth += d[i] == ")"
            depth -= d[i] == "("
            if depth == 0:
                return d[:i]
    return d
res = {}
out = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
for m in re.finditer(r"Function (\S+):\n\s*(.*)", out):
    f = dict(kv.split(":") for kv in m.group(2).split() if ":" in kv)
    res[m.group(1)] = f
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, cur = collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        for k, p in pat.items():
            if re.search(p, line):
                counts[cur][k] += 1
rows = []
for fn, f in res.items():
    c = counts.get(fn, {})
    rows.append((demangle(fn), f.get("REG", "?"), f.get("STACK", "?"), f.get("SHARED", "?"), c))
rows.sort(key=lambda r: r[0])
with open(os.path.join(ROOT, "profiles", f"sass_evidence_r{sys.argv[1] if len(sys.argv) > 1 else 2}.txt"), "w") as fo:
    fo.write("libvlsat_b200.so (sm_100a), static per-kernel evidence: registers / stack bytes / static smem, then SASS instruction counts\n")
    fo.write("(dynamic shared memory is set at launch; UTC*MMA = tcgen05.mma, UTMALDG/UTMASTG = TMA, LDTM/STTM = tcgen05.ld/st)\n\n")
    for name, reg, st, sh, c in rows:
        extra = "  ".join(f"{k.split(' ')[0]}={v}" for k, v in c.items())
        fo.write(f"{name:<110s} REG {reg:>3s} STACK {st:>4s} SMEM {sh:>5s}  {extra}\n")
print(len(rows), "kernels")
