"""Top stall sites of an ncu --set full --import-source on capture (SASS view): python tools/ncu_hot.py report.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["# Samples"] or 0) for r in rows)
stalls = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
print("total samples", tot)
agg = {k: sum(int(r[k] or 0) for r in rows) for k in stalls}
print({k: round(v / max(tot, 1), 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot})
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:top]
for i in sorted(idx):
    r = rows[i]
    s = int(r["# Samples"] or 0)
    main = sorted(((int(r[k] or 0), k) for k in stalls), reverse=True)[:2]
    print(f"{i:5d} {100 * s / tot:5.1f}%  {r['Source'].strip()[:70]:70s} exec={r['Instructions Executed']:>9s}  " + " ".join(f"{k[6:]}={v}" for v, k in main if v))
