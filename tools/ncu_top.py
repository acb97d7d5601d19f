"""Top stall sites of one ncu capture (development aid): python tools/ncu_top.py <report.ncu-rep> [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(h)]
ci = {k: i for i, k in enumerate(h)}
stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {k: sum(int(r[ci[k]] or 0) for r in body) for k in stall_cols}
print("by reason:", {k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
idx = sorted(range(len(body)), key=lambda i: -int(body[i][ci["# Samples"]] or 0))[:N]
for i in sorted(idx):
    r = body[i]
    st = {k[6:]: int(r[ci[k]]) for k in stall_cols if int(r[ci[k]] or 0) > 0}
    top = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(f"{i:5d} {int(r[ci['# Samples']]):6d} {100.0*int(r[ci['# Samples']])/tot:5.1f}%  {r[ci['Source']].strip()[:70]:70s} {top}")
