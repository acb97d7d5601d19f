"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor...` launch
list (one row per launch and metric) into per-kernel totals.  python tools/summarize_launches.py in.csv out.json steps"""
import csv, json, re, sys
from collections import defaultdict

src, dst, steps = sys.argv[1], sys.argv[2], int(sys.argv[3])
workload = json.loads(sys.argv[4]) if len(sys.argv) > 4 else None     # [model, h, w, batch, rois, reid_bn]: bench.py matches it
rows = [r for r in csv.reader(open(src)) if len(r) >= 15 and r[0].isdigit()]
per = defaultdict(dict)
name = {}
for r in rows:
    per[int(r[0])][r[12]] = float(r[14].replace(",", ""))
    name[int(r[0])] = re.sub(r"\(.*", "", r[4]).replace("vcb::", "").replace("void ", "")
agg = defaultdict(lambda: {"launches": 0, "us": 0.0, "dram_read_MB": 0.0, "dram_write_MB": 0.0, "tensor_pct_x_us": 0.0})
for i, m in per.items():
    a = agg[name[i]]
    us = m.get("gpu__time_duration.sum", 0.0) / 1e3
    a["launches"] += 1; a["us"] += us
    a["dram_read_MB"] += m.get("dram__bytes_read.sum", 0.0) / 1e6
    a["dram_write_MB"] += m.get("dram__bytes_write.sum", 0.0) / 1e6
    a["tensor_pct_x_us"] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * us
tot = sum(a["us"] for a in agg.values())
out = {"source": src, "workload": workload, "steps_captured": steps, "launches": len(per), "total_us": round(tot, 1), "by_kernel": []}
for k, a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
    out["by_kernel"].append({"kernel": k, "launches": a["launches"], "total_us": round(a["us"], 1), "share": round(a["us"] / tot, 4),
                             "dram_read_MB_per_step": round(a["dram_read_MB"] / steps, 1), "dram_write_MB_per_step": round(a["dram_write_MB"] / steps, 1),
                             "tensor_pipe_active_pct_time_weighted": round(a["tensor_pct_x_us"] / max(a["us"], 1e-9), 1)})
conv = [b for b in out["by_kernel"] if "conv_umma" in b["kernel"] or "conv_patch" in b["kernel"] or "reid_stem_pool" in b["kernel"] or "reid_stem_direct" in b["kernel"]]
out["conv_kernels"] = {"launches_per_step": sum(b["launches"] for b in conv) / steps, "us_per_step": round(sum(b["total_us"] for b in conv) / steps, 1),
                       "dram_bytes_per_step": int(sum(b["dram_read_MB_per_step"] + b["dram_write_MB_per_step"] for b in conv) * 1e6),
                       "share_of_step": round(sum(b["total_us"] for b in conv) / tot, 4),
                       "tensor_pipe_active_pct_time_weighted": round(sum(b["tensor_pipe_active_pct_time_weighted"] * b["total_us"] for b in conv) / max(sum(b["total_us"] for b in conv), 1e-9), 1)}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out["conv_kernels"]))
for b in out["by_kernel"]:
    print(b)
