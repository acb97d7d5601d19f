"""Key figures of every kernel in an `ncu --set full` report: duration, DRAM bytes and achieved DRAM GB/s, tensor-pipe activity,
L2 -> SM bytes.  python tools/ncu_kernel_summary.py report.ncu-rep [algorithmic_bytes] -> one JSON line per kernel"""
import csv, io, json, subprocess, sys
rep = sys.argv[1]
alg = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
names, units = rows[hdr], rows[hdr + 1]
want = {"gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_bytes", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__grid_size": "grid", "launch__block_size": "block",
        "launch__registers_per_thread": "regs"}
mult = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6,
        "Gbyte": 1e9}
for r in rows[hdr + 2:]:
    if len(r) != len(names):
        continue
    d = {"kernel": r[names.index("Kernel Name")][:90]}
    for i, n in enumerate(names):
        if n in want and r[i] not in ("", "n/a"):
            v = float(r[i].replace(",", ""))
            d[want[n]] = v * mult.get(units[i], 1.0) if units[i] in mult else v
    if "duration" in d and "dram_read" in d:
        d["dram_GBps"] = (d["dram_read"] + d.get("dram_write", 0.0)) / d["duration"] / 1e9
        d["duration_us"] = d.pop("duration") * 1e6
        if alg:
            d["algorithmic_bytes"] = alg
            d["algorithmic_GBps"] = alg / (d["duration_us"] * 1e-6) / 1e9
    print(json.dumps(d))
