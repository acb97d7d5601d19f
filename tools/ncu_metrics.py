"""Selected raw metrics of ncu captures (development aid): python tools/ncu_metrics.py a.ncu-rep [b.ncu-rep ...]"""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum", "smsp__thread_inst_executed.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "launch__grid_size",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    h, u, v = r[0], r[1], r[-1]
    d = {k: (val, un) for k, un, val in zip(h, u, v)}
    print("==", rep, d.get("Kernel Name", ("",))[0][:60])
    for k in WANT:
        if k in d:
            print(f"   {k:80s} {d[k][0]:>16s} {d[k][1]}")
