#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c32_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/c32_pytest.log | tail -2; grep -E "^E|FAILED" gpurun_out/c32_pytest.log | head -5 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
( time python bench.py ) > gpurun_out/c32_bench_default.json 2> gpurun_out/c32_bench_default.err
python - <<P
import json
d=json.load(open("gpurun_out/c32_bench_default.json"))
print("fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"],3), "kernels", d["kernels_per_step"], d["clocks"])
P
VCB_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 600 --csv --log-file gpurun_out/c32_ncu_launches.csv python bench.py --batch 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c32_ncu_bench.log 2>&1
wc -l gpurun_out/c32_ncu_launches.csv
timeout 300 python tools/profile_engine.py --batch 64 --reid 4096 --out gpurun_out/c32_profile_b64.json > gpurun_out/c32_profile_b64.log 2>&1
head -3 gpurun_out/c32_profile_b64.log; grep "^reid" gpurun_out/c32_profile_b64.log
