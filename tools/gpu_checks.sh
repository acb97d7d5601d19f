#!/bin/bash
# End-of-change GPU checks (run under gpurun, one B200): parity tests, smoke, default bench, the ncu launch list of the timed
# steps (time + DRAM bytes + tensor-pipe activity per launch) and the per-layer event profile.  Outputs land in gpurun_out/;
# copy what should be judged into profiles/ (tools/summarize_launches.py turns the launch list into r01_launch_summary.json).
#   gpurun --timeout 1500 -- 'bash tools/gpu_checks.sh'
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2; grep -E "^E|FAILED" gpurun_out/pytest_gpu.log | head -5 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<P
import json
d = json.load(open("gpurun_out/bench_default.json"))
print("fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"], 3), d["clocks"])
P
VCB_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -c 600 --csv --log-file gpurun_out/ncu_launches.csv python bench.py --batch 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/ncu_launches.csv
timeout 300 python tools/profile_engine.py --batch 64 --reid 4096 --out gpurun_out/profile_engine.json > gpurun_out/profile_engine.log 2>&1
head -12 gpurun_out/profile_engine.log; grep -A6 "^reid" gpurun_out/profile_engine.log
