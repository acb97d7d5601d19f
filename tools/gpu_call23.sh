#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c23_pytest.log 2>&1
grep -E "passed|failed|Error" gpurun_out/c23_pytest.log | tail -3; tail -12 gpurun_out/c23_pytest.log | grep -E "^E|assert" | head -8 | cut -c1-300
VCB_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 600 --csv --log-file gpurun_out/c23_ncu_launches.csv python bench.py --batch 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c23_ncu_bench.log 2>&1
wc -l gpurun_out/c23_ncu_launches.csv
( time python bench.py ) > gpurun_out/c23_bench_default.json 2> gpurun_out/c23_bench_default.err; tail -c 300 gpurun_out/c23_bench_default.json; tail -3 gpurun_out/c23_bench_default.err
