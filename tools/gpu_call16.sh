#!/bin/bash
# final round-1 measurements: tests, bench at B=64/128, ncu launch list (time + DRAM bytes) of the timed steps, full captures
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c16_pytest.log 2>&1
tail -3 gpurun_out/c16_pytest.log | cut -c1-200
for B in 64 128; do
  timeout 300 python bench.py --batch $B --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c16_bench_b$B.json 2> gpurun_out/c16_bench_b$B.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/c16_bench_b$B.json"))
    print("B=$B", "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["achieved"]), d["clocks"])
except Exception as e:
    print("B=$B FAILED", e); print(open("gpurun_out/c16_bench_b$B.err").read()[-1500:])
P
done
VCB_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -c 600 --csv --log-file gpurun_out/c16_ncu_launches.csv python bench.py --batch 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c16_ncu_bench.log 2>&1
wc -l gpurun_out/c16_ncu_launches.csv
for name in fast-big-3x3-192-res auto-big-reid-l1 fast-big-1x1-96 auto-big-reid-l4; do
  idx=$(python - <<P
import sys; sys.path.insert(0,'tests')
import bringup_conv as b
print([i for i,(n,_) in enumerate(b.CASES) if n=="$name"][0])
P
)
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -f -o gpurun_out/c16_ncu_$name python tests/bringup_conv.py --case $idx > gpurun_out/c16_ncu_$name.log 2>&1
  tail -1 gpurun_out/c16_ncu_$name.log | cut -c1-200
done
timeout 300 python tools/profile_engine.py --batch 64 --reid 4096 --out gpurun_out/c16_profile_b64.json > gpurun_out/c16_profile_b64.log 2>&1
head -2 gpurun_out/c16_profile_b64.log; grep "^reid" gpurun_out/c16_profile_b64.log
