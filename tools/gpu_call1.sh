#!/bin/bash
# round-1 re-entry check: GPU tests, bench at several batch sizes, ncu full captures of representative conv launches
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.log 2>&1
tail -3 gpurun_out/c1_pytest.log
for B in 32 64 128; do
  timeout 300 python bench.py --batch $B --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_b$B.json 2> gpurun_out/c1_bench_b$B.err
  tail -c 600 gpurun_out/c1_bench_b$B.json
done
timeout 300 python tools/profile_engine.py --batch 64 --reid 4096 --out gpurun_out/c1_profile_b64.json > gpurun_out/c1_profile_b64.log 2>&1
timeout 300 python tools/profile_engine.py --batch 128 --reid 8192 --out gpurun_out/c1_profile_b128.json > gpurun_out/c1_profile_b128.log 2>&1
# ncu --set full on three representative layer shapes (third launch of the conv kernel in each bring-up case)
i=0
for name in tiled-1x1-192-192 prof-3x3-192-192-p4 prof-1x1-96-96-m819k tma-bk16-s2dstem; do
  idx=$(python - <<P
import sys; sys.path.insert(0,'tests')
import bringup_conv as b
print([i for i,(n,_) in enumerate(b.CASES) if n=="$name"][0])
P
)
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -f -o gpurun_out/c1_ncu_$name python tests/bringup_conv.py --case $idx > gpurun_out/c1_ncu_$name.log 2>&1
  tail -2 gpurun_out/c1_ncu_$name.log
done
ls -la gpurun_out
