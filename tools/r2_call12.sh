#!/bin/bash
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "stem" ) > $O/r2c12_pytest_stem.log 2>&1
tail -3 $O/r2c12_pytest_stem.log; grep -E "^E  |FAILED" $O/r2c12_pytest_stem.log | head -10 | cut -c1-300
for bn in eval train; do
  timeout 300 python tools/profile_engine.py --batch 8 --reid 4096 --reid-bn $bn --out $O/r2c12_prof_r_${bn}.json > $O/r2c12_prof_r_${bn}.log 2>&1; grep "^reid" $O/r2c12_prof_r_${bn}.log; grep -E "stem|roi" $O/r2c12_prof_r_${bn}.log | head -4
done
