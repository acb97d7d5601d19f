#!/bin/bash
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "stem" ) > $O/r2c11_pytest_stem.log 2>&1
tail -3 $O/r2c11_pytest_stem.log; grep -E "^E  |FAILED" $O/r2c11_pytest_stem.log | head -10 | cut -c1-300
( time timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_parity2_gpu.py -m gpu -q -x ) > $O/r2c11_pytest_eng.log 2>&1
tail -3 $O/r2c11_pytest_eng.log; grep -E "^E  |FAILED" $O/r2c11_pytest_eng.log | head -10 | cut -c1-300
for bn in eval train; do
  timeout 300 python tools/profile_engine.py --batch 8 --reid 4096 --reid-bn $bn --out $O/r2c11_prof_r_${bn}.json > $O/r2c11_prof_r_${bn}.log 2>&1; grep "^reid" $O/r2c11_prof_r_${bn}.log; grep -E "stem|roi" $O/r2c11_prof_r_${bn}.log | head -4
  VCB_REID_STEM=patches timeout 300 python tools/profile_engine.py --batch 8 --reid 4096 --reid-bn $bn --out $O/r2c11_prof_r_${bn}_patches.json > $O/r2c11_prof_r_${bn}_patches.log 2>&1; grep "^reid" $O/r2c11_prof_r_${bn}_patches.log; grep -E "stem|roi" $O/r2c11_prof_r_${bn}_patches.log | head -4
done
