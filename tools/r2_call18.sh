#!/bin/bash
O=gpurun_out; mkdir -p $O
for tag in park nopark; do
  if [ $tag = nopark ]; then export VCB_LIB_PATH=$PWD/vehicle_counting_b200/libvcb200_nopark.so; else unset VCB_LIB_PATH; fi
  for rep in 1 2; do
    timeout 300 python tools/profile_engine.py --batch 64 --reid 4096 --reid-bn eval --out $O/r2c18_prof_${tag}_$rep.json > $O/r2c18_prof_${tag}_$rep.log 2>&1
    echo $tag $rep: $(head -1 $O/r2c18_prof_${tag}_$rep.log | cut -c1-120) '|' $(grep "^reid" $O/r2c18_prof_${tag}_$rep.log)
  done
done
unset VCB_LIB_PATH
( time timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q -x ) > $O/r2c18_pytest.log 2>&1
tail -3 $O/r2c18_pytest.log | head -1
