#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reid_stem_direct -s 2 -c 2 -o $O/r2c13_ncu_stem_direct_m1 -f \
  python tools/profile_engine.py --batch 8 --reid 4096 --reid-bn train --iters 2 --out $O/r2c13_tmp.json > $O/r2c13_ncu.log 2>&1
tail -2 $O/r2c13_ncu.log
ls -la $O/r2c13_ncu_stem_direct_m1.ncu-rep
