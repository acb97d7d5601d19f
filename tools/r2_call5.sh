#!/bin/bash
# round 2, GPU call 5: full suite incl. pipeline / v5 / scale tests, train-mode per-kernel profile, cuDNN comparator
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/r2c5_pytest.log 2>&1
tail -4 $O/r2c5_pytest.log; grep -E "^E  |FAILED" $O/r2c5_pytest.log | head -20 | cut -c1-300
cat $O/pipeline_csv_agreement.json 2>/dev/null; echo
timeout 400 python tools/profile_engine.py --batch 64 --reid 4096 --reid-bn train --out $O/r2c5_profile_train.json > $O/r2c5_profile_train.log 2>&1
grep -A60 "^reid" $O/r2c5_profile_train.log | head -64
timeout 900 python tools/cudnn_compare.py --out $O/r02_cudnn_compare.json > $O/r2c5_cudnn.log 2>&1
tail -70 $O/r2c5_cudnn.log
