#!/bin/bash
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_parity2_gpu.py -m gpu -q -x ) > $O/r2c17_pytest.log 2>&1
tail -3 $O/r2c17_pytest.log; grep -E "^E  |FAILED" $O/r2c17_pytest.log | head -10 | cut -c1-300
timeout 300 python tools/profile_engine.py --batch 64 --reid 4096 --reid-bn train --out $O/r2c17_prof_train.json > $O/r2c17_prof_train.log 2>&1; head -1 $O/r2c17_prof_train.log; grep "^reid" $O/r2c17_prof_train.log
python - <<P
import json
d=json.load(open("$O/r2c17_prof_train.json"))["reid"]
import collections
agg=collections.Counter()
for l,m in zip(d["labels"],d["eager_ms"]):
    k = "stem" if "stem" in l else "conv" if l.startswith("conv") else "bn stats" if "stats" in l else "bn apply" if "apply" in l else "finalize" if "finalize" in l else l
    agg[k]+=m
print({k:round(v,3) for k,v in agg.items()})
P
