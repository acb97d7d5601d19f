#!/bin/bash
O=gpurun_out; mkdir -p $O
for mode in exp tanh; do
  rm -f $O/parity_yolo.jsonl
  VCB_SILU=$mode timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_parity2_gpu.py tests/test_fullsize_gpu.py -m gpu -q -k "yolo or heads" > $O/r2c19_pytest_$mode.log 2>&1
  tail -2 $O/r2c19_pytest_$mode.log | head -1; grep -E "^E  .*assert|FAILED" $O/r2c19_pytest_$mode.log | head -6 | cut -c1-200
  cp $O/parity_yolo.jsonl $O/r2c19_parity_$mode.jsonl 2>/dev/null
  VCB_SILU=$mode timeout 300 python tools/profile_engine.py --batch 64 --out $O/r2c19_prof_$mode.json > $O/r2c19_prof_$mode.log 2>&1; head -1 $O/r2c19_prof_$mode.log
done
python - <<P
import json
for mode in ("exp","tanh"):
    try:
        for l in open("$O/r2c19_parity_%s.jsonl" % mode):
            d=json.loads(l); print(mode, {k:(round(v,5) if isinstance(v,float) else v) for k,v in d.items() if k in ("name","hw","head_rel_l2_fp32","head_rel_l2_twin","rel_l2_fp32","rel_l2_twin","case","model")} , str(d)[:300])
    except Exception as e: print(mode, e)
P
