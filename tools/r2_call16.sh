#!/bin/bash
O=gpurun_out; mkdir -p $O
( time timeout 300 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -x -k pinned ) > $O/r2c16_pytest.log 2>&1
tail -3 $O/r2c16_pytest.log; grep -E "^E  |FAILED" $O/r2c16_pytest.log | head -10 | cut -c1-300
( time timeout 1200 python tests/bringup_conv.py --only bnsweep --out $O/r2c16_bnsweep.jsonl ) > $O/r2c16_bnsweep.log 2>&1
tail -3 $O/r2c16_bnsweep.log
python - <<P
import json
for l in open("$O/r2c16_bnsweep.jsonl"):
    d=json.loads(l); print(f"{d['case']:42s} {d['us']:8.1f} us {d['tflops']:8.1f} TF rel {d['rel']:.1e} fault {d['fault'][0]}")
P
