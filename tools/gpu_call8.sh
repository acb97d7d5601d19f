#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c8_pytest.log 2>&1
tail -4 gpurun_out/c8_pytest.log | cut -c1-300
VCB_PROF=0 timeout 900 python tests/bringup_conv.py --only big- --skip sweep,persistent,c4-,tanh,stem,fast-big-1x1-96,m256,xp- --out gpurun_out/c8_big.jsonl > gpurun_out/c8_big.log 2>&1
python - <<P
import json
for l in open("gpurun_out/c8_big.jsonl"):
    d=json.loads(l); print(d.get("case"), d.get("us"), d.get("tflops"), d.get("ok"), d.get("fault"), (d.get("stderr") or "")[-300:])
P
