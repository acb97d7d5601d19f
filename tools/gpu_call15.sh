#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tests/bringup_conv.py --only xq- --out gpurun_out/c15_xq.jsonl > gpurun_out/c15_xq.log 2>&1
python - <<P
import json
for l in open("gpurun_out/c15_xq.jsonl"):
    d=json.loads(l); print(d.get("case"), d.get("us"), d.get("tflops"), d.get("fault"), (d.get("stderr") or "")[-300:])
P
