#!/bin/bash
# role timers on the big layer shapes (128-row two-CTA mode vs 256-row mode)
mkdir -p gpurun_out
VCB_PROF=1 timeout 900 python tests/bringup_conv.py --only big- --skip sweep,persistent,c4-,tanh --out gpurun_out/c4_prof_big.jsonl > gpurun_out/c4_prof_big.log 2>&1
python - <<P
import json
for l in open("gpurun_out/c4_prof_big.jsonl"):
    d=json.loads(l); print(d.get("case"), d.get("us"), d.get("tflops"), d.get("ok")); print("    ", d.get("prof"))
P
