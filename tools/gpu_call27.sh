#!/bin/bash
mkdir -p gpurun_out
for c in patch-big-reid-l1 patch-big-3x3-48 patch-big-3x3-192-res; do
  VCB_PROF=1 timeout 300 python tests/bringup_conv.py --only $c --out gpurun_out/c27_one.jsonl > /dev/null 2>&1
  python - <<P
import json
for l in open("gpurun_out/c27_one.jsonl"):
    d=json.loads(l)
    if d.get("case") == "$c": print(d.get("case"), d.get("us"), d.get("tflops"), d.get("prof"), (d.get("stderr") or "")[-200:])
P
done
