#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tests/bringup_conv.py --only patch- --skip big --out gpurun_out/c26_patch.jsonl > gpurun_out/c26_patch.log 2>&1
python - <<P
import json
for l in open("gpurun_out/c26_patch.jsonl"):
    d=json.loads(l); print(d.get("case"), d.get("us"), d.get("ok"), d.get("rel"), d.get("fault"), d.get("bad_rows"), (d.get("stderr") or "")[-300:])
P
for c in patch-big-reid-l1 patchk1-big-reid-l1 auto-big-reid-l1 patch-big-3x3-48 patchk1-big-3x3-48 auto-big-3x3-48 patch-big-yolos-64 auto-big-yolos-64 patch-big-reid-l2 auto-big-reid-l2 patch-big-3x3-96 auto-big-3x3-96 patch-big-3x3-192-res fast-big-3x3-192-res; do
  timeout 300 python tests/bringup_conv.py --only $c --out gpurun_out/c26_one.jsonl > /dev/null 2>&1
  python - <<P
import json
for l in open("gpurun_out/c26_one.jsonl"):
    d=json.loads(l)
    if d.get("case") == "$c": print(d.get("case"), d.get("us"), d.get("tflops"), d.get("ok"), d.get("fault"), (d.get("stderr") or "")[-200:])
P
done
