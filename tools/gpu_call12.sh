#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tests/bringup_conv.py --only patch- --skip big --out gpurun_out/c14_patch.jsonl > gpurun_out/c14_patch.log 2>&1
python - <<P
import json
for l in open("gpurun_out/c14_patch.jsonl"):
    d=json.loads(l); print(d.get("case"), d.get("us"), d.get("ok"), d.get("rel"), d.get("fault"), d.get("bad_rows"), d.get("n_rows"), d.get("bad_cols"), (d.get("first_bad_rows") or [])[:8], (d.get("stderr") or "")[-300:])
P
VCB_PROF=0 timeout 900 python tests/bringup_conv.py --only big --skip sweep,persistent,c4-,tanh,stem,fast-big-1x1,m256,xp-,1x1,s2,l3,l4,m128,pair2,192,384,l2,3x3-96 --out gpurun_out/c14_big.jsonl > gpurun_out/c14_big.log 2>&1
python - <<P
import json
for l in open("gpurun_out/c14_big.jsonl"):
    d=json.loads(l); print(d.get("case"), d.get("us"), d.get("tflops"), d.get("ok"), d.get("fault"), (d.get("stderr") or "")[-300:])
P
