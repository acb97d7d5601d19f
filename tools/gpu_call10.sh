#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c10_pytest.log 2>&1
tail -30 gpurun_out/c10_pytest.log | cut -c1-300
for v in 1 0; do
  VCB_REID_FUSED_STEM=$v timeout 300 python bench.py --batch 64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench_fused$v.json 2> gpurun_out/c10_bench_fused$v.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/c10_bench_fused$v.json"))
    print("fused=$v", "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["achieved"]), d["clocks"])
except Exception as e:
    print("fused=$v FAILED", e); print(open("gpurun_out/c10_bench_fused$v.err").read()[-1500:])
P
done
timeout 300 python tools/profile_engine.py --batch 64 --reid 4096 --out gpurun_out/c10_profile_b64.json > gpurun_out/c10_profile_b64.log 2>&1
grep -A8 "^reid" gpurun_out/c10_profile_b64.log
