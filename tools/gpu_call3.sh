#!/bin/bash
# 256-row tiles: parity + timing
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c3_pytest.log 2>&1
tail -5 gpurun_out/c3_pytest.log
timeout 600 python tests/bringup_conv.py --only big- --skip sweep,persistent,c4- --out gpurun_out/c3_bringup_big.jsonl > gpurun_out/c3_bringup_big.log 2>&1
python - <<P
import json
for l in open("gpurun_out/c3_bringup_big.jsonl"):
    d=json.loads(l); print(d.get("case"), d.get("us"), d.get("tflops"), d.get("ok"), d.get("rel"))
P
for B in 64 128; do
  timeout 300 python bench.py --batch $B --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c3_bench_b$B.json 2> gpurun_out/c3_bench_b$B.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/c3_bench_b$B.json"))
    print("B=$B", "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["achieved"]), d["clocks"])
except Exception as e:
    print("B=$B FAILED", e); print(open("gpurun_out/c3_bench_b$B.err").read()[-1500:])
P
done
timeout 300 python tools/profile_engine.py --batch 64 --reid 4096 --out gpurun_out/c3_profile_b64.json > gpurun_out/c3_profile_b64.log 2>&1
head -30 gpurun_out/c3_profile_b64.log
