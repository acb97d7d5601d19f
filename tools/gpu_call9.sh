#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c9_pytest.log 2>&1
tail -4 gpurun_out/c9_pytest.log | cut -c1-300
for B in 64 128; do
  timeout 300 python bench.py --batch $B --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench_b$B.json 2> gpurun_out/c9_bench_b$B.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/c9_bench_b$B.json"))
    print("B=$B", "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["achieved"]), d["clocks"])
except Exception as e:
    print("B=$B FAILED", e); print(open("gpurun_out/c9_bench_b$B.err").read()[-1500:])
P
done
timeout 300 python tools/profile_engine.py --batch 64 --reid 4096 --out gpurun_out/c9_profile_b64.json > gpurun_out/c9_profile_b64.log 2>&1
head -3 gpurun_out/c9_profile_b64.log; grep "^reid" gpurun_out/c9_profile_b64.log
