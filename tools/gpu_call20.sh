#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c20_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/c20_pytest.log | tail -2
timeout 300 python bench.py --batch 64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err
python - <<P
import json
d=json.load(open("gpurun_out/c20_bench.json"))
print("fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["achieved"]), "kernels", d["kernels_per_step"], d["clocks"])
P
timeout 300 python tools/profile_engine.py --batch 64 --reid 4096 --out gpurun_out/c20_profile_b64.json > gpurun_out/c20_profile_b64.log 2>&1
head -14 gpurun_out/c20_profile_b64.log; grep -A5 "^reid" gpurun_out/c20_profile_b64.log
