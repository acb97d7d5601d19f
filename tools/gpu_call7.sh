#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c7_pytest.log 2>&1
tail -4 gpurun_out/c7_pytest.log
VCB_PROF=0 timeout 900 python tests/bringup_conv.py --only xp- --out gpurun_out/c7_xp.jsonl > gpurun_out/c7_xp.log 2>&1
python - <<P
import json
for l in open("gpurun_out/c7_xp.jsonl"):
    d=json.loads(l); print(d.get("case"), d.get("us"), d.get("tflops"), d.get("ok"), d.get("fault"), (d.get("stderr") or "")[-300:])
P
for B in 64; do
  timeout 300 python bench.py --batch $B --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c7_bench_b$B.json 2> gpurun_out/c7_bench_b$B.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/c7_bench_b$B.json"))
    print("B=$B", "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["achieved"]), d["clocks"])
except Exception as e:
    print("B=$B FAILED", e); print(open("gpurun_out/c7_bench_b$B.err").read()[-1500:])
P
done
