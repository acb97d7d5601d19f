#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c31_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/c31_pytest.log | tail -2; grep -E "^E|FAILED" gpurun_out/c31_pytest.log | head -5 | cut -c1-300
for c in patch-big-reid-l1 auto-big-reid-l1 patch-big-3x3-48 auto-big-3x3-48 patch-big-yolos-64 auto-big-yolos-64 patch-big-reid-l2 auto-big-reid-l2 patch-big-3x3-96 auto-big-3x3-96 patch-big-3x3-192-res fast-big-3x3-192-res fast-big-1x1-96 auto-big-reid-l4; do
  timeout 300 python tests/bringup_conv.py --only $c --out gpurun_out/c31_one.jsonl > /dev/null 2>&1
  python - <<P
import json
for l in open("gpurun_out/c31_one.jsonl"):
    d=json.loads(l)
    if d.get("case") == "$c": print(d.get("case"), d.get("us"), d.get("tflops"), d.get("ok"), d.get("fault"), (d.get("stderr") or "")[-200:])
P
done
timeout 300 python bench.py --batch 64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c31_bench.json 2> gpurun_out/c31_bench.err
python - <<P
import json
d=json.load(open("gpurun_out/c31_bench.json"))
print("fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["achieved"]), "kernels", d["kernels_per_step"], d["clocks"])
P
