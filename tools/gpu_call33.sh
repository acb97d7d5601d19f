#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c33_pytest.log 2>&1
grep -E "passed|failed" gpurun_out/c33_pytest.log | tail -2; grep -E "^E|FAILED" gpurun_out/c33_pytest.log | head -5 | cut -c1-300
for c in auto-big-3x3-96 auto-big-reid-l2 fast-big-1x1-96 auto-big-1x1-192 m128-big-reid-l1 m128-big-3x3-192-res; do
  timeout 300 python tests/bringup_conv.py --only $c --out gpurun_out/c33_one.jsonl > /dev/null 2>&1
  python - <<P
import json
for l in open("gpurun_out/c33_one.jsonl"):
    d=json.loads(l)
    if d.get("case") == "$c": print(d.get("case"), d.get("us"), d.get("tflops"), d.get("ok"), d.get("fault"), (d.get("stderr") or "")[-200:])
P
done
for i in 1 2; do
timeout 300 python bench.py --batch 64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c33_bench$i.json 2> gpurun_out/c33_bench$i.err
python - <<P
import json
d=json.load(open("gpurun_out/c33_bench$i.json"))
print("fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["achieved"]), "kernels", d["kernels_per_step"], d["clocks"])
P
done
