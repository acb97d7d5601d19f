#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c19_pytest.log 2>&1
tail -4 gpurun_out/c19_pytest.log | cut -c1-300
for v in 1 0 1 0; do
  VCB_C3_FUSE=$v timeout 300 python bench.py --batch 64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c19_bench_fuse$v.json 2> gpurun_out/c19_bench_fuse$v.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/c19_bench_fuse$v.json"))
    print("c3fuse=$v", "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["achieved"]), "kernels", d["kernels_per_step"], d["clocks"])
except Exception as e:
    print("c3fuse=$v FAILED", e); print(open("gpurun_out/c19_bench_fuse$v.err").read()[-1500:])
P
done
