#!/bin/bash
O=gpurun_out; mkdir -p $O
( time timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_adapter_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x ) > $O/r2c15_pytest.log 2>&1
tail -3 $O/r2c15_pytest.log; grep -E "^E  |FAILED" $O/r2c15_pytest.log | head -10 | cut -c1-300
python tools/e2e_breakdown2.py 2>&1 | tail -17
( time python bench.py ) > $O/r2c15_bench.json 2> $O/r2c15_bench.err; tail -3 $O/r2c15_bench.err
python - <<P
import json
d = json.loads(open("$O/r2c15_bench.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("N=1 fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pipe", round(d["e2e_pipelined"]["value"]), "b1", round(d["dropin_b1"]["value"]), "folded", round(d["folded_bn"]["value"]), "frac", round(r["frac"],3), "yolo_frac", round(r["yolo_frac"],3), "cpu", d["cpu_baseline"]["value"], d["clocks"])
P
