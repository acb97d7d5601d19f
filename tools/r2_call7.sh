#!/bin/bash
# round 2, GPU call 7: fused train-mode stem
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity2_gpu.py tests/test_engine_gpu.py tests/test_pipeline_gpu.py -m gpu -q ) > $O/r2c7_pytest.log 2>&1
tail -3 $O/r2c7_pytest.log; grep -E "^E  |FAILED" $O/r2c7_pytest.log | head -10 | cut -c1-300
timeout 400 python tools/profile_engine.py --batch 64 --reid 4096 --reid-bn train --out $O/r2c7_profile_train.json > $O/r2c7_profile_train.log 2>&1
grep -A12 "^reid" $O/r2c7_profile_train.log
( timeout 600 python bench.py --no-cpu-baseline ) > $O/r2c7_bench.json 2> $O/r2c7_bench.err
python - <<P
import json
d = json.loads(open("$O/r2c7_bench.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pipe", round(d["e2e_pipelined"]["value"]), "folded", round(d["folded_bn"]["value"]), "reid_ms", round(r["reid_ms_per_step"],3), "other", round(r["other_kernels_ms_per_step"],3), "kernels", d["kernels_per_step"])
P
