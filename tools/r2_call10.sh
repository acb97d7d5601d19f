#!/bin/bash
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -m gpu -q -x ) > $O/r2c10_pytest.log 2>&1
tail -3 $O/r2c10_pytest.log; grep -E "^E  |FAILED" $O/r2c10_pytest.log | head -10 | cut -c1-300
for b in 8 16 32; do
  timeout 300 python tools/profile_engine.py --batch $b --out $O/r2c10_prof_y_b$b.json > $O/r2c10_prof_y_b$b.log 2>&1; head -1 $O/r2c10_prof_y_b$b.log
done
for n in 512 1024 4096; do
  timeout 300 python tools/profile_engine.py --batch 8 --reid $n --reid-bn train --out $O/r2c10_prof_r_$n.json > $O/r2c10_prof_r_$n.log 2>&1; grep -A3 "^reid" $O/r2c10_prof_r_$n.log | head -4
done
VCB_EPI_STATS=0 timeout 300 python tools/profile_engine.py --batch 8 --reid 4096 --reid-bn train --out $O/r2c10_prof_r_4096_nostats.json > $O/r2c10_prof_r_4096_nostats.log 2>&1; grep -A3 "^reid" $O/r2c10_prof_r_4096_nostats.log | head -4
( time python bench.py ) > $O/r2c10_bench.json 2> $O/r2c10_bench.err; tail -3 $O/r2c10_bench.err
python - <<P
import json
d = json.loads(open("$O/r2c10_bench.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("N=1 fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pipe", round(d["e2e_pipelined"]["value"]), "b1", round(d["dropin_b1"]["value"]), "folded", round(d["folded_bn"]["value"]), "frac", round(r["frac"],3), "yolo_frac", round(r["yolo_frac"],3), "cpu", d["cpu_baseline"]["value"], d["clocks"])
P
