#!/bin/bash
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/r2c14_pytest.log 2>&1
tail -3 $O/r2c14_pytest.log; grep -E "^E  |FAILED" $O/r2c14_pytest.log | head -10 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py ) > $O/r2c14_bench.json 2> $O/r2c14_bench.err; tail -3 $O/r2c14_bench.err
python - <<P
import json
d = json.loads(open("$O/r2c14_bench.json").read().strip().splitlines()[-1]); r = d["roofline"]
print("N=1 fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pipe", round(d["e2e_pipelined"]["value"]), "b1", round(d["dropin_b1"]["value"]), "folded", round(d["folded_bn"]["value"]), "frac", round(r["frac"],3), "yolo_frac", round(r["yolo_frac"],3), "cpu", d["cpu_baseline"]["value"], d["clocks"])
P
timeout 300 ncu --set full --clock-control none -k regex:reid_stem_direct -s 2 -c 1 -o $O/r2c14_ncu_stem_direct_eval -f python tools/profile_engine.py --batch 8 --reid 4096 --reid-bn eval --iters 2 --out $O/r2c14_tmp.json > $O/r2c14_ncu.log 2>&1
