#!/bin/bash
# round 2, GPU call 4: full GPU suite (new parity cases, train-mode plan), new bench (train / eval), per-layer profile
O=gpurun_out; mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/r2c4_pytest.log 2>&1
tail -4 $O/r2c4_pytest.log; grep -E "^E  |FAILED" $O/r2c4_pytest.log | head -20 | cut -c1-260
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
b() { tag=$1; shift; ( time timeout 600 python bench.py "$@" ) > $O/r2c4_bench_$tag.json 2> $O/r2c4_bench_$tag.err; python - <<P
import json
try:
    d = json.loads(open("$O/r2c4_bench_$tag.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$tag", "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pipe", round(d.get("e2e_pipelined",{}).get("value",0)), "b1", round(d.get("dropin_b1",{}).get("value",0)),
          "folded", round(d.get("folded_bn",{}).get("value",0)), "conv TF", round(r["achieved"]), "frac", round(r["frac"],3), "yolo_frac", round(r["yolo_frac"],3),
          "reid_ms", round(r.get("reid_ms_per_step",0),3), "other", round(r["other_kernels_ms_per_step"],3), "cpu", d.get("cpu_baseline",{}).get("value"))
except Exception as e:
    print("$tag", "FAILED", e); print(open("$O/r2c4_bench_$tag.err").read()[-1500:])
P
}
b train
b eval --reid-bn eval --no-cpu-baseline
VCB_TILE_REV=1 b train_rev --no-cpu-baseline --quick
grep real $O/r2c4_bench_train.err
