#!/bin/bash
# round 2, GPU call 3: split epilogue parity + A/B
O=gpurun_out; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2c3_pytest.log 2>&1
tail -4 $O/r2c3_pytest.log; grep -E "^E|FAILED" $O/r2c3_pytest.log | head -8 | cut -c1-300
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > $O/r2c3_bench_$tag.json 2> $O/r2c3_bench_$tag.err; python - <<P
import json
try:
    d = json.load(open("$O/r2c3_bench_$tag.json"))
    print("$tag", "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"], 3), "conv_ms", round(d["roofline"]["conv_ms_per_step"],3), "other", round(d["roofline"]["other_kernels_ms_per_step"],3), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag", "FAILED", e); print(open("$O/r2c3_bench_$tag.err").read()[-800:])
P
}
b split X=1
b nosplit VCB_EPI_SPLIT=0
b split_rev VCB_TILE_REV=1
b split_norowwin VCB_STEM_ROWWIN=0
b split_b128 X=1 
for c in fast-big-stem rowwin-big-stem m128-big-1x1-192 fast-big-1x1-96 auto-big-3x3-96 m128-big-3x3-192-res pair2-big-3x3-192-res auto-big-reid-l4; do
  idx=$(python - <<P
import sys; sys.path.insert(0,'tests')
import bringup_conv as B
print([i for i,(n,_) in enumerate(B.CASES) if n=="$c"][0])
P
)
  for sp in 1 0; do
    VCB_EPI_SPLIT=$sp timeout 120 python tests/bringup_conv.py --case $idx 2>&1 | grep RESULT | python -c "
import sys, json
for l in sys.stdin:
    r=json.loads(l[7:]); print('split=$sp', r['case'], r['us'], 'us', r['tflops'], 'TF rel', round(r['rel'],5))
"
  done
done
timeout 300 python tools/profile_engine.py --batch 64 --reid 4096 --out $O/r2c3_profile_engine.json > $O/r2c3_profile_engine.log 2>&1
head -3 $O/r2c3_profile_engine.log; grep "^reid" $O/r2c3_profile_engine.log
