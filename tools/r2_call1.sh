#!/bin/bash
# round 2, GPU call 1: parity of the new kernel paths, then A/B of the step-level switches (one process each)
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/r2c1_pytest.log 2>&1
tail -4 $O/r2c1_pytest.log; grep -E "^E|FAILED" $O/r2c1_pytest.log | head -8 | cut -c1-300
timeout 300 python tests/bringup_conv.py --only rowwin --out $O/r2c1_rowwin.jsonl 2>&1 | tail -7 | cut -c1-400
b() { tag=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > $O/r2c1_bench_$tag.json 2> $O/r2c1_bench_$tag.err; python - <<P
import json
try:
    d = json.load(open("$O/r2c1_bench_$tag.json"))
    print("$tag", "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"], 3), "conv_ms", round(d["roofline"]["conv_ms_per_step"],3), "other", round(d["roofline"]["other_kernels_ms_per_step"],3), d["clocks"]["sm_mhz"])
except Exception as e:
    print("$tag", "FAILED", e); print(open("$O/r2c1_bench_$tag.err").read()[-800:])
P
}
b default X=1
b plain VCB_LIB_PATH=$PWD/vehicle_counting_b200/libvcb200_plain.so
b norowwin VCB_STEM_ROWWIN=0
b rev VCB_TILE_REV=1
b l2hint VCB_L2_HINT=1
b rev_l2hint VCB_TILE_REV=1 VCB_L2_HINT=1
b pdl VCB_PDL=1
b default2 X=1
timeout 300 python tools/profile_engine.py --batch 64 --reid 4096 --out $O/r2c1_profile_engine.json > $O/r2c1_profile_engine.log 2>&1
head -3 $O/r2c1_profile_engine.log
VCB_TILE_REV=1 timeout 300 python tools/profile_engine.py --batch 64 --reid 4096 --out $O/r2c1_profile_engine_rev.json > $O/r2c1_profile_engine_rev.log 2>&1
head -3 $O/r2c1_profile_engine_rev.log
