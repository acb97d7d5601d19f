#!/bin/bash
# fast epilogue / tanh SiLU / PDL: parity + A/B timing
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/c2_pytest.log 2>&1
tail -5 gpurun_out/c2_pytest.log
timeout 600 python tests/bringup_conv.py --only big --out gpurun_out/c2_bringup_big.jsonl > gpurun_out/c2_bringup_big.log 2>&1
grep -E "PASS|FAIL" gpurun_out/c2_bringup_big.log | cut -c1-260
for v in "exp 0" "tanh 0" "exp 1" "tanh 1"; do
  set -- $v
  VCB_SILU=$1 VCB_PDL=$2 timeout 300 python bench.py --batch 64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench_$1_pdl$2.json 2> gpurun_out/c2_bench_$1_pdl$2.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/c2_bench_$1_pdl$2.json"))
    print("$1 pdl=$2", "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "conv TF", round(d["roofline"]["achieved"]), d["clocks"])
except Exception as e:
    print("$1 pdl=$2 FAILED", e); print(open("gpurun_out/c2_bench_$1_pdl$2.err").read()[-1500:])
P
done
timeout 300 python tools/profile_engine.py --batch 64 --reid 4096 --out gpurun_out/c2_profile_b64.json > gpurun_out/c2_profile_b64.log 2>&1
head -40 gpurun_out/c2_profile_b64.log
