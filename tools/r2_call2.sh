#!/bin/bash
# round 2, GPU call 2: role timers + ncu source-level captures of the latency-bound small-K layers
O=gpurun_out; mkdir -p $O
for c in fast-big-stem rowwin-big-stem noepi-big-stem m128-big-1x1-192 noepi-big-1x1-192 fast-big-1x1-96 noepi-big-1x1-96 auto-big-3x3-96 auto-big-3x3-48 m128-big-3x3-192-res; do
  idx=$(python - <<P
import sys; sys.path.insert(0,'tests')
import bringup_conv as B
print([i for i,(n,_) in enumerate(B.CASES) if n=="$c"][0])
P
)
  VCB_PROF=1 timeout 120 python tests/bringup_conv.py --case $idx 2>&1 | grep RESULT | cut -c1-900 >> $O/r2c2_prof.jsonl
done
cat $O/r2c2_prof.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r=json.loads(l[7:]); p=r.get('prof',{})
    print(r['case'], r['us'], r['tflops'], {k:v for k,v in p.items() if 'frac' in k or 'tiles' in k})
"
for c in m128-big-1x1-192 fast-big-stem fast-big-1x1-96; do
  idx=$(python - <<P
import sys; sys.path.insert(0,'tests')
import bringup_conv as B
print([i for i,(n,_) in enumerate(B.CASES) if n=="$c"][0])
P
)
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 2 -c 1 -o $O/r2c2_ncu_$c -f python tests/bringup_conv.py --case $idx > $O/r2c2_ncu_$c.log 2>&1
  tail -2 $O/r2c2_ncu_$c.log | cut -c1-200
done
ls -la $O/*.ncu-rep | tail -5
