for lib in "" $PWD/vehicle_counting_b200/libvcb200_nr.so; do
echo "LIB=$lib"
for c in m128-big-1x1-192 fast-big-1x1-96 actprobe-big-1x1-384-silu fast-big-stem auto-big-3x3-48 auto-big-3x3-96 m128-big-3x3-192-res; do
  idx=$(python - <<P
import sys; sys.path.insert(0,'tests')
import bringup_conv as B
print([i for i,(n,_) in enumerate(B.CASES) if n=="$c"][0])
P
)
  VCB_LIB_PATH=$lib python tests/bringup_conv.py --case $idx 2>&1 | grep RESULT | python -c "
import sys, json
for l in sys.stdin:
    r=json.loads(l[7:]); print(r['case'], r['us'], 'us', r['tflops'], 'TF rel', round(r['rel'],5))
"
done
VCB_LIB_PATH=$lib python bench.py --no-cpu-baseline --quick --reid-bn eval 2>/dev/null | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('bench eval fps', round(d['value']), 'yolo_ms', round(r['yolo_conv_ms_per_step'],3), 'yolo_frac', round(r['yolo_frac'],3))"
done
