for c in m128-big-1x1-192 actprobe-big-1x1-192-none actprobe-big-1x1-192-relu actprobe-big-1x1-192-tanh fast-big-1x1-96 fast-big-1x1-96-tanh actprobe-big-1x1-96-none actprobe-big-1x1-384-silu actprobe-big-1x1-384-none; do
  idx=$(python - <<P
import sys; sys.path.insert(0,'tests')
import bringup_conv as B
print([i for i,(n,_) in enumerate(B.CASES) if n=="$c"][0])
P
)
  python tests/bringup_conv.py --case $idx 2>&1 | grep RESULT | python -c "
import sys, json
for l in sys.stdin:
    r=json.loads(l[7:]); print(r['case'], r['us'], 'us', r['tflops'], 'TF rel', round(r['rel'],5))
"
done
