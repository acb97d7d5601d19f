#!/bin/bash
O=gpurun_out; mkdir -p $O
( time timeout 900 python -m pytest tests/test_adapter_gpu.py tests/test_parity2_gpu.py tests/test_pipeline_gpu.py tests/test_kernels_gpu.py -m gpu -q ) > $O/r2c8_pytest.log 2>&1
tail -3 $O/r2c8_pytest.log; grep -E "^E  |FAILED" $O/r2c8_pytest.log | head -10 | cut -c1-300
b() { tag=$1; shift; env "$@" timeout 600 python bench.py --no-cpu-baseline --quick > $O/r2c8_bench_$tag.json 2> $O/r2c8_bench_$tag.err; python - <<P
import json
try:
    d = json.loads(open("$O/r2c8_bench_$tag.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print("$tag", "fps", round(d["value"]), "e2e", round(d["e2e"]["value"]), "pipe", round(d["e2e_pipelined"]["value"]), "yolo_ms", round(r["yolo_conv_ms_per_step"],3), "yolo_frac", round(r["yolo_frac"],3), "other", round(r["other_kernels_ms_per_step"],3))
except Exception as e:
    print("$tag FAILED", e); print(open("$O/r2c8_bench_$tag.err").read()[-1000:])
P
}
b default X=1
b fp16logits VCB_FP16_LOGITS=1
