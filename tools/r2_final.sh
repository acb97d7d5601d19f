#!/bin/bash
# round 2, final records (run under gpurun): the default bench + reference arm, the other BASELINE configurations, the tracked
# pipeline, the ncu launch list of the timed steps and ncu --set full of the HBM-bound kernels.  Outputs in gpurun_out/ with the
# names they carry under profiles/.
O=gpurun_out; mkdir -p $O; rm -f $O/r02_pipeline_tracked.jsonl $O/r02_ncu_hbm_kernels.jsonl
( time timeout 1500 python -m pytest tests -m gpu -q ) > $O/r02_final_pytest.log 2>&1
tail -3 $O/r02_final_pytest.log | head -1; grep -E "^E  |FAILED" $O/r02_final_pytest.log | head -10 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
run() { out=$1; shift; ( time timeout 900 python bench.py "$@" ) > $O/$out.json 2> $O/$out.err; python - <<P
import json
try:
    d = json.loads(open("$O/$out.json").read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print("$out", d["config"]["workload"][:60], "| fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "pipe", round(d.get("e2e_pipelined",{}).get("value",0)), "b1", round(d.get("dropin_b1",{}).get("value",0)), "folded", round(d.get("folded_bn",{}).get("value",0)),
          "conv TF", round(r.get("achieved",0)), "frac", round(r.get("frac",0),3), "yolo_frac", round(r.get("yolo_frac",0),3), "yolo TF", round(r.get("yolo_tflops",0)), "cpu", d.get("cpu_baseline",{}).get("value"))
except Exception as e:
    print("$out", "FAILED", e); print(open("$O/$out.err").read()[-1200:])
P
}
run r02_bench_reference --impl reference --steps 5 --warmup 3
run r02_bench_ours
run r02_bench_ours_eval --reid-bn eval --no-cpu-baseline
run r02_bench_ours_b128 --batch 128 --no-cpu-baseline --quick
for b in 1 8 32; do run r02_bench_config1_b$b --config 1 --batch $b --steps 50 --no-cpu-baseline; done
run r02_bench_config2_train --config 2 --no-cpu-baseline
run r02_bench_config2_eval --config 2 --reid-bn eval --no-cpu-baseline --quick
run r02_bench_config4_384x640 --config 4 --no-cpu-baseline
run r02_bench_config4_736x1280 --config 4b --no-cpu-baseline
for bn in train eval; do for sz in 640 1280; do
  timeout 600 python tools/pipeline_demo.py --model yolov5l --frames 64 --batch 32 --bn $bn --size $sz 2>/dev/null | tail -1 >> $O/r02_pipeline_tracked.jsonl
done; done
cat $O/r02_pipeline_tracked.jsonl
VCB_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -c 800 --csv --log-file $O/r02_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick > $O/r02_ncu_bench.log 2>&1
wc -l $O/r02_ncu_launches.csv
python tools/summarize_launches.py $O/r02_ncu_launches.csv $O/r02_launch_summary.json 2 '["yolov5m", 640, 640, 64, 64, "train"]' | head -1
for k in detect_decode nms_kernel bn_seg_stats bn_seg_apply frames_to_f16 upsample2x sppf_pool avgpool_l2norm reid_stem_direct; do
  VCB_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:$k -c 1 -o $O/r02_ncu_$k -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --quick > $O/r02_ncu_$k.log 2>&1
  python tools/ncu_kernel_summary.py $O/r02_ncu_$k.ncu-rep >> $O/r02_ncu_hbm_kernels.jsonl 2>/dev/null
done
cat $O/r02_ncu_hbm_kernels.jsonl | cut -c1-400
