"""Per-op CUDA-event timing of the YOLO / ReID plans (development aid; run under gpurun).

    python tools/profile_engine.py --model yolov5m --batch 32 --size 640
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="yolov5m")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--reid", type=int, default=0, help="crops for the ReID plan (0 = skip)")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--mode", default="auto", choices=["auto", "gather"])
    ap.add_argument("--reid-bn", default="eval", choices=["eval", "train"])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "profile_engine.json"))
    a = ap.parse_args()
    from vehicle_counting_b200 import _lib as L
    from vehicle_counting_b200.engine import YoloEngine, ReidEngine
    from vehicle_counting_b200.weights import synth_yolov5_state_dict, synth_reid_state_dict
    a_mode = L.A_GATHER if a.mode == "gather" else L.A_AUTO
    dev = torch.device("cuda:0")
    res = {}
    sd = synth_yolov5_state_dict(a.model, seed=0)
    eng = YoloEngine(sd, a.batch, a.size, a.size, model_name=a.model, a_mode=a_mode)
    eng.frames.copy_(torch.randint(0, 256, tuple(eng.frames.shape), dtype=torch.uint8, device=dev))
    plan = eng.plan
    st = plan.stream

    def time_steps(plan, label):
        n = len(plan.steps)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(a.iters)]
        with torch.cuda.stream(plan.stream):
            for it in range(a.iters):
                ev[it][0].record(plan.stream)
                for i, fn in enumerate(plan.steps):
                    fn(plan.stream)
                    ev[it][i + 1].record(plan.stream)
        torch.cuda.synchronize()
        per = np.array([[ev[it][i].elapsed_time(ev[it][i + 1]) for i in range(n)] for it in range(1, a.iters)]).min(0)
        return per

    per = time_steps(plan, "yolo")
    # graph replay timing
    plan.run(True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(10):
            plan.run(True)
        e1.record(st)
    torch.cuda.synchronize()
    ms_graph = e0.elapsed_time(e1) / 10
    flops = plan.conv_flops
    print(f"{a.model} B={a.batch} {a.size}x{a.size}: eager sum {per.sum():.3f} ms, graph {ms_graph:.3f} ms/batch -> "
          f"{a.batch / ms_graph * 1e3:.0f} FPS, conv {flops / ms_graph / 1e9:.1f} TFLOP/s ({plan.num_convs} convs, "
          f"{eng.plan.graph.num_kernels} kernels)")
    order = np.argsort(-per)
    for i in order[:30]:
        fl = plan.step_flops[i]
        print(f"  step {i:3d}: {per[i] * 1e3:8.1f} us  {fl / per[i] / 1e9 if fl else 0:7.1f} TF/s  {plan.labels[i]}")
    res["yolo"] = {"model": a.model, "batch": a.batch, "size": a.size, "ms_graph": ms_graph, "eager_ms": per.tolist(),
                   "conv_flops": flops, "kernels": eng.plan.graph.num_kernels, "dets": eng.download()[1].tolist(),
                   "labels": plan.labels, "step_flops": plan.step_flops}
    if a.reid:
        rsd = synth_reid_state_dict(0)
        r = ReidEngine(rsd, capacity=a.reid, bn_mode=a.reid_bn, a_mode=a_mode, max_segments=max(a.reid // 64, 8))
        seg = [64] * (a.reid // 64) if a.reid % 64 == 0 else [a.reid]
        rng = np.random.default_rng(0)
        wh = rng.uniform(32, 256, (a.reid, 2)); tl = rng.uniform(0, 1, (a.reid, 2)) * (a.size - wh)
        rois = np.concatenate([np.zeros((a.reid, 1)), tl, tl + wh], 1).astype(np.int32)
        rois[:, 0] = np.arange(a.reid) % a.batch
        r.run(eng.frames, rois, seg_sizes=seg); torch.cuda.synchronize()
        key = next(iter(r._plans)); rp = r._plans[key]["plan"]
        per_r = time_steps(rp, "reid")
        with torch.cuda.stream(r.stream):
            e0.record(r.stream)
            for _ in range(10):
                rp.run(True)
            e1.record(r.stream)
        torch.cuda.synchronize()
        ms_r = e0.elapsed_time(e1) / 10
        print(f"reid n={a.reid}: eager sum {per_r.sum():.3f} ms, graph {ms_r:.3f} ms -> {rp.conv_flops / ms_r / 1e9:.1f} TFLOP/s")
        for i in np.argsort(-per_r)[:60 if a.reid_bn == "train" else 14]:
            fl = rp.step_flops[i]
            print(f"  step {i:3d}: {per_r[i] * 1e3:8.1f} us  {fl / per_r[i] / 1e9 if fl else 0:7.1f} TF/s  {rp.labels[i]}")
        res["reid"] = {"n": a.reid, "ms_graph": ms_r, "eager_ms": per_r.tolist(), "conv_flops": rp.conv_flops,
                       "labels": rp.labels, "step_flops": rp.step_flops}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"))


if __name__ == "__main__":
    main()
