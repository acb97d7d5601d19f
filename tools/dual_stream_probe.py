"""Experiment: does running the detector as TWO independent half-batches on two streams (kernel tails / launch gaps of one overlap
the other's compute) beat one full batch?  python tools/dual_stream_probe.py --batch 64"""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="yolov5m")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--parts", type=int, default=2)
    a = ap.parse_args()
    from vehicle_counting_b200.engine import YoloEngine
    from vehicle_counting_b200.weights import synth_yolov5_state_dict
    sd = synth_yolov5_state_dict(a.model, seed=0, obj_bias=-3.0)
    dev = torch.device("cuda:0")
    res = {}

    def timed(engs, iters=20):
        for e in engs:
            e.frames.copy_(torch.randint(0, 256, tuple(e.frames.shape), dtype=torch.uint8, device=dev))
            e.forward()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main_s = engs[0].plan.stream
        ev = torch.cuda.Event()
        e0.record(main_s)
        for _ in range(iters):
            ev.record(main_s)
            for e in engs[1:]:
                e.plan.stream.wait_event(ev)
            for e in engs:
                e.forward()
            for e in engs[1:]:
                main_s.wait_stream(e.plan.stream)
        e1.record(main_s)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    one = YoloEngine(sd, a.batch, a.size, a.size, model_name=a.model)
    res["one_engine_ms"] = timed([one])
    del one
    torch.cuda.empty_cache()
    parts = [YoloEngine(sd, a.batch // a.parts, a.size, a.size, model_name=a.model) for _ in range(a.parts)]
    res[f"{a.parts}_engines_ms"] = timed(parts)
    res["fps_one"] = a.batch / res["one_engine_ms"] * 1e3
    res["fps_parts"] = a.batch / res[f"{a.parts}_engines_ms"] * 1e3
    print(json.dumps(res))


if __name__ == "__main__":
    main()
