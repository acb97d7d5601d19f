#!/usr/bin/env python
"""The library comparator on the same B200 (SURVEY 2b / BASELINE.md 3): the network the reference reaches through
torch.hub (YOLOv5 v6.0, restated in oracle/yolov5.py) run by PyTorch + cuDNN in fp16, channels-last, Conv+BN fused as upstream's
attempt_load(fuse=True) does, decode in torch, `torchvision.ops.nms` per image -- and the DeepSORT ReID net through torch in
fp16 (BatchNorm as shipped = batch statistics per 64-crop call, and eval) -- timed with CUDA events next to this repository's
engines in the same process.  Also reports how far the reference's own CUDA numerics (fp16 autocast) sit from the fp32 CPU
oracle on the head tensors: that is the yardstick for the head tolerances of tests/test_engine_gpu.py.

Test / measurement infrastructure (imports oracle/): not part of the product path.

    python tools/cudnn_compare.py --out gpurun_out/r02_cudnn_compare.json
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def fuse_model(model):
    from oracle import yolov5 as Y
    for m in model.modules():
        if isinstance(m, Y.Conv) and isinstance(m.bn, nn.BatchNorm2d):
            m.conv = torch.nn.utils.fusion.fuse_conv_bn_eval(m.conv, m.bn)
            m.bn = nn.Identity()
    return model


def backbone_heads(model, x):
    """DetectionModel.forward up to the raw Detect logits (the oracle's Detect builds its grids on the CPU)"""
    from oracle import yolov5 as Y
    y = []
    for i, m in enumerate(model.model):
        f = model.froms[i]
        xin = (x if f == -1 else y[f]) if isinstance(f, int) else [x if j == -1 else y[j] for j in f]
        if isinstance(m, nn.Identity):
            x = torch.cat(xin, 1)
        elif isinstance(m, Y.Detect):
            return [m.m[k](xin[k]) for k in range(m.nl)], m
        else:
            x = m(xin)
        y.append(x)


def decode(raw, det, grids):
    z = []
    for i, x in enumerate(raw):
        bs, _, ny, nx = x.shape
        x = x.view(bs, det.na, det.no, ny, nx).permute(0, 1, 3, 4, 2)
        yv = x.float().sigmoid()
        grid, ag, stride = grids[i]
        xy = (yv[..., 0:2] * 2.0 - 0.5 + grid) * stride
        wh = (yv[..., 2:4] * 2.0) ** 2 * ag
        z.append(torch.cat((xy, wh, yv[..., 4:]), -1).reshape(bs, -1, det.no))
    return torch.cat(z, 1)


def nms_upstream(pred, conf=0.25, iou=0.45, max_det=300, max_wh=4096.0, max_nms=30000):
    import torchvision
    out = []
    for x in pred:                                   # per image, as upstream non_max_suppression
        x = x[x[:, 4] > conf]
        if not x.shape[0]:
            out.append(x[:, :6]); continue
        x = x.clone()
        x[:, 5:] *= x[:, 4:5]
        box = torch.stack([x[:, 0] - x[:, 2] / 2, x[:, 1] - x[:, 3] / 2, x[:, 0] + x[:, 2] / 2, x[:, 1] + x[:, 3] / 2], 1)
        c, j = x[:, 5:].max(1, keepdim=True)
        x = torch.cat((box, c, j.float()), 1)[c.view(-1) > conf]
        if x.shape[0] > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        keep = torchvision.ops.nms(x[:, :4] + x[:, 5:6] * max_wh, x[:, 4], iou)[:max_det]
        out.append(x[keep])
    return out


def time_ms(fn, iters, stream=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="yolov5m")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--crops", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_cudnn_compare.json"))
    a = ap.parse_args()
    from oracle import reid as R
    from oracle import yolov5 as Y
    from vehicle_counting_b200.engine import ReidEngine, YoloEngine
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "workload": {"model": a.model, "batch": a.batch, "size": a.size, "crops": a.crops}}
    B, S = a.batch, a.size
    rng = np.random.default_rng(0)
    frames = torch.from_numpy(rng.integers(0, 256, (B, S, S, 3), dtype=np.uint8)).to(dev)

    # ---------------- detector: cuDNN fp16 channels-last, fused Conv+BN ----------------
    model = Y.build(a.model, seed=0, obj_bias=-3.0)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    lib_model = fuse_model(Y.build(a.model, seed=0, obj_bias=-3.0)).to(dev).half().to(memory_format=torch.channels_last).eval()
    det = lib_model.model[-1]
    grids = []
    for i, st in enumerate((8, 16, 32)):
        ny, nx = S // st, S // st
        yv, xv = torch.meshgrid(torch.arange(ny, device=dev), torch.arange(nx, device=dev), indexing="ij")
        grids.append((torch.stack((xv, yv), 2).view(1, 1, ny, nx, 2).float(),
                      (det.anchors[i].to(dev).float() * st).view(1, det.na, 1, 1, 2), float(st)))

    @torch.no_grad()
    def lib_step():
        x = (frames.permute(0, 3, 1, 2).half() / 255.0).contiguous(memory_format=torch.channels_last)
        raw, d = backbone_heads(lib_model, x)
        return nms_upstream(decode(raw, d, grids))

    @torch.no_grad()
    def lib_convs_only():
        x = (frames.permute(0, 3, 1, 2).half() / 255.0).contiguous(memory_format=torch.channels_last)
        backbone_heads(lib_model, x)

    ms_lib = time_ms(lib_step, a.iters)
    ms_lib_conv = time_ms(lib_convs_only, a.iters)
    eng = YoloEngine(sd, B, S, S, model_name=a.model)
    eng.frames.copy_(frames)
    eng.forward(); torch.cuda.synchronize()
    ms_ours = time_ms(lambda: eng.forward(), a.iters, eng.plan.stream)
    gflop = eng.plan.conv_flops / 1e9
    res["detector"] = {"conv_gflop_per_step": gflop,
                       "cudnn_fp16_channels_last": {"ms_per_step": ms_lib, "fps": B / ms_lib * 1e3, "ms_backbone_heads_only": ms_lib_conv,
                                                    "conv_tflops": gflop / ms_lib_conv},
                       "ours": {"ms_per_step": ms_ours, "fps": B / ms_ours * 1e3, "conv_tflops_whole_step": gflop / ms_ours},
                       "speedup_ours_over_cudnn": ms_lib / ms_ours}
    del eng
    torch.cuda.empty_cache()

    # ---------------- head deviation of the reference's own CUDA numerics (fp16 autocast) from fp32 ----------------
    small = Y.build("yolov5s", seed=0, obj_bias=-4.0)
    imgs = np.random.default_rng(0).integers(0, 256, (4, 640, 640, 3), dtype=np.uint8)
    x32 = torch.from_numpy(imgs).to(dev).permute(0, 3, 1, 2).float() / 255.0
    with torch.no_grad():
        small_gpu = small.to(dev).eval()
        raw32, _ = backbone_heads(small_gpu, x32)
        with torch.autocast("cuda", dtype=torch.float16):
            raw_ac, _ = backbone_heads(small_gpu, x32)
        fused_half = fuse_model(Y.build("yolov5s", seed=0, obj_bias=-4.0)).to(dev).half().eval()
        raw_h, _ = backbone_heads(fused_half, x32.half())
        cpu_model = Y.build("yolov5s", seed=0, obj_bias=-4.0)
        raw_cpu, _ = backbone_heads(cpu_model, x32.cpu())
    eng_s = YoloEngine(small.state_dict(), 4, 640, 640, model_name="yolov5s")
    eng_s.frames.copy_(torch.from_numpy(imgs).to(dev)); eng_s.forward(); torch.cuda.synchronize()
    dev_rows = []
    for li in range(3):
        ref = raw_cpu[li].float()
        ours = eng_s.logits[li].float().cpu()[..., :255].permute(0, 3, 1, 2)

        def rel(t):
            t = t.float().cpu()
            return float((t - ref).norm() / ref.norm()), float((t - ref).abs().max() / ref.abs().max())
        dev_rows.append({"head": li, "torch_fp32_gpu": rel(raw32[li]), "torch_fp16_autocast (reference CUDA path)": rel(raw_ac[li]),
                         "torch_fp16_fused_half": rel(raw_h[li]), "ours": rel(ours)})
    res["head_deviation_vs_fp32_cpu_oracle (rel L2, rel max) yolov5s 640 B=4"] = dev_rows
    del eng_s
    torch.cuda.empty_cache()

    # ---------------- ReID: torch fp16 (cuDNN) vs ours, both BatchNorm modes ----------------
    ckpt = os.path.join(ROOT, "oracle", "_ref", "reid_ckpt.npz")
    rsd = R.load_state_dict(ckpt) if os.path.isfile(ckpt) else R.seeded_state_dict(0)
    n = a.crops
    xin = torch.randn(n, 3, 50, 50, device=dev).half().contiguous(memory_format=torch.channels_last)
    sd_h = {k: (v.to(dev).half() if v.is_floating_point() else v.to(dev)) for k, v in rsd.items()}
    for k in list(sd_h):
        if sd_h[k].ndim == 4:
            sd_h[k] = sd_h[k].contiguous(memory_format=torch.channels_last)

    def reid_lib(mode):
        if mode == "eval":
            return R.net_forward(sd_h, xin, "eval")
        outs = [R.net_forward(sd_h, xin[i:i + 64], "train") for i in range(0, n, 64)]      # one reference call per 64 crops
        return torch.cat(outs, 0)

    reid_gflop = 0.844 * n
    res["reid"] = {"crops": n, "conv_gflop": reid_gflop}
    for mode in ("eval", "train"):
        ms_l = time_ms(lambda: reid_lib(mode), max(a.iters // 2, 3))
        r = ReidEngine(rsd, capacity=n, bn_mode=mode, max_segments=max(n // 64, 8))
        fr = torch.from_numpy(rng.integers(0, 256, (n // 64, S, S, 3), dtype=np.uint8)).to(dev)
        wh = rng.uniform(32, 256, (n, 2)); tl = rng.uniform(0, 1, (n, 2)) * (S - wh)
        rois = np.concatenate([np.repeat(np.arange(n // 64), 64)[:, None], tl, tl + wh], 1).astype(np.int32)
        seg = [64] * (n // 64)
        r.run(fr, rois, seg_sizes=seg); torch.cuda.synchronize()
        ms_o = time_ms(lambda: r.run(fr, None, n=n, seg_sizes=seg), a.iters, r.stream)
        res["reid"][mode] = {"torch_fp16_ms": ms_l, "torch_fp16_tflops": reid_gflop / ms_l, "ours_ms (incl. crop+resize)": ms_o,
                             "ours_tflops": reid_gflop / ms_o, "speedup_ours_over_torch": ms_l / ms_o}
        del r
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
