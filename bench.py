#!/usr/bin/env python
"""Benchmark of the per-frame hot path: YOLOv5 detect + decode/NMS + ROI crop/resize + ReID embedding.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one JSON line on rank 0)
    python bench.py --impl reference --gpus N ...             # the CPU reference path (oracle port) on host cores
    torchrun --nnodes=1 --nproc-per-node N bench.py --gpus N  # one rank per GPU, frames sharded, weak scaling
    python bench.py --config {1,2,4}                          # the other BASELINE.json configurations (records under profiles/)

Default workload (BASELINE.json: "end-to-end FPS (detect+ReID+NMS) 640x640", configs[3] per-GPU shard): per step and per GPU a
batch of B synthetic random-uint8 640x640 frames goes through YOLOv5m (seeded synthetic weights) + decode + NMS, and 64 synthetic
ROIs per frame (w,h~U(32,256)) go through crop/resize/normalise + the ReID CNN with the reference's own
BatchNorm behaviour (batch statistics per call -- its Extractor never calls .eval(); one 64-crop segment per frame).  A "step" =
one such batch.  The JSON line carries, next to the contract keys:
  value          whole-job frames/s, inputs resident in HBM (CUDA events on the plan stream, graph replay)
  e2e            frames/s through the reference-shaped Python surface: ImageDetect.run(batch of numpy frames) -> dict of lists,
                 then Extractor.from_frames(BGR numpy frames, float64 boxes) -> numpy embeddings; every H2D / D2H copy, the
                 host-side gathering of the frame lists and the result conversion are inside the timed region
  e2e_pipelined  the same work through the double-buffered FramePipeline (pinned tensors in / out; uploads overlap compute)
  dropin_b1      the reference's own granularity: one frame per call (ImageDetect.run on a 1-frame batch, then one ReID call
                 for that frame's boxes), i.e. what modules/__init__.py:54-70 drives
  folded_bn      the device-resident number with BatchNorm folded (eval statistics) instead of the reference's batch statistics
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "end_to_end_fps_detect_reid_nms_640x640"
UNIT = "frames/s"


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "tflops_burst": p["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_sustained": 1400.0, "tflops_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def synth_boxes(rng, frames: int, per_frame: int, h: int, w: int) -> np.ndarray:
    """w,h ~ U(32,256), box inside the frame; float64 xyxy [frames, per_frame, 4]."""
    wh = rng.uniform(32, min(256, h - 2, w - 2), (frames, per_frame, 2))
    tl = rng.uniform(0, 1, (frames, per_frame, 2)) * (np.array([w, h]) - wh)
    return np.concatenate([tl, tl + wh], 2)


def boxes_to_rois(boxes: np.ndarray, h: int, w: int) -> np.ndarray:
    """reference crop rule (deep_sort.py:78-95): centre/size in float64, int() truncation, clip -> int32 [n, 5] (frame, x1, y1, x2, y2)"""
    F, P = boxes.shape[:2]
    b = boxes.reshape(-1, 4)
    bw, bh = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    cx, cy = b[:, 0] + bw / 2, b[:, 1] + bh / 2
    x1 = np.maximum(np.trunc(cx - bw / 2), 0); x2 = np.minimum(np.trunc(cx + bw / 2), w - 1)
    y1 = np.maximum(np.trunc(cy - bh / 2), 0); y2 = np.minimum(np.trunc(cy + bh / 2), h - 1)
    f = np.repeat(np.arange(F), P)
    return np.stack([f, x1, y1, x2, y2], 1).astype(np.int32)


# ---------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline leg (the only place outside tests/ that executes oracle/)
# ---------------------------------------------------------------------------------------------
def cpu_reference_step(model_name: str, h: int, w: int, frames: int, rois_per_frame: int, state=None, bn="train"):
    """One bounded sample of the workload on the host cores: the reference path restated (oracle YOLOv5 v6.0 + the ReID
    restatement with the same weights, BatchNorm in the reference's mode: batch statistics per frame's call)."""
    from oracle import reid as R
    from oracle import yolov5 as Y
    if state is None:
        torch.set_num_threads(os.cpu_count() or 1)
        rng = np.random.default_rng(0)
        ckpt = os.path.join(ROOT, "oracle", "_ref", "reid_ckpt.npz")
        boxes = synth_boxes(rng, frames, rois_per_frame, h, w)
        state = {"model": Y.build(model_name, seed=0), "rng": rng,
                 "reid_sd": R.load_state_dict(ckpt) if os.path.isfile(ckpt) else R.seeded_state_dict(0),
                 "imgs": [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(frames)],
                 "rois": boxes_to_rois(boxes, h, w)}
    t0 = time.perf_counter()
    Y.yolo_backbone_detect(state["model"], {"imgs": state["imgs"]}, size=max(h, w))
    if rois_per_frame > 0:
        for f in range(frames):
            bgr = state["imgs"][f][:, :, ::-1]
            crops = [bgr[y1:y2, x1:x2] for (_, x1, y1, x2, y2) in state["rois"][f * rois_per_frame:(f + 1) * rois_per_frame]]
            R.extract(state["reid_sd"], crops, bn)
    return time.perf_counter() - t0, state


def run_reference_arm(a) -> dict:
    frames = a.ref_frames
    dt, st = cpu_reference_step(a.model, a.h, a.w, frames, a.rois, None, a.reid_bn)      # warm-up (thread pools, allocations)
    for _ in range(max(a.warmup - 1, 0)):
        cpu_reference_step(a.model, a.h, a.w, frames, a.rois, st, a.reid_bn)
    times = [cpu_reference_step(a.model, a.h, a.w, frames, a.rois, st, a.reid_bn)[0] for _ in range(a.steps)]
    ms = 1e3 * sum(times) / len(times)
    fps = frames / (ms / 1e3)
    sample = (f"{frames} frames x (YOLOv5 {a.model} {a.h}x{a.w} fp32 oracle + {a.rois} ReID crops, BatchNorm {a.reid_bn}) per step, "
              f"torch CPU")
    return {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, frames),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def workload_config(a, batch):
    return {"workload": f"{a.model} {a.h}x{a.w} detect+decode+NMS" + (f", {a.rois} ROI/frame crop+resize+ReID CNN" if a.rois else " (detect only)")
            + f" ({a.config_name})",
            "frames_per_step_per_gpu": batch, "rois_per_frame": a.rois,
            "reid_bn": "batch statistics per frame's call (reference: Extractor never calls .eval())" if a.reid_bn == "train" else "folded (eval)",
            "conf": 0.25, "iou": 0.45, "max_det": 300,
            "sharding": "frames round-robin over ranks; one NCCL all-gather of int64[5] counters (frames, detections, crops, H2D probe, bound cores) at the end",
            "l2": "inputs rotate through a pool of distinct batches larger than L2; per-step activations (GBs) exceed L2"}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def _event_time(stream, fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        fn(i)
    e1.record(stream)
    return e0, e1


def per_launch_times(plans, reps=3):
    """CUDA events around every launch of the given plans (same streams): (conv ms, other ms, yolo conv ms) summed per step"""
    out = []
    for plan in plans:
        n = len(plan.steps)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(reps)]
        for r in range(reps):
            ev[r][0].record(plan.stream)
            for j, fn in enumerate(plan.steps):
                fn(plan.stream)
                ev[r][j + 1].record(plan.stream)
        torch.cuda.synchronize()
        per = np.array([[ev[r][j].elapsed_time(ev[r][j + 1]) for j in range(n)] for r in range(1, reps)]).mean(0)
        conv = sum(per[j] for j in range(n) if plan.step_flops[j] > 0)
        out.append((conv, per.sum() - conv))
    return out


def pin_to_gpu_local_cores(gpu_index: int, local_rank: int, local_world: int) -> list:
    """Multi-rank runs: bind this process to its share of the cores NVML reports as local to its GPU (ranks whose GPUs report the
    same core set -- one NUMA domain for all eight GPUs on the round-1 box -- split it evenly), so that the host side of the numpy
    surface (gathers, result conversion) and the H2D staging of a rank stay on memory next to its PCIe root.  Returns the cores."""
    try:
        all_cpus = sorted(os.sched_getaffinity(0))
        cpus = all_cpus
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            n = (max(all_cpus) + 64) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
            near = [c for c in all_cpus if (mask[c // 64] >> (c % 64)) & 1]
            if near:
                cpus = near
        except Exception:
            pass
        k = max(1, len(cpus) // max(local_world, 1))
        mine = cpus[(local_rank % max(len(cpus) // k, 1)) * k:][:k] or cpus
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return []


def h2d_probe_gbps(dev, nbytes: int = 80 << 20, reps: int = 5) -> float:
    """pinned -> device copy bandwidth of this rank (all ranks copy at the same time when the caller puts a barrier in front)"""
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) / 1e3) / 1e9


def run_ours(a) -> dict:
    import torch.distributed as dist
    from vehicle_counting_b200 import _lib as L
    from vehicle_counting_b200.engine import ReidEngine, YoloEngine
    from vehicle_counting_b200.modules import ImageDetect
    from vehicle_counting_b200.networks import yolo as NY
    from vehicle_counting_b200.networks.deepsort.deep_sort import Extractor
    from vehicle_counting_b200.pipeline import FramePipeline
    from vehicle_counting_b200.sharding import gather_counters, max_over_ranks
    from vehicle_counting_b200.weights import load_reid_state_dict, synth_reid_state_dict, synth_yolov5_state_dict

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = pin_to_gpu_local_cores(local, local, int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))) if world > 1 else []
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    L.init(local)                      # raises if libvcb200.so is missing or the device is not sm_100
    B, H, W, R_ = a.batch, a.h, a.w, a.rois
    rng = np.random.default_rng(1000 + rank)
    with_reid = R_ > 0

    ysd = synth_yolov5_state_dict(a.model, seed=0, obj_bias=a.obj_bias)
    ckpt = os.path.join(ROOT, "oracle", "_ref", "reid_ckpt.npz")
    reid_path = ckpt if os.path.isfile(ckpt) else "synthetic"
    rsd = load_reid_state_dict(ckpt) if os.path.isfile(ckpt) else synth_reid_state_dict(0)
    yolo = YoloEngine(ysd, B, H, W, device=str(dev), model_name=a.model)
    ncrops = B * R_
    reid = ReidEngine(rsd, capacity=max(ncrops, 8), device=str(dev), bn_mode=a.reid_bn, max_segments=max(B, 8)) if with_reid else None

    # synthetic stream shard: a pool of distinct batches (device copies for `value`, pinned host copies for the pipelined arm)
    pool_n = max(2, min(8, int(np.ceil(140e6 / (B * H * W * 3)))))
    host_pool = [torch.from_numpy(rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)).pin_memory() for _ in range(pool_n)]
    dev_pool = [h.to(dev) for h in host_pool]
    boxes = synth_boxes(rng, B, max(R_, 1), H, W)                  # float64 xyxy, what the tracker stage receives
    rois_np = boxes_to_rois(boxes, H, W) if with_reid else None
    seg = [R_] * B                                                  # one reference call (BatchNorm segment) per frame

    ys = yolo.plan.stream
    rs = reid.stream if with_reid else None

    def step_device(i):
        """inputs already in HBM: D2D from the pool into the plan's static input, detect, then ReID on the same frames"""
        with torch.cuda.stream(ys):
            yolo.frames.copy_(dev_pool[i % pool_n], non_blocking=True)
        yolo.forward()
        if with_reid:
            rs.wait_stream(ys)
            reid.run(yolo.frames, None, n=ncrops, seg_sizes=seg)
            ys.wait_stream(rs)

    if with_reid:                       # ROIs (and segment tables) resident on the device for the HBM-resident arm
        reid.run(yolo.frames, rois_np, seg_sizes=seg)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident ---------------------------------------------------------------------------------
    for i in range(a.warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    profile_region = os.environ.get("VCB_BENCH_PROFILE", "0") == "1"      # ncu --profile-from-start off: capture only the timed steps
    if profile_region:
        torch.cuda.profiler.start()
    e0, e1 = _event_time(ys, step_device, a.steps)
    barrier()
    if profile_region:
        torch.cuda.profiler.stop()
    ms_dev = e0.elapsed_time(e1) / a.steps
    clocks = sampler.stop()
    det_total = int(yolo.det_count.sum().item())

    # ---- folded-BN device-resident number (the designed fast path; not the reference's numerics) ------------------
    folded = None
    if with_reid and a.reid_bn == "train" and not a.quick:
        reid_f = ReidEngine(rsd, capacity=max(ncrops, 8), device=str(dev), bn_mode="eval")
        reid_f.run(yolo.frames, rois_np)

        def step_folded(i):
            with torch.cuda.stream(ys):
                yolo.frames.copy_(dev_pool[i % pool_n], non_blocking=True)
            yolo.forward()
            reid_f.stream.wait_stream(ys)
            reid_f.run(yolo.frames, None, n=ncrops)
            ys.wait_stream(reid_f.stream)
        for i in range(3):
            step_folded(i)
        barrier()
        f0, f1 = _event_time(ys, step_folded, a.steps)
        barrier()
        folded = f0.elapsed_time(f1) / a.steps
        folded_plan = next(iter(reid_f._plans.values()))["plan"]
        folded_times = per_launch_times([folded_plan])[0]
        del reid_f

    # ---- e2e through the reference-shaped surface ----------------------------------------------------------------
    # ImageDetect(args, config).run(batch) exactly as modules/__init__.py:57 calls it (lists of numpy frames in, dict of lists
    # out), then the tracker stage's feature call for the same frames (BGR numpy frames + float64 boxes in, numpy embeddings out)
    os.environ["VCB_REID_CAPACITY"] = str(max(ncrops, 512))
    monkey = NY.synth_yolov5_state_dict
    NY.synth_yolov5_state_dict = lambda name, seed=0: ysd              # same weights as the engine above (obj_bias included)
    os.environ["VCB_SYNTH_WEIGHTS"] = "1"
    cfg = types.SimpleNamespace(model_name=a.model, min_iou=0.45, min_conf=0.25, max_det=300)
    det_stage = ImageDetect(types.SimpleNamespace(weight=None, mapping=None, mapping_dict=None), cfg)
    det_stage.model.model.size = max(H, W)
    NY.synth_yolov5_state_dict = monkey
    extractor = Extractor(reid_path, use_cuda=True, bn_mode=a.reid_bn) if with_reid else None
    np_pool = [[hp.numpy()[k] for k in range(B)] for hp in host_pool]     # lists of HWC uint8 frames (RGB for the detector)
    np_pool_bgr = [[f[:, :, ::-1] for f in fr] for fr in np_pool] if a.bgr_views else np_pool
    boxes_list = [boxes[k] for k in range(B)]

    def step_plugin(i):
        frames = np_pool[i % pool_n]
        out = det_stage.run({"imgs": frames, "frames": list(range(B)), "ori_imgs": np_pool_bgr[i % pool_n]})
        n_det = sum(len(s) for s in out["scores"])
        n_feat = 0
        if with_reid:
            feats = extractor.from_frames(np_pool_bgr[i % pool_n], boxes_list)
            n_feat = sum(f.shape[0] for f in feats)
        return n_det, n_feat

    for i in range(max(3, min(a.warmup, 4))):
        step_plugin(i)
    barrier()
    t0 = time.perf_counter()
    n_det = n_feat = 0
    for i in range(a.steps):
        d_, f_ = step_plugin(i)
        n_det += d_; n_feat += f_
    torch.cuda.synchronize()
    ms_e2e = 1e3 * (time.perf_counter() - t0) / a.steps
    barrier()

    # ---- the reference's own granularity: one frame per call ------------------------------------------------------
    b1 = None
    if not a.quick and world == 1:
        one = [np_pool[0][0]]
        one_bgr = [np_pool_bgr[0][0]]
        for _ in range(5):
            det_stage.run({"imgs": one, "frames": [0], "ori_imgs": one_bgr})
            if with_reid:
                extractor.from_frames(one_bgr, boxes_list[:1])
        torch.cuda.synchronize()
        nfr = 64
        t0 = time.perf_counter()
        for k in range(nfr):
            fr = [np_pool[k % pool_n][k % B]]
            det_stage.run({"imgs": fr, "frames": [k], "ori_imgs": fr})
            if with_reid:
                extractor.from_frames(fr, boxes_list[k % B:k % B + 1])
        torch.cuda.synchronize()
        b1 = nfr / (time.perf_counter() - t0)

    # ---- pipelined host path (pinned tensors in / out, double buffered) -------------------------------------------
    pipe = FramePipeline(yolo, reid)
    ms_pipe = None
    if True:
        def run_pipe(nsteps):
            for i in range(nsteps):
                pipe.submit(host_pool[i % pool_n], rois_np, seg_sizes=seg if with_reid else None)
                if pipe.submitted - pipe.collected == 2:
                    pipe.collect()
            while pipe.collected < pipe.submitted:
                pipe.collect()
        run_pipe(3)
        barrier()
        t0 = time.perf_counter()
        run_pipe(a.steps)
        torch.cuda.synchronize()
        ms_pipe = 1e3 * (time.perf_counter() - t0) / a.steps
        barrier()

    # ---- per-kernel pass: CUDA events around every launch, for the roofline of the dominant kernel ----------------
    plans = [yolo.plan] + ([next(iter(reid._plans.values()))["plan"]] if with_reid else [])
    times = per_launch_times(plans)
    conv_ms = sum(t[0] for t in times)
    other_ms = sum(t[1] for t in times)
    conv_flops = sum(p.conv_flops for p in plans)
    launches = sum(p.graph.num_kernels for p in plans if p.graph is not None)

    ms_dev, ms_e2e = max_over_ranks([ms_dev, ms_e2e], device=dev)
    barrier()
    h2d_gbps = h2d_probe_gbps(dev)                      # every rank copies at the same time: the host-side ceiling of the e2e arm
    per_rank = gather_counters([a.steps * B, n_det, n_feat, int(h2d_gbps * 1000), len(cores)], device=dev)
    totals = [sum(c[i] for c in per_rank) for i in range(3)]

    peaks = _peaks()
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r02_launch_summary.json")) as f:
            ls = json.load(f)
        if ls.get("workload") == [a.model, H, W, B, R_, a.reid_bn]:
            traffic = int(ls["conv_kernels"]["dram_bytes_per_step"])
            traffic_src = "profiles/r02_launch_summary.json (ncu launch list of this command)"
    except Exception:
        pass
    out = None
    if rank == 0:
        fps = world * B / (ms_dev / 1e3)
        fps_e2e = world * B / (ms_e2e / 1e3)
        ach = conv_flops / (conv_ms / 1e3) / 1e12
        yolo_tf = yolo.plan.conv_flops / (times[0][0] / 1e3) / 1e12
        h2d = B * H * W * 3 * (2 if with_reid else 1)               # the detector's RGB batch and the tracker stage's BGR batch
        d2h = yolo.det_host.numel() * 4 + yolo.det_count_host.numel() * 4 + ncrops * 512 * 4
        out = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
               "data": "synthetic", "config": workload_config(a, B), "clocks": clocks,
               "e2e": {"value": fps_e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "path": "ImageDetect.run(batch of numpy frames) -> dict of lists; Extractor.from_frames(numpy BGR frames, float64 boxes) -> numpy embeddings"},
               "gpu_launches": int(launches * a.steps),
               "roofline": {"kernel": "conv_umma / conv_umma_2cta / conv_patch (all conv launches of one step)", "bound": "tensor", "achieved": ach,
                            "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tflops_sustained"],
                            "frac_of_burst": ach / peaks["tflops_burst"], "peak_source": peaks["source"], "traffic": traffic,
                            "traffic_source": traffic_src, "conv_gflop_per_step": conv_flops / 1e9, "conv_ms_per_step": conv_ms,
                            "other_kernels_ms_per_step": other_ms,
                            "yolo_frac": yolo_tf / peaks["tflops_sustained"], "yolo_tflops": yolo_tf, "yolo_conv_ms_per_step": times[0][0],
                            "whole_step_tensor_frac": conv_flops / (ms_dev / 1e3) / 1e12 / peaks["tflops_sustained"]},
               "counters": {"frames": totals[0], "detections": totals[1], "crops": totals[2], "detections_last_step_rank0": det_total},
               "host": {"h2d_probe_GBps_per_rank": [round(c[3] / 1000.0, 1) for c in per_rank], "pinned_cores_per_rank": [int(c[4]) for c in per_rank],
                        "note": "pinned->device copy rate with all ranks copying at once; ranks > 1 are bound to their share of the GPU-local cores"},
               "kernels_per_step": int(launches)}
        if with_reid:
            out["roofline"]["reid_tflops"] = plans[1].conv_flops / (times[1][0] / 1e3) / 1e12
            out["roofline"]["reid_ms_per_step"] = times[1][0] + times[1][1]
        if ms_pipe is not None:
            out["e2e_pipelined"] = {"value": world * B / (ms_pipe / 1e3), "unit": UNIT, "ms_per_step": ms_pipe,
                                    "path": "FramePipeline.submit/collect (pinned tensors, uploads overlap compute)"}
        if b1 is not None:
            out["dropin_b1"] = {"value": b1, "unit": UNIT, "path": "one frame per call: ImageDetect.run([frame]) + one ReID call for its boxes"}
        if folded is not None:
            out["folded_bn"] = {"value": world * B / (folded / 1e3), "unit": UNIT, "ms_per_step": folded,
                                "reid_ms_per_step": folded_times[0] + folded_times[1],
                                "note": "BatchNorm folded with running statistics: faster, but not what the reference computes"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


CONFIGS = {
    # name: (config_name, model, h, w, batch, rois)
    "default": ("BASELINE configs[3] per-GPU shard", "yolov5m", 640, 640, 64, 64),
    "1": ("BASELINE configs[1]: detect only", "yolov5s", 640, 640, 32, 0),
    "2": ("BASELINE configs[2]", "yolov5m", 1024, 1024, 32, 64),
    "4": ("BASELINE configs[4] detector at the reference's inference shape (size=640)", "yolov5l", 384, 640, 64, 64),
    "4b": ("BASELINE configs[4] detector at size=1280", "yolov5l", 736, 1280, 16, 64),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="default", choices=sorted(CONFIGS), help="BASELINE.json configuration preset")
    ap.add_argument("--model", default=None)
    ap.add_argument("--size", type=int, default=None, help="square inference size (overrides the preset)")
    ap.add_argument("--batch", type=int, default=None, help="frames per step per GPU")
    ap.add_argument("--rois", type=int, default=None, help="ReID crops per frame (0 = detect only)")
    ap.add_argument("--reid-bn", default="train", choices=["train", "eval"],
                    help="train = the reference's BatchNorm (batch statistics per call); eval = folded running statistics")
    ap.add_argument("--obj-bias", type=float, default=-3.0, help="synthetic Detect objectness bias (controls #candidates)")
    ap.add_argument("--ref-frames", type=int, default=4, help="frames per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the secondary measurements (folded BN, one-frame-per-call)")
    ap.add_argument("--bgr-views", action="store_true", help="hand the tracker stage reversed-channel VIEWS of the frames (as cv2 would not)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    name, model, h, w, batch, rois = CONFIGS[a.config]
    a.config_name = name
    a.model = a.model or model
    a.h, a.w = (a.size, a.size) if a.size else (h, w)
    a.batch = a.batch or batch
    a.rois = rois if a.rois is None else a.rois
    rank = int(os.environ.get("RANK", "0"))

    if a.impl == "reference":
        if rank != 0:
            return
        print(json.dumps(run_reference_arm(a)), flush=True)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the hot path has no CPU fallback")
    # stdout carries exactly ONE JSON line: anything libraries print while the run is in flight (e.g. NCCL's version banner,
    # which goes to fd 1) is diverted to stderr, and the real stdout is restored for the result line
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        out = run_ours(a)
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)
    if rank == 0 and out is not None:
        if int(os.environ.get("WORLD_SIZE", "1")) == 1 and not a.no_cpu_baseline:
            frames = 2
            dt, st = cpu_reference_step(a.model, a.h, a.w, frames, a.rois, None, a.reid_bn)
            dts = [cpu_reference_step(a.model, a.h, a.w, frames, a.rois, st, a.reid_bn)[0] for _ in range(3)]
            v = frames / (sum(dts) / len(dts))
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": f"3 x {frames} frames (YOLOv5 {a.model} {a.h}x{a.w} fp32 oracle + {a.rois} ReID crops/frame, BatchNorm {a.reid_bn}), torch CPU"}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
