#!/usr/bin/env python
"""Benchmark of the per-frame hot path: YOLOv5 detect + decode/NMS + ROI crop/resize + ReID embedding.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one JSON line on rank 0)
    python bench.py --impl reference --gpus N ...             # the CPU reference path (oracle port) on host cores
    torchrun --nnodes=1 --nproc-per-node N bench.py --gpus N  # one rank per GPU, frames sharded, weak scaling

Workload (BASELINE.json: "end-to-end FPS (detect+ReID+NMS) 640x640", config "YOLOv5m 640x640 ... frames sharded
over ranks"): per step and per GPU a batch of B synthetic random-uint8 640x640 frames goes through YOLOv5m
(seeded synthetic weights) + decode + NMS, and 64 synthetic ROIs per frame (w,h~U(32,256))
go through crop/resize/normalise + the ReID CNN (folded BatchNorm).  A "step" = one such batch.
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
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


def synth_rois(rng, frames: int, per_frame: int, size: int) -> np.ndarray:
    """w,h ~ U(32,256), box inside the frame; reference crop rule applied (int truncation)."""
    wh = rng.uniform(32, 256, (frames, per_frame, 2))
    tl = rng.uniform(0, 1, (frames, per_frame, 2)) * (size - wh)
    x1y1 = np.maximum(tl.astype(np.int64), 0)
    x2y2 = np.minimum((tl + wh).astype(np.int64), size - 1)
    f = np.repeat(np.arange(frames)[:, None, None], per_frame, 1)
    return np.concatenate([f, x1y1, x2y2], 2).reshape(-1, 5).astype(np.int32)


# ---------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline leg (the only place outside tests/ that executes oracle/)
# ---------------------------------------------------------------------------------------------
def cpu_reference_step(model_name: str, size: int, frames: int, rois_per_frame: int, state=None):
    """One bounded sample of the workload on the host cores: the reference path restated (oracle YOLOv5 v6.0 +
    the ReID restatement with the same weights layout, eval-mode BN like the GPU arm)."""
    from oracle import reid as R
    from oracle import yolov5 as Y
    if state is None:
        torch.set_num_threads(os.cpu_count() or 1)
        rng = np.random.default_rng(0)
        ckpt = os.path.join(ROOT, "oracle", "_ref", "reid_ckpt.npz")
        state = {"model": Y.build(model_name, seed=0), "rng": rng,
                 "reid_sd": R.load_state_dict(ckpt) if os.path.isfile(ckpt) else R.seeded_state_dict(0),
                 "imgs": [rng.integers(0, 256, (size, size, 3), dtype=np.uint8) for _ in range(frames)],
                 "rois": synth_rois(rng, frames, rois_per_frame, size)}
    t0 = time.perf_counter()
    Y.yolo_backbone_detect(state["model"], {"imgs": state["imgs"]}, size=size)
    for f in range(frames):
        bgr = state["imgs"][f][:, :, ::-1]
        crops = [bgr[y1:y2, x1:x2] for (_, x1, y1, x2, y2) in state["rois"][f * rois_per_frame:(f + 1) * rois_per_frame]]
        R.extract(state["reid_sd"], crops, "eval")
    return time.perf_counter() - t0, state


def run_reference_arm(a) -> dict:
    frames = a.ref_frames
    dt, st = cpu_reference_step(a.model, a.size, frames, a.rois, None)      # warm-up (thread pools, allocations)
    for _ in range(max(a.warmup - 1, 0)):
        cpu_reference_step(a.model, a.size, frames, a.rois, st)
    times = [cpu_reference_step(a.model, a.size, frames, a.rois, st)[0] for _ in range(a.steps)]
    ms = 1e3 * sum(times) / len(times)
    fps = frames / (ms / 1e3)
    sample = f"{frames} frames x (YOLOv5 {a.model} {a.size}x{a.size} fp32 oracle + {a.rois} ReID crops) per step, torch CPU"
    return {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(a, frames),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def workload_config(a, batch):
    return {"workload": f"{a.model} {a.size}x{a.size} detect+decode+NMS, {a.rois} ROI/frame crop+resize+ReID CNN (BASELINE configs[3] per-GPU shard)",
            "frames_per_step_per_gpu": batch, "rois_per_frame": a.rois, "reid_bn": "folded (eval)", "conf": 0.25, "iou": 0.45,
            "max_det": 300, "sharding": "frames round-robin over ranks; one NCCL all-gather of int64[3] counters at the end",
            "l2": "inputs rotate through a pool of distinct batches larger than L2; per-step activations (GBs) exceed L2"}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(a) -> dict:
    import torch.distributed as dist
    from vehicle_counting_b200 import _lib as L
    from vehicle_counting_b200.engine import ReidEngine, YoloEngine
    from vehicle_counting_b200.weights import load_reid_state_dict, synth_reid_state_dict, synth_yolov5_state_dict

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    L.init(local)                      # raises if libvcb200.so is missing or the device is not sm_100
    B, S, R_ = a.batch, a.size, a.rois
    rng = np.random.default_rng(1000 + rank)

    ysd = synth_yolov5_state_dict(a.model, seed=0, obj_bias=a.obj_bias)
    ckpt = os.path.join(ROOT, "oracle", "_ref", "reid_ckpt.npz")
    rsd = load_reid_state_dict(ckpt) if os.path.isfile(ckpt) else synth_reid_state_dict(0)
    yolo = YoloEngine(ysd, B, S, S, device=str(dev), model_name=a.model)
    reid = ReidEngine(rsd, capacity=B * R_, device=str(dev), bn_mode="eval")

    # synthetic stream shard: a pool of distinct batches (device copies for `value`, pinned host copies for `e2e`)
    pool_n = max(2, min(8, int(np.ceil(140e6 / (B * S * S * 3)))))
    host_pool = [torch.from_numpy(rng.integers(0, 256, (B, S, S, 3), dtype=np.uint8)).pin_memory() for _ in range(pool_n)]
    dev_pool = [h.to(dev) for h in host_pool]
    rois_np = synth_rois(rng, B, R_, S)
    ncrops = rois_np.shape[0]

    ys, rs = yolo.plan.stream, reid.stream

    def step_device(i):
        """inputs already in HBM: D2D from the pool into the plan's static input, detect, then ReID on the same frames"""
        with torch.cuda.stream(ys):
            yolo.frames.copy_(dev_pool[i % pool_n], non_blocking=True)
        yolo.forward()
        rs.wait_stream(ys)
        reid.run(yolo.frames, None, n=ncrops)
        ys.wait_stream(rs)

    from vehicle_counting_b200.pipeline import FramePipeline
    pipe = FramePipeline(yolo, reid)

    def run_e2e(nsteps):
        """host buffers in, host results out: every step's frames + ROIs are uploaded from pinned memory and its detections +
        embeddings are read back; uploads of step i+1 overlap the compute of step i (double buffering)"""
        n_det = n_feat = 0
        for i in range(nsteps):
            pipe.submit(host_pool[i % pool_n], rois_np)
            if pipe.submitted - pipe.collected == 2:
                det, cnt, feats = pipe.collect()
                n_det += int(cnt.sum()); n_feat += feats.shape[0]
        while pipe.collected < pipe.submitted:
            det, cnt, feats = pipe.collect()
            n_det += int(cnt.sum()); n_feat += feats.shape[0]
        return n_det, n_feat

    # ROIs resident on the device for the HBM-resident arm
    reid.run(yolo.frames, rois_np)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(a.warmup):
        step_device(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    profile_region = os.environ.get("VCB_BENCH_PROFILE", "0") == "1"      # ncu --profile-from-start off: capture only the timed steps
    if profile_region:
        torch.cuda.profiler.start()
    e0.record(ys)
    for i in range(a.steps):
        step_device(i)
    e1.record(ys)
    barrier()
    if profile_region:
        torch.cuda.profiler.stop()
    ms_dev = e0.elapsed_time(e1) / a.steps
    clocks = sampler.stop()
    det_total = int(yolo.det_count.sum().item())

    # end-to-end through host buffers
    run_e2e(max(a.warmup, 2))
    barrier()
    t0 = time.perf_counter()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record(pipe.copy_stream)
    n_det, n_feat = run_e2e(a.steps)
    ee1.record(ys)
    barrier()
    ms_e2e = max(ee0.elapsed_time(ee1), 1e3 * (time.perf_counter() - t0)) / a.steps

    # per-kernel pass: CUDA events around every conv launch (same streams), for the roofline of the dominant kernel
    conv_ms, other_ms = 0.0, 0.0
    for plan in (yolo.plan, next(iter(reid._plans.values()))["plan"]):
        n = len(plan.steps)
        reps = 3
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(n + 1)] for _ in range(reps)]
        for r in range(reps):
            ev[r][0].record(plan.stream)
            for j, fn in enumerate(plan.steps):
                fn(plan.stream)
                ev[r][j + 1].record(plan.stream)
        torch.cuda.synchronize()
        per = np.array([[ev[r][j].elapsed_time(ev[r][j + 1]) for j in range(n)] for r in range(1, reps)]).mean(0)
        for j in range(n):
            if plan.step_flops[j] > 0:
                conv_ms += per[j]
            else:
                other_ms += per[j]
    reid_plan = next(iter(reid._plans.values()))["plan"]
    conv_flops = yolo.plan.conv_flops + reid_plan.conv_flops
    launches = yolo.plan.graph.num_kernels + reid_plan.graph.num_kernels

    # max over ranks, counters all-gather (the only collective of the path)
    from vehicle_counting_b200.sharding import gather_counters, max_over_ranks
    ms_dev, ms_e2e = max_over_ranks([ms_dev, ms_e2e], device=dev)
    per_rank = gather_counters([a.steps * B, n_det, n_feat], device=dev)
    totals = [sum(c[i] for c in per_rank) for i in range(3)]

    peaks = _peaks()
    # DRAM bytes of the conv launches of one step, from the committed ncu launch list of this very command at the default workload
    # (profiles/r01_launch_summary.json: dram__bytes_read.sum + dram__bytes_write.sum summed over the conv kernels of a step)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_launch_summary.json")) as f:
            ls = json.load(f)
        if (a.model, a.size, B, R_) == ("yolov5m", 640, 64, 64):
            traffic = int(ls["conv_kernels"]["dram_bytes_per_step"])
            traffic_src = "profiles/r01_launch_summary.json (ncu, batch 64)"
    except Exception:
        pass
    out = None
    if rank == 0:
        fps = world * B / (ms_dev / 1e3)
        fps_e2e = world * B / (ms_e2e / 1e3)
        ach = conv_flops / (conv_ms / 1e3) / 1e12
        out = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
               "data": "synthetic", "config": workload_config(a, B), "clocks": clocks,
               "e2e": {"value": fps_e2e, "unit": UNIT, "ms_per_step": ms_e2e,
                       "h2d_bytes_per_step": int(B * S * S * 3 + rois_np.nbytes),
                       "d2h_bytes_per_step": int(yolo.det_host.numel() * 4 + yolo.det_count_host.numel() * 4 + ncrops * 512 * 4)},
               "gpu_launches": int(launches * a.steps),
               "roofline": {"kernel": "conv_umma_kernel (all conv launches of one step)", "bound": "tensor", "achieved": ach,
                            "peak": peaks["tflops_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["tflops_sustained"],
                            "frac_of_burst": ach / peaks["tflops_burst"], "peak_source": peaks["source"], "traffic": traffic, "traffic_source": traffic_src,
                            "conv_gflop_per_step": conv_flops / 1e9, "conv_ms_per_step": conv_ms, "other_kernels_ms_per_step": other_ms,
                            "whole_step_tensor_frac": conv_flops / (ms_dev / 1e3) / 1e12 / peaks["tflops_sustained"]},
               "counters": {"frames": totals[0], "detections": totals[1], "crops": totals[2], "detections_last_step_rank0": det_total},
               "kernels_per_step": int(launches)}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="yolov5m")
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--batch", type=int, default=64, help="frames per step per GPU")
    ap.add_argument("--rois", type=int, default=64, help="ReID crops per frame")
    ap.add_argument("--obj-bias", type=float, default=-3.0, help="synthetic Detect objectness bias (controls #candidates)")
    ap.add_argument("--ref-frames", type=int, default=4, help="frames per step of the CPU reference arm (bounded sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
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
            dt, st = cpu_reference_step(a.model, a.size, frames, a.rois, None)
            dts = [cpu_reference_step(a.model, a.size, frames, a.rois, st)[0] for _ in range(3)]
            v = frames / (sum(dts) / len(dts))
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": f"3 x {frames} frames (YOLOv5 {a.model} {a.size}x{a.size} fp32 oracle + {a.rois} ReID crops/frame), torch CPU"}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
