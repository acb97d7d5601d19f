"""Per-layer kernel choice, measured (run under gpurun on a B200):  python tools/autotune_layers.py  ->  vehicle_counting_b200/data/tuned_layers.json

Every convolution shape the BASELINE configurations launch (logged while the engines are built) is timed with the library's own
choice and with the variants the descriptor exposes -- N-tile width 64 / 128, single-CTA / CTA-pair (one or two clusters per SM
pair) / patch kernel -- CUDA events around 8 launches after 3 warm-ups, the variant's output checked against the default's.  A
variant enters the table only if it is >= 3 % faster.  The engine reads the table at plan time (engine.tuned_choice)."""
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["VCB_LOG_CONV_SHAPES"] = "1"
os.environ["VCB_TUNED"] = "0"
from vehicle_counting_b200 import _lib as L, ops                      # noqa: E402
from vehicle_counting_b200 import engine as E                          # noqa: E402
from vehicle_counting_b200.weights import synth_reid_state_dict, synth_yolov5_state_dict   # noqa: E402

DEV = torch.device("cuda:0")


DEFAULT_YOLO = (("yolov5m", 64, 640, 640), ("yolov5m", 128, 640, 640), ("yolov5s", 32, 640, 640), ("yolov5s", 8, 640, 640),
                ("yolov5m", 32, 1024, 1024), ("yolov5l", 64, 384, 640), ("yolov5l", 32, 384, 640), ("yolov5l", 16, 736, 1280))


def collect(yolo=DEFAULT_YOLO, reid=True, reid_ns=(64, 1024, 2048, 4096)):
    L.init(0)
    for name, b, h, w in yolo:
        eng = E.YoloEngine(synth_yolov5_state_dict(name, seed=0), b, h, w, model_name=name)
        del eng
        torch.cuda.empty_cache()
    rsd = synth_reid_state_dict(0)
    for mode in (("eval", "train") if reid else ()):
        r = E.ReidEngine(rsd, capacity=4096, bn_mode=mode, max_segments=64)
        for nc in reid_ns:                           # one frame per call, configs[4] at 16 frames, configs[2], the default step
            per = min(nc, 64)
            rois = np.zeros((nc, 5), np.int32); rois[:, 0] = np.repeat(np.arange(nc // per), per); rois[:, 3:] = 100
            r.run(torch.zeros(64, 640, 640, 3, dtype=torch.uint8, device=DEV), rois, seg_sizes=[per] * (nc // per))
            torch.cuda.synchronize()
        del r
        torch.cuda.empty_cache()
    uniq = {}
    for c in E.CONV_SHAPE_LOG:
        if c["a_mode"] != L.A_AUTO:
            continue
        uniq[tuple(sorted(c.items()))] = c
    return list(uniq.values())


def time_variant(c, block_n, cta_pair, ref=None):
    g = torch.Generator().manual_seed(1)
    n, h, w = c["n"], c["h"], c["w"]
    x = (torch.randn(n, h, w, c["cin_pitch"], generator=g)).half().to(DEV)
    wt = (torch.randn(c["cout"], c["cin"], c["k"], c["k"], generator=g) / (c["cin"] * c["k"] ** 2) ** 0.5).to(DEV)
    bias = (torch.randn(c["cout"], generator=g) * 0.3).to(DEV)
    d = ops.make_conv_desc(n, h, w, c["cin"], c["cout"], c["k"], c["s"], c["p"], cin_pitch=c["cin_pitch"], cout_pitch=c["cout_pitch"], act=c["act"],
                           res_mode=c["res"], res_pitch=c["cout_pitch"] if c["res"] else 0, out_dtype=c["out_dtype"], block_n=block_n, cta_pair=cta_pair)
    ho, wo = ops.conv_out_hw(d)
    wp, bp = ops.pack_conv_weights(d, wt, bias)
    y = torch.zeros(n, ho, wo, c["cout_pitch"], dtype=torch.float32 if c["out_dtype"] == L.F32 else torch.float16, device=DEV)
    res = (torch.randn(n, ho, wo, c["cout_pitch"], generator=g)).half().to(DEV) if c["res"] else None
    for _ in range(3):
        ops.conv2d(d, x, wp, bp, y, residual=res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        ops.conv2d(d, x, wp, bp, y, residual=res)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 8 * 1e3
    if tuple(L.last_fault()) != (0, 0, 0, 0):
        raise RuntimeError("kernel fault")
    yc = y[..., :c["cout"]].float()
    if ref is not None:
        err = (yc - ref).abs().max().item()
        if err > 2e-3 * max(1.0, ref.abs().max().item()):
            raise RuntimeError(f"variant output differs from the default's by {err}")
    return us, yc


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--only-k", type=int, default=0, help="1 or 3: tune only layers with this filter size and MERGE into the existing table")
    ap.add_argument("--pairs-1x1", default="1", help="cta_pair candidates for 1x1 layers (comma separated)")
    ap.add_argument("--only-f32", action="store_true", help="only the fp32-output layers (Detect heads); merges into the existing table")
    ap.add_argument("--reid-n", default="", help="crop counts for the ReID plans, e.g. 8,16,32 (implies merging)")
    ap.add_argument("--yolo", default="", help="name:batch:h:w,... instead of the BASELINE list (implies merging, no ReID shapes)")
    a = ap.parse_args()
    if a.yolo or a.reid_n:
        ys = tuple((t.split(":")[0],) + tuple(int(v) for v in t.split(":")[1:]) for t in a.yolo.split(",")) if a.yolo else ()
        shapes = collect(ys, reid=bool(a.reid_n), reid_ns=tuple(int(v) for v in a.reid_n.split(",")) if a.reid_n else ())
        a.merge_all = True
    else:
        shapes = collect()
        a.merge_all = False
    if a.only_k:
        shapes = [c for c in shapes if c["k"] == a.only_k]
    if a.only_f32:
        shapes = [c for c in shapes if c["out_dtype"] == L.F32]
        a.only_k = a.only_k or 1
    print(len(shapes), "distinct convolution shapes", flush=True)
    table, rows = {}, []
    path0 = os.path.join(ROOT, "vehicle_counting_b200", "data", "tuned_layers.json")
    if (a.only_k or a.merge_all) and os.path.isfile(path0):
        table = json.load(open(path0))["layers"]
    for c in sorted(shapes, key=lambda c: (c["k"], c["cin"], c["cout"], -c["n"] * c["h"] * c["w"])):
        ho = (c["h"] + 2 * c["p"] - c["k"]) // c["s"] + 1
        wo = (c["w"] + 2 * c["p"] - c["k"]) // c["s"] + 1
        m = c["n"] * ho * wo
        key = f'{c["k"]},{c["s"]},{c["cin"]},{c["cout"]},{int(c["res"] != 0)},{c["h"]}x{c["w"]},{int(round(math.log2(max(c["n"], 1))))}'
        try:
            base, ref = time_variant(c, 0, 0)
        except Exception as e:
            print(key, "default failed:", e)
            continue
        best = (base, 0, 0)
        tried = []
        if key in table:                     # merging: the entry to beat is the table's
            try:
                us, _ = time_variant(c, table[key][0], table[key][1], ref)
                if us < best[0]:
                    best = (us, table[key][0], table[key][1])
            except Exception:
                pass
        bns = [bn for bn in (64, 128) if bn < c["cout"] and c["cout"] % bn == 0]
        pairs = [1, 2, 4] + ([5] if (c["k"], c["s"], c["p"]) == (3, 1, 1) else []) if c["k"] > 1 else [int(v) for v in a.pairs_1x1.split(",")]
        for bn in [0] + bns:
            for cp in [0] + pairs:
                if bn == 0 and cp == 0:
                    continue
                try:
                    us, _ = time_variant(c, bn, cp, ref)
                except Exception as e:
                    L.load().vcb_last_error_string()
                    continue
                tried.append((round(us, 1), bn, cp))
                if us < best[0]:
                    best = (us, bn, cp)
        gain = 1.0 - best[0] / base
        rows.append({"key": key, "shape": c, "default_us": round(base, 1), "best_us": round(best[0], 1), "block_n": best[1], "cta_pair": best[2],
                     "tried": sorted(tried)[:4]})
        if gain >= 0.03 and (best[1], best[2]) != (0, 0):
            prev = table.get(key)
            if prev is None or prev[3] > best[0] or a.only_k or a.merge_all:
                table[key] = [best[1], best[2], round(base, 1), round(best[0], 1)]
        print(f'{key:28s} M={m:8d} default {base:7.1f} us  best {best[0]:7.1f} us (block_n={best[1]}, cta_pair={best[2]})  {"*" if key in table else ""}', flush=True)
    out = {"device": torch.cuda.get_device_name(0), "note": "key = k,s,cin,cout,res,HxW of the input,round(log2(n)); value = [block_n, cta_pair, default us, tuned us]",
           "layers": table}
    path = os.path.join(ROOT, "vehicle_counting_b200", "data", "tuned_layers.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    with open(os.path.join(ROOT, "gpurun_out", "autotune_rows.jsonl"), "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
    print("wrote", path, len(table), "entries;", "sum default", round(sum(r["default_us"] for r in rows)), "us, sum best", round(sum(r["best_us"] for r in rows)), "us")
    # the file must travel back from the GPU box
    import shutil
    shutil.copy(path, os.path.join(ROOT, "gpurun_out", "tuned_layers.json"))


if __name__ == "__main__":
    main()
