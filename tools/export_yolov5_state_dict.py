#!/usr/bin/env python
"""One-time export of an upstream YOLOv5 checkpoint to the tensor-only file vehicle_counting_b200 loads.

Upstream `yolov5*.pt` files (what the reference fetches through torch.hub, /root/reference/networks/yolo.py:14-17,58)
pickle the ultralytics `Model` object, so they can only be opened where the ultralytics/yolov5 sources are importable.
Run this there (e.g. inside a yolov5 checkout, or after `torch.hub.load('ultralytics/yolov5', 'custom', ...)` has
cached the sources):

    python tools/export_yolov5_state_dict.py yolov5s.pt yolov5s_sd.pt

The output holds {'state_dict': fp32 tensors in the v6.0 key layout, 'names': class names} and loads with
torch.load(weights_only=True); pass it as `--weight`.
"""
import sys

import torch


def main():
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    src, dst = sys.argv[1:]
    ckpt = torch.load(src, map_location="cpu", weights_only=False)
    model = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt
    if hasattr(model, "model") and hasattr(model, "names") and not hasattr(model, "state_dict"):
        model = model.model                      # AutoShape wrapper
    names = getattr(model, "names", None)
    if isinstance(names, dict):
        names = [names[i] for i in sorted(names)]
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in model.float().state_dict().items()}
    out = {"state_dict": sd}
    if names is not None:
        out["names"] = list(names)
    torch.save(out, dst)
    print(f"wrote {dst}: {len(sd)} tensors, {len(names) if names is not None else 0} class names")


if __name__ == "__main__":
    main()
