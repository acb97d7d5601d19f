"""Weights: loading checkpoints in the layouts the reference uses, and deterministic synthetic stand-ins.

The reference downloads `yolov5{s,m,l,x}.pt` (v6.0) at run time (/root/reference/networks/yolo.py:14-17,
utilities/utils.py:204-212) and ships `ckpt.t7` for the ReID net (feature_extractor.py:13).  Neither a
network nor a YOLO checkpoint exists on the GPU box, so benchmarks and tests use seeded synthetic
weights of the same architecture, in the same state_dict key layout.
"""
from __future__ import annotations

import json
import math
import os
from typing import Dict, Iterator, Optional, Tuple

import numpy as np
import torch

from .engine import LAYERS_V6, MODEL_SCALES, REID_BLOCKS, _ceil8

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def iter_yolov5_convs(name: str, nc: int = 80) -> Iterator[Tuple[str, int, int, int, bool]]:
    """(prefix, cout, cin, k, has_bn) for every conv of a v6.0 model, in module order."""
    gd, gw = MODEL_SCALES[name]
    ch = []
    for i, (f, n, kind, args) in enumerate(LAYERS_V6):
        def cin(j):
            return (3 if i == 0 else ch[i - 1]) if j == -1 else ch[j]
        n = max(round(n * gd), 1) if n > 1 else n
        p = f"model.{i}"
        if kind == "Conv":
            c2 = _ceil8(args[0] * gw)
            yield p, c2, cin(f), args[1], True
            ch.append(c2)
        elif kind == "C3":
            c1, c2 = cin(f), _ceil8(args[0] * gw)
            c_ = c2 // 2
            yield p + ".cv1", c_, c1, 1, True
            yield p + ".cv2", c_, c1, 1, True
            yield p + ".cv3", c2, 2 * c_, 1, True
            for j in range(n):
                yield f"{p}.m.{j}.cv1", c_, c_, 1, True
                yield f"{p}.m.{j}.cv2", c_, c_, 3, True
            ch.append(c2)
        elif kind == "SPPF":
            c1, c2 = cin(f), _ceil8(args[0] * gw)
            yield p + ".cv1", c1 // 2, c1, 1, True
            yield p + ".cv2", c2, 2 * c1, 1, True
            ch.append(c2)
        elif kind == "Up":
            ch.append(cin(f))
        elif kind == "Cat":
            ch.append(sum(cin(j) for j in f))
        elif kind == "Detect":
            for li, j in enumerate(f):
                yield f"{p}.m.{li}", 3 * (nc + 5), ch[j], 1, False
            ch.append(0)


def _load_calib(name: str) -> Dict[str, float]:
    path = os.path.join(_DATA, "synth_calib.json")
    if os.path.isfile(path):
        with open(path) as fh:
            return json.load(fh).get(name, {})
    return {}


def synth_yolov5_state_dict(name: str = "yolov5s", seed: int = 0, nc: int = 80, obj_bias: float = -4.0, cls_bias: float = 0.0,
                            calib: Optional[Dict[str, float]] = None) -> Dict[str, torch.Tensor]:
    """Seeded YOLOv5 v6.0 state_dict (unfused Conv+BN layout).  Conv weights ~ U(+-sqrt(3/fan_in)); BN
    gamma~U(.5,1.5), beta~N(0,.1), mean~N(0,.1)*sqrt(v), var~U(.5,1.5)*v, where v is the layer's conv-output
    variance from data/synth_calib.json (measured once with the CPU oracle by tools/make_synth_calib.py) so
    that activations stay O(1) through all ~60-100 layers; Detect biases give O(10^2) candidates/frame."""
    g = torch.Generator().manual_seed(seed)
    calib = _load_calib(name) if calib is None else calib
    sd: Dict[str, torch.Tensor] = {}
    for prefix, co, ci, k, has_bn in iter_yolov5_convs(name, nc):
        bound = math.sqrt(3.0 / (ci * k * k))
        w = (torch.rand(co, ci, k, k, generator=g) * 2 - 1) * bound
        if has_bn:
            v = float(calib.get(prefix, 1.0 / 3.0))
            sd[prefix + ".conv.weight"] = w
            sd[prefix + ".bn.weight"] = torch.rand(co, generator=g) + 0.5
            sd[prefix + ".bn.bias"] = torch.randn(co, generator=g) * 0.1
            sd[prefix + ".bn.running_mean"] = torch.randn(co, generator=g) * 0.1 * math.sqrt(v)
            sd[prefix + ".bn.running_var"] = (torch.rand(co, generator=g) + 0.5) * v
        else:
            sd[prefix + ".weight"] = w
            b = torch.zeros(3, nc + 5)
            b[:, 4] = obj_bias
            b[:, 5:] = cls_bias
            sd[prefix + ".bias"] = b.view(-1)
    return sd


def synth_reid_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded weights in ckpt.t7's `net_dict` key layout (model.py:48-98)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, co, ci, k, bias=False):
        bound = (6.0 / (ci * k * k)) ** 0.5
        sd[name + ".weight"] = (torch.rand(co, ci, k, k, generator=g) * 2 - 1) * bound
        if bias:
            sd[name + ".bias"] = torch.randn(co, generator=g) * 0.05

    def bn(name, c):
        sd[name + ".weight"] = torch.rand(c, generator=g) * 0.5 + 0.75
        sd[name + ".bias"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_var"] = torch.rand(c, generator=g) * 0.5 + 0.75

    conv("conv.0", 64, 3, 3, True); bn("conv.1", 64)
    for prefix, ci, co, down in REID_BLOCKS:
        conv(prefix + ".conv1", co, ci, 3); bn(prefix + ".bn1", co)
        conv(prefix + ".conv2", co, co, 3); bn(prefix + ".bn2", co)
        if down:
            conv(prefix + ".downsample.0", co, ci, 1); bn(prefix + ".downsample.1", co)
    return sd


def load_yolov5_checkpoint(path: str):
    """-> (state_dict, class names or None).

    Accepted: a plain v6.0 state_dict file, or {'state_dict' | 'model': state_dict, 'names': [...]} as written by
    tools/export_yolov5_state_dict.py.  The reference itself loads upstream `yolov5*.pt` files through
    torch.hub (networks/yolo.py:58): those pickle the upstream `Model` CLASS under ckpt['model'] (fp16 weights, `.names`),
    which can only be unpickled with the ultralytics/yolov5 sources on the path -- they are not available offline, so such a
    file is rejected here with the one-line export recipe instead of an opaque UnpicklingError."""
    try:
        obj = torch.load(path, map_location="cpu", weights_only=True)
    except Exception as e:     # pickle.UnpicklingError (unsupported global models.yolo.Model), or a zip/EOF error
        raise ValueError(
            f"{path}: cannot be read as a tensor-only checkpoint ({type(e).__name__}). Upstream yolov5*.pt files pickle the "
            "ultralytics Model class; convert once where the yolov5 sources are importable:\n"
            "    python tools/export_yolov5_state_dict.py yolov5s.pt yolov5s_sd.pt\n"
            "and pass the result as --weight.") from e
    names = None
    if isinstance(obj, dict) and "names" in obj:
        nm = obj["names"]
        names = [nm[i] for i in sorted(nm)] if isinstance(nm, dict) else list(nm)
    for key in ("state_dict", "model"):
        if isinstance(obj, dict) and key in obj and isinstance(obj[key], dict):
            obj = obj[key]
    if not (isinstance(obj, dict) and "model.0.conv.weight" in obj):
        raise ValueError(f"{path}: not a YOLOv5 v6.0 state_dict (expected key 'model.0.conv.weight')")
    sd = {k: v.float() if v.is_floating_point() else v for k, v in obj.items() if torch.is_tensor(v)}
    return sd, names


def load_yolov5_state_dict(path: str) -> Dict[str, torch.Tensor]:
    return load_yolov5_checkpoint(path)[0]


def load_reid_state_dict(path: str) -> Dict[str, torch.Tensor]:
    """ckpt.t7 (`torch.load(path)['net_dict']`, feature_extractor.py:13) or an .npz repack of it."""
    if path.endswith(".npz"):
        z = np.load(path)
        return {k: torch.from_numpy(z[k]) for k in z.files}
    obj = torch.load(path, map_location="cpu", weights_only=True)
    sd = obj.get("net_dict", obj)
    return {k: v.float() for k, v in sd.items() if not k.startswith("classifier") and "num_batches" not in k}
