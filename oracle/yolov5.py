"""ORACLE (test infrastructure, not product): CPU fp32 restatement of the YOLOv5 v6.0 detector.

PARITY UNPINNED for this half of the path: the arithmetic lives in the third-party
`ultralytics/yolov5` repository which the reference fetches at run time with
`torch.hub.load('ultralytics/yolov5', 'custom', path=weight, force_reload=True)`
(/root/reference/networks/yolo.py:58); it is not vendored, no checkpoint is on disk and there
is no network.  The only version evidence is the v6.0 weight URLs
(/root/reference/utilities/utils.py:204-209).  This file restates the published v6.0
algorithm (models/common.py Conv/Bottleneck/C3/SPPF/Focus/SPP/AutoShape, models/yolo.py
Detect/parse_model, models/yolov5{n,s,m,l,x}.yaml, utils/general.py non_max_suppression /
scale_coords / make_divisible, utils/augmentations.py letterbox) and the reference adapter
contract (/root/reference/networks/yolo.py:68-99).  Offline self-checks: fused parameter
counts 1 867 405 / 7 225 885 / 21 172 173 / 46 533 693 / 86 705 005 (n/s/m/l/x), asserted in
tests/test_oracle_yolo.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# models/yolov5*.yaml [upstream v6.0]: depth_multiple, width_multiple
MODEL_SCALES: Dict[str, Tuple[float, float]] = {
    "yolov5n": (0.33, 0.25),
    "yolov5s": (0.33, 0.50),
    "yolov5m": (0.67, 0.75),
    "yolov5l": (1.00, 1.00),
    "yolov5x": (1.33, 1.25),
}

ANCHORS_PX = (
    (10, 13, 16, 30, 33, 23),        # P3/8
    (30, 61, 62, 45, 59, 119),       # P4/16
    (116, 90, 156, 198, 373, 326),   # P5/32
)
STRIDES = (8, 16, 32)

# (from, number, module, args) rows of models/yolov5s.yaml v6.0 (backbone + head)
YAML_V6 = [
    (-1, 1, "Conv", (64, 6, 2, 2)),      # 0  P1/2
    (-1, 1, "Conv", (128, 3, 2)),        # 1  P2/4
    (-1, 3, "C3", (128,)),               # 2
    (-1, 1, "Conv", (256, 3, 2)),        # 3  P3/8
    (-1, 6, "C3", (256,)),               # 4
    (-1, 1, "Conv", (512, 3, 2)),        # 5  P4/16
    (-1, 9, "C3", (512,)),               # 6
    (-1, 1, "Conv", (1024, 3, 2)),       # 7  P5/32
    (-1, 3, "C3", (1024,)),              # 8
    (-1, 1, "SPPF", (1024, 5)),          # 9
    (-1, 1, "Conv", (512, 1, 1)),        # 10
    (-1, 1, "Upsample", ()),             # 11
    ((-1, 6), 1, "Concat", ()),          # 12
    (-1, 3, "C3", (512, False)),         # 13
    (-1, 1, "Conv", (256, 1, 1)),        # 14
    (-1, 1, "Upsample", ()),             # 15
    ((-1, 4), 1, "Concat", ()),          # 16
    (-1, 3, "C3", (256, False)),         # 17 -> Detect P3
    (-1, 1, "Conv", (256, 3, 2)),        # 18
    ((-1, 14), 1, "Concat", ()),         # 19
    (-1, 3, "C3", (512, False)),         # 20 -> Detect P4
    (-1, 1, "Conv", (512, 3, 2)),        # 21
    ((-1, 10), 1, "Concat", ()),         # 22
    (-1, 3, "C3", (1024, False)),        # 23 -> Detect P5
    ((17, 20, 23), 1, "Detect", ()),     # 24
]


# models/yolov5s.yaml of the v5.0 release (checkpoints <= v5.0): Focus stem, C3 x9 at P3, SPP before the last backbone C3
YAML_V5 = [
    (-1, 1, "Focus", (64, 3)),           # 0  P1/2
    (-1, 1, "Conv", (128, 3, 2)),        # 1  P2/4
    (-1, 3, "C3", (128,)),               # 2
    (-1, 1, "Conv", (256, 3, 2)),        # 3  P3/8
    (-1, 9, "C3", (256,)),               # 4
    (-1, 1, "Conv", (512, 3, 2)),        # 5  P4/16
    (-1, 9, "C3", (512,)),               # 6
    (-1, 1, "Conv", (1024, 3, 2)),       # 7  P5/32
    (-1, 1, "SPP", (1024, (5, 9, 13))),  # 8
    (-1, 3, "C3", (1024, False)),        # 9
] + YAML_V6[10:]


def make_divisible(x: float, divisor: int) -> int:
    """utils/general.py make_divisible [upstream]: ceil to a multiple of divisor."""
    return int(math.ceil(x / divisor) * divisor)


class Conv(nn.Module):
    """models/common.py Conv [upstream]: Conv2d(bias=False) -> BatchNorm2d(eps=1e-3, momentum=.03) -> SiLU."""

    def __init__(self, c1, c2, k=1, s=1, p=None, act=True):
        super().__init__()
        p = k // 2 if p is None else p
        self.conv = nn.Conv2d(c1, c2, k, s, p, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
        self.act = nn.SiLU() if act else nn.Identity()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class Bottleneck(nn.Module):
    def __init__(self, c1, c2, shortcut=True, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_, c2, 3, 1)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        return x + self.cv2(self.cv1(x)) if self.add else self.cv2(self.cv1(x))


class C3(nn.Module):
    def __init__(self, c1, c2, n=1, shortcut=True, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, e=1.0) for _ in range(n)))

    def forward(self, x):
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), dim=1))


class SPPF(nn.Module):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        x = self.cv1(x)
        y1 = self.m(x)
        y2 = self.m(y1)
        return self.cv2(torch.cat((x, y1, y2, self.m(y2)), 1))


class SPP(nn.Module):
    """<= v5.0 checkpoints: cv2(cat(x, mp5, mp9, mp13)); equals the SPPF cascade."""

    def __init__(self, c1, c2, k=(5, 9, 13)):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * (len(k) + 1), c2, 1, 1)
        self.m = nn.ModuleList([nn.MaxPool2d(kernel_size=x, stride=1, padding=x // 2) for x in k])

    def forward(self, x):
        x = self.cv1(x)
        return self.cv2(torch.cat([x] + [m(x) for m in self.m], 1))


class Focus(nn.Module):
    """<= v5.0 checkpoints: space-to-depth then Conv(4*c1, c2, k)."""

    def __init__(self, c1, c2, k=1, s=1, p=None, act=True):
        super().__init__()
        self.conv = Conv(c1 * 4, c2, k, s, p, act)

    def forward(self, x):
        return self.conv(torch.cat((x[..., ::2, ::2], x[..., 1::2, ::2], x[..., ::2, 1::2], x[..., 1::2, 1::2]), 1))


class Detect(nn.Module):
    """models/yolo.py Detect [upstream v6.0], inference branch only."""

    def __init__(self, nc=80, anchors=ANCHORS_PX, ch=()):
        super().__init__()
        self.nc = nc
        self.no = nc + 5
        self.nl = len(anchors)
        self.na = len(anchors[0]) // 2
        a = torch.tensor(anchors, dtype=torch.float32).view(self.nl, -1, 2)
        # upstream stores anchors in grid units (px / stride) in the checkpoint
        self.register_buffer("anchors", a / torch.tensor(STRIDES, dtype=torch.float32).view(-1, 1, 1))
        self.stride = torch.tensor(STRIDES, dtype=torch.float32)
        self.m = nn.ModuleList(nn.Conv2d(x, self.no * self.na, 1) for x in ch)

    def forward(self, feats: Sequence[torch.Tensor]):
        z, raw = [], []
        for i in range(self.nl):
            x = self.m[i](feats[i])                                 # [B, na*no, ny, nx]
            bs, _, ny, nx = x.shape
            raw.append(x)
            x = x.view(bs, self.na, self.no, ny, nx).permute(0, 1, 3, 4, 2).contiguous()
            yv, xv = torch.meshgrid(torch.arange(ny), torch.arange(nx), indexing="ij")
            grid = torch.stack((xv, yv), 2).view(1, 1, ny, nx, 2).float()
            anchor_grid = (self.anchors[i] * self.stride[i]).view(1, self.na, 1, 1, 2)
            y = x.sigmoid()
            xy = (y[..., 0:2] * 2.0 - 0.5 + grid) * self.stride[i]
            wh = (y[..., 2:4] * 2.0) ** 2 * anchor_grid
            y = torch.cat((xy, wh, y[..., 4:]), -1)
            z.append(y.view(bs, -1, self.no))
        return torch.cat(z, 1), raw


class DetectionModel(nn.Module):
    """models/yolo.py Model + parse_model [upstream v6.0]; `self.model` keeps upstream's module
    numbering so `state_dict()` keys are `model.{i}.…` exactly like a v6.0 checkpoint."""

    def __init__(self, name: str = "yolov5s", nc: int = 80, ch: int = 3, version: str = "v6"):
        super().__init__()
        gd, gw = MODEL_SCALES[name]
        self.version = version
        self.name, self.nc = name, nc
        layers, self.froms, outs = [], [], []

        def cin(f, i):
            return (ch if i == 0 else outs[i - 1]) if f == -1 else outs[f]

        for i, (f, n, m, args) in enumerate(YAML_V6 if version == "v6" else YAML_V5):
            n = max(round(n * gd), 1) if n > 1 else n
            if m in ("Focus", "SPP"):
                c1 = cin(f, i)
                c2 = make_divisible(args[0] * gw, 8)
                mod = Focus(c1, c2, args[1]) if m == "Focus" else SPP(c1, c2, args[1])
            elif m in ("Conv", "C3", "SPPF"):
                c1 = cin(f, i)
                c2 = make_divisible(args[0] * gw, 8)
                if m == "Conv":
                    mod = Conv(c1, c2, *args[1:])
                elif m == "C3":
                    mod = C3(c1, c2, n, *(args[1:]))
                else:
                    mod = SPPF(c1, c2, args[1])
            elif m == "Upsample":
                c2 = cin(f, i)
                mod = nn.Upsample(None, 2, "nearest")
            elif m == "Concat":
                c2 = sum(cin(x, i) for x in f)
                mod = nn.Identity()
            elif m == "Detect":
                c2 = 0
                mod = Detect(nc, ANCHORS_PX, [cin(x, i) for x in f])
            else:
                raise ValueError(m)
            layers.append(mod)
            self.froms.append(f)
            outs.append(c2)
        self.model = nn.ModuleList(layers)
        self.names = [f"class{i}" for i in range(nc)]

    def forward(self, x):
        y = []
        for i, m in enumerate(self.model):
            f = self.froms[i]
            if isinstance(f, int):
                xin = x if f == -1 else y[f]
            else:
                xin = [x if j == -1 else y[j] for j in f]
            if isinstance(m, nn.Identity):           # Concat(dim=1)
                x = torch.cat(xin, 1)
            elif isinstance(m, Detect):
                x = m(xin)
            else:
                x = m(xin)
            y.append(x)
        return x  # (pred [B,P,no], [raw heads])


def seeded_init_(model: DetectionModel, seed: int = 0, obj_bias: float = -4.0, cls_bias: float = 0.0,
                 head_gain: float = 1.0, calib_hw: int = 320) -> DetectionModel:
    """Deterministic synthetic weights (stand-in for the checkpoint the reference would download,
    /root/reference/networks/yolo.py:14-17). Conv weights ~ U(+-sqrt(3/fan_in)); BatchNorm
    gamma~U(.5,1.5), beta~N(0,.1), mean~N(0,.1)*sqrt(v), var~U(.5,1.5)*v where v is the layer's
    measured conv-output variance on a seeded noise image (one calibration forward, layer by
    layer) -- without it activations explode/vanish exponentially over ~60 layers and the parity
    test degenerates.  Detect: weights x sqrt(head_gain), objectness / class biases chosen so that
    random weights yield a realistic O(10^2) candidates per frame above conf=0.25 (upstream's
    _initialize_biases values would give zero)."""
    g = torch.Generator().manual_seed(seed)
    for mod in model.modules():
        if isinstance(mod, nn.Conv2d):
            fan_in = mod.in_channels * mod.kernel_size[0] * mod.kernel_size[1]
            bound = math.sqrt(3.0 / fan_in)
            mod.weight.data = (torch.rand(mod.weight.shape, generator=g) * 2 - 1) * bound
            if mod.bias is not None:
                mod.bias.data.zero_()
        elif isinstance(mod, nn.BatchNorm2d):
            n = mod.num_features
            mod.weight.data = torch.rand(n, generator=g) + 0.5
            mod.bias.data = torch.randn(n, generator=g) * 0.1
            mod.running_mean.data = torch.randn(n, generator=g) * 0.1
            mod.running_var.data = torch.rand(n, generator=g) + 0.5
    det: Detect = model.model[-1]
    for mi in det.m:
        mi.weight.data *= math.sqrt(head_gain)
        b = mi.bias.data.view(det.na, -1)
        b[:, 4] = obj_bias
        b[:, 5:] = cls_bias
        mi.bias.data = b.view(-1)
    model.eval()

    # calibration forward: scale each BN's (mean, var) by the measured conv-output variance
    def pre_hook(bn, inp):
        v = inp[0].var().item()
        bn.running_mean.data *= math.sqrt(v)
        bn.running_var.data *= v

    hs = [m.bn.register_forward_pre_hook(pre_hook) for m in model.modules() if isinstance(m, Conv)]
    with torch.no_grad():
        model(torch.rand(1, 3, calib_hw, calib_hw, generator=g))
    for h in hs:
        h.remove()
    return model


def build(name: str = "yolov5s", seed: int = 0, nc: int = 80, version: str = "v6", **kw) -> DetectionModel:
    return seeded_init_(DetectionModel(name, nc, version=version), seed, **kw)


def fp16_storage_twin(model: DetectionModel) -> DetectionModel:
    """The same algorithm with the CUDA path's STORAGE precision emulated, arithmetic still fp32 on the CPU:
    Conv+BN folded (upstream fuse_conv_and_bn) with the folded weights rounded to fp16, the input and every
    activation tensor rounded to fp16 where the CUDA path stores it (after SiLU, or after the shortcut add of a
    Bottleneck), Detect logits left in fp32.  Separates "fp16 storage" error (inherent to the precision the
    north star prescribes) from kernel error when comparing head tensors."""
    import copy
    m = copy.deepcopy(model).eval()
    r16 = lambda t: t.half().float()
    for mod in m.modules():
        if isinstance(mod, Conv):
            bn = mod.bn
            scale = bn.weight.data / torch.sqrt(bn.running_var.data + bn.eps)
            w = mod.conv.weight.data * scale.view(-1, 1, 1, 1)
            b = bn.bias.data - bn.running_mean.data * scale
            fused = nn.Conv2d(mod.conv.in_channels, mod.conv.out_channels, mod.conv.kernel_size, mod.conv.stride,
                              mod.conv.padding, bias=True)
            fused.weight.data = r16(w)
            fused.bias.data = b
            mod.conv = fused
            mod.bn = nn.Identity()
    for mod in m.modules():
        if isinstance(mod, Detect):
            for c in mod.m:
                c.weight.data = r16(c.weight.data)
    skip = set()
    for mod in m.modules():
        if isinstance(mod, Bottleneck) and mod.add:
            skip.add(mod.cv2)                               # rounded once, after the residual add
            mod.register_forward_hook(lambda _m, _i, o: r16(o))
    for mod in m.modules():
        if isinstance(mod, Conv) and mod not in skip:
            mod.register_forward_hook(lambda _m, _i, o: r16(o))
    m.model[0].register_forward_pre_hook(lambda _m, inp: (r16(inp[0]),))
    return m


def fused_param_count(model: DetectionModel) -> int:
    """Parameter count after Conv+BN fusion (upstream `model.fuse()` + model_info)."""
    n = 0
    for mod in model.modules():
        if isinstance(mod, Conv):
            n += mod.conv.weight.numel() + mod.conv.out_channels     # fused conv gets a bias
        elif isinstance(mod, Detect):
            n += sum(p.numel() for p in mod.m.parameters())
    return n


def conv_flops(model: DetectionModel, h: int, w: int) -> float:
    """2*MAC over every Conv2d at an h x w input (SURVEY §8(d) algorithmic work)."""
    total = [0.0]

    def hook(mod, inp, out):
        kh, kw = mod.kernel_size
        total[0] += 2.0 * out.numel() * mod.in_channels * kh * kw

    hs = [m.register_forward_hook(hook) for m in model.modules() if isinstance(m, nn.Conv2d)]
    with torch.no_grad():
        model(torch.zeros(1, 3, h, w))
    for x in hs:
        x.remove()
    return total[0]


# ---------------------------------------------------------------------------------------------
# AutoShape preprocessing [upstream models/common.py AutoShape.forward, utils/augmentations.py letterbox]
# ---------------------------------------------------------------------------------------------

def autoshape_shapes(shapes0: Sequence[Tuple[int, int]], size: int = 640, stride: int = 32):
    """shape1 (inference H,W) for a batch: per image g=size/max(h,w); max over batch; ceil to stride."""
    shape1 = []
    for (h, w) in shapes0:
        g = size / max(h, w)
        shape1.append([h * g, w * g])
    return [make_divisible(x, stride) for x in np.stack(shape1, 0).max(0)]


def letterbox_params(shape0: Tuple[int, int], new_shape: Tuple[int, int]):
    """Returns (new_unpad (w,h), top, bottom, left, right) of letterbox(auto=False, scaleup=True)."""
    h, w = shape0
    r = min(new_shape[0] / h, new_shape[1] / w)
    new_unpad = (int(round(w * r)), int(round(h * r)))
    dw, dh = (new_shape[1] - new_unpad[0]) / 2, (new_shape[0] - new_unpad[1]) / 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_unpad, top, bottom, left, right


def letterbox(im: np.ndarray, new_shape: Tuple[int, int], color=(114, 114, 114)) -> np.ndarray:
    import cv2
    new_unpad, top, bottom, left, right = letterbox_params(im.shape[:2], new_shape)
    if im.shape[:2][::-1] != new_unpad:
        im = cv2.resize(im, new_unpad, interpolation=cv2.INTER_LINEAR)
    return cv2.copyMakeBorder(im, top, bottom, left, right, cv2.BORDER_CONSTANT, value=color)


def preprocess(imgs: Sequence[np.ndarray], size: int = 640):
    """list of HWC uint8 RGB -> (x [B,3,H1,W1] fp32 in [0,1], shape0 list, shape1)."""
    shape0 = [im.shape[:2] for im in imgs]
    shape1 = autoshape_shapes(shape0, size)
    x = np.stack([letterbox(im, shape1) for im in imgs], 0)
    x = np.ascontiguousarray(x.transpose(0, 3, 1, 2))
    return torch.from_numpy(x).float() / 255.0, shape0, shape1


# ---------------------------------------------------------------------------------------------
# NMS [upstream utils/general.py non_max_suppression v6.0 + torchvision.ops.nms semantics]
# ---------------------------------------------------------------------------------------------

def greedy_nms(boxes: np.ndarray, scores: np.ndarray, iou_thres: float) -> np.ndarray:
    """torchvision.ops.nms semantics in float32: descending score (ties: lower index first),
    suppress iff IoU > thr (strict), areas (x2-x1)*(y2-y1)."""
    boxes = boxes.astype(np.float32)
    n = boxes.shape[0]
    order = np.lexsort((np.arange(n), -scores.astype(np.float32)))
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = ((x2 - x1) * (y2 - y1)).astype(np.float32)
    suppressed = np.zeros(n, bool)
    keep = []
    thr = np.float32(iou_thres)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest]); yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest]); yy2 = np.minimum(y2[i], y2[rest])
        w = np.maximum(np.float32(0), xx2 - xx1); h = np.maximum(np.float32(0), yy2 - yy1)
        inter = (w * h).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr > thr]] = True
    return np.asarray(keep, dtype=np.int64)


def non_max_suppression(pred: torch.Tensor, conf_thres=0.25, iou_thres=0.45, classes=None,
                        max_det=300, max_wh=4096, max_nms=30000) -> List[torch.Tensor]:
    """v6.0 non_max_suppression with multi_label=False, agnostic=False (reference settings,
    /root/reference/networks/yolo.py:62-66). The `time_limit` early exit is NOT replicated."""
    out = []
    for x in pred:                                   # per image [P, 85]
        x = x[x[:, 4] > conf_thres]
        if not x.shape[0]:
            out.append(torch.zeros((0, 6)))
            continue
        x = x.clone()
        x[:, 5:] *= x[:, 4:5]
        box = torch.empty_like(x[:, :4])
        box[:, 0] = x[:, 0] - x[:, 2] / 2
        box[:, 1] = x[:, 1] - x[:, 3] / 2
        box[:, 2] = x[:, 0] + x[:, 2] / 2
        box[:, 3] = x[:, 1] + x[:, 3] / 2
        conf, j = x[:, 5:].max(1, keepdim=True)
        x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
        if classes is not None:
            x = x[(x[:, 5:6] == torch.tensor(list(classes), dtype=torch.float32)).any(1)]
        n = x.shape[0]
        if not n:
            out.append(torch.zeros((0, 6)))
            continue
        if n > max_nms:
            x = x[x[:, 4].argsort(descending=True, stable=True)[:max_nms]]
        c = x[:, 5:6] * max_wh
        keep = greedy_nms((x[:, :4] + c).numpy(), x[:, 4].numpy(), iou_thres)[:max_det]
        out.append(x[torch.from_numpy(keep)])
    return out


def scale_coords(img1_shape, coords: torch.Tensor, img0_shape) -> torch.Tensor:
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    coords = coords.clone()
    coords[:, [0, 2]] -= pad[0]
    coords[:, [1, 3]] -= pad[1]
    coords[:, :4] /= gain
    coords[:, [0, 2]] = coords[:, [0, 2]].clamp(0, img0_shape[1])
    coords[:, [1, 3]] = coords[:, [1, 3]].clamp(0, img0_shape[0])
    return coords


@torch.no_grad()
def autoshape_forward(model: DetectionModel, imgs: Sequence[np.ndarray], size=640, conf=0.25, iou=0.45,
                      classes=None, max_det=300, max_wh=4096, return_raw=False):
    """AutoShape.forward on CPU fp32 (no autocast): list of RGB HWC uint8 -> list of [n,6] xyxy,conf,cls
    in original-image pixels."""
    x, shape0, shape1 = preprocess(imgs, size)
    pred, raw = model(x)
    dets = non_max_suppression(pred, conf, iou, classes, max_det, max_wh)
    for i in range(len(dets)):
        if dets[i].shape[0]:
            dets[i][:, :4] = scale_coords(shape1, dets[i][:, :4], shape0[i])
    return (dets, pred, raw) if return_raw else dets


def yolo_backbone_detect(model: DetectionModel, batch: dict, **kw) -> List[dict]:
    """Adapter output contract of /root/reference/networks/yolo.py:68-99."""
    out = []
    for d in autoshape_forward(model, batch["imgs"], **kw):
        if d.shape[0]:
            d = d.double().numpy()
            boxes = np.stack([d[:, 0], d[:, 1], d[:, 2] - d[:, 0], d[:, 3] - d[:, 1]], 1)
            out.append({"bboxes": boxes, "classes": d[:, 5].astype(np.int64), "scores": d[:, 4]})
        else:
            out.append({"bboxes": np.array(()), "classes": np.array(()), "scores": np.array(())})
    return out
