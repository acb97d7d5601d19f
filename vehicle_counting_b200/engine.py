"""Static execution plans for the two networks on the hot path.

`YoloEngine`  -- YOLOv5 v6.0 DetectionModel forward + Detect decode + NMS + scale_coords for a fixed
                 (batch, H, W): what `YoloBackbone.detect` needs behind
                 /root/reference/networks/yolo.py:70 (the AutoShape call).
`ReidEngine`  -- ROI crop/resize/normalise + the DeepSORT `Net` (reid=True) forward: what
                 `Extractor.__call__` computes (/root/reference/networks/deepsort/deep/
                 feature_extractor.py:42-47, model.py:83-95).

Both build a list of C-ABI calls over statically planned NHWC fp16 buffers (concat-free: producers
write channel slices of the consumer's buffer), then capture the list into one CUDA graph that is
replayed per batch.  PyTorch only owns the memory and the stream; BN folding / weight packing use
torch elementwise ops once at load time, never per frame.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L
from . import ops

# ---------------------------------------------------------------------------------------------
# YOLOv5 v6.0 topology (models/yolov5*.yaml, upstream) -- the product's own statement of it
# ---------------------------------------------------------------------------------------------
ANCHORS_PX = ((10, 13, 16, 30, 33, 23), (30, 61, 62, 45, 59, 119), (116, 90, 156, 198, 373, 326))
STRIDES = (8, 16, 32)
MODEL_SCALES = {"yolov5n": (0.33, 0.25), "yolov5s": (0.33, 0.50), "yolov5m": (0.67, 0.75),
                "yolov5l": (1.00, 1.00), "yolov5x": (1.33, 1.25)}
# (from, repeats, kind, args): Conv(c, k, s, p) / C3(c, shortcut) / SPPF(c, k) / Up / Cat / Detect
LAYERS_V6 = [
    (-1, 1, "Conv", (64, 6, 2, 2)), (-1, 1, "Conv", (128, 3, 2, 1)), (-1, 3, "C3", (128, True)),
    (-1, 1, "Conv", (256, 3, 2, 1)), (-1, 6, "C3", (256, True)), (-1, 1, "Conv", (512, 3, 2, 1)),
    (-1, 9, "C3", (512, True)), (-1, 1, "Conv", (1024, 3, 2, 1)), (-1, 3, "C3", (1024, True)),
    (-1, 1, "SPPF", (1024, 5)), (-1, 1, "Conv", (512, 1, 1, 0)), (-1, 1, "Up", ()), ((-1, 6), 1, "Cat", ()),
    (-1, 3, "C3", (512, False)), (-1, 1, "Conv", (256, 1, 1, 0)), (-1, 1, "Up", ()), ((-1, 4), 1, "Cat", ()),
    (-1, 3, "C3", (256, False)), (-1, 1, "Conv", (256, 3, 2, 1)), ((-1, 14), 1, "Cat", ()),
    (-1, 3, "C3", (512, False)), (-1, 1, "Conv", (512, 3, 2, 1)), ((-1, 10), 1, "Cat", ()),
    (-1, 3, "C3", (1024, False)), ((17, 20, 23), 1, "Detect", ()),
]
# checkpoints <= v5.0 (north_star: "CSP/Focus conv stacks ... SPP"): Focus stem (space-to-depth + 3x3 conv == the s2d stem path
# below with a channel permutation), nine C3 repeats at P3, SPP(5, 9, 13) before the last backbone C3 (== the SPPF cascade:
# mp9 = mp5 o mp5, mp13 = mp5 o mp5 o mp5, exactly)
LAYERS_V5 = [
    (-1, 1, "Focus", (64, 3)), (-1, 1, "Conv", (128, 3, 2, 1)), (-1, 3, "C3", (128, True)),
    (-1, 1, "Conv", (256, 3, 2, 1)), (-1, 9, "C3", (256, True)), (-1, 1, "Conv", (512, 3, 2, 1)),
    (-1, 9, "C3", (512, True)), (-1, 1, "Conv", (1024, 3, 2, 1)), (-1, 1, "SPP", (1024, 5)), (-1, 3, "C3", (1024, False)),
] + LAYERS_V6[10:]
YOLO_BN_EPS = 1e-3     # upstream initialize_weights() sets eps=1e-3 on every BatchNorm2d


def _ceil8(x: float) -> int:
    return int(math.ceil(x / 8) * 8)


def infer_version(sd: Dict[str, torch.Tensor]) -> str:
    """'v5' for checkpoints with the Focus stem (model.0.conv.conv.weight), else 'v6'"""
    return "v5" if "model.0.conv.conv.weight" in sd else "v6"


def focus_weights_to_s2d(w: torch.Tensor) -> torch.Tensor:
    """Focus concatenates the four pixel phases as (dy, dx) = (0,0), (1,0), (0,1), (1,1) [upstream models/common.py Focus]; the
    s2d ingest stores them as (dy*2 + dx).  [cout, 12, 3, 3] -> [cout, 16, 3, 3] in the ingest's channel order."""
    co = w.shape[0]
    out = torch.zeros(co, 16, 3, 3, dtype=w.dtype, device=w.device)
    for dy in range(2):
        for dx in range(2):
            out[:, (dy * 2 + dx) * 3:(dy * 2 + dx) * 3 + 3] = w[:, (dx * 2 + dy) * 3:(dx * 2 + dy) * 3 + 3]
    return out


def infer_model_name(sd: Dict[str, torch.Tensor]) -> str:
    """Recover (depth, width) multiples from a state_dict: stem width and C3 repeat count."""
    c0 = sd["model.0.conv.conv.weight" if infer_version(sd) == "v5" else "model.0.conv.weight"].shape[0]
    n2 = len({k.split(".")[3] for k in sd if k.startswith("model.2.m.")})
    for name, (gd, gw) in MODEL_SCALES.items():
        if _ceil8(64 * gw) == c0 and max(round(3 * gd), 1) == n2:
            return name
    raise ValueError(f"not a YOLOv5 v6.0 n/s/m/l/x state_dict (stem width {c0}, C3 repeats {n2})")


def fold_conv_bn(sd, prefix: str, eps: float, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Conv2d(bias=False)+BatchNorm2d -> (w', b') fp32 on `device` [upstream fuse_conv_and_bn].
    Already-fused checkpoints (conv.bias present, no bn.*) pass through."""
    w = sd[prefix + ".conv.weight"].to(device=device, dtype=torch.float32)
    if prefix + ".bn.weight" in sd:
        g = sd[prefix + ".bn.weight"].to(device=device, dtype=torch.float32)
        b = sd[prefix + ".bn.bias"].to(device=device, dtype=torch.float32)
        m = sd[prefix + ".bn.running_mean"].to(device=device, dtype=torch.float32)
        v = sd[prefix + ".bn.running_var"].to(device=device, dtype=torch.float32)
        s = g / torch.sqrt(v + eps)
        return w * s.view(-1, 1, 1, 1), b - m * s
    bias = sd.get(prefix + ".conv.bias")
    bias = torch.zeros(w.shape[0], device=device) if bias is None else bias.to(device=device, dtype=torch.float32)
    return w, bias


def stem_weights_to_s2d(w: torch.Tensor) -> torch.Tensor:
    """[cout, 3, 6, 6] stride-2 stem -> the equivalent [cout, 16, 3, 3] stride-1 kernel over the space-to-depth input:
    w'[o, (dy*2+dx)*3 + c, a, b] = w[o, c, 2a+dy, 2b+dx]; channels 12..15 are zero."""
    co = w.shape[0]
    w6 = w.view(co, 3, 3, 2, 3, 2)                      # [o, c, a, dy, b, dx]
    w2 = w6.permute(0, 3, 5, 1, 2, 4).reshape(co, 12, 3, 3)   # [o, (dy, dx, c), a, b]
    out = torch.zeros(co, 16, 3, 3, dtype=w.dtype, device=w.device)
    out[:, :12] = w2
    return out


@dataclass
class TRef:
    """A channel slice of an NHWC fp16 buffer."""
    buf: torch.Tensor       # [n, h, w, pitch]
    c0: int
    c: int

    @property
    def pitch(self) -> int:
        return self.buf.shape[-1]

    @property
    def h(self) -> int:
        return self.buf.shape[1]

    @property
    def w(self) -> int:
        return self.buf.shape[2]

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr() + self.c0 * self.buf.element_size()

    def view(self) -> torch.Tensor:
        return self.buf[..., self.c0:self.c0 + self.c]


class _Plan:
    """Ordered list of launch closures + bookkeeping of algorithmic work."""

    def __init__(self, device: torch.device):
        self.device = device
        self.steps: List[Callable[[object], None]] = []
        self.labels: List[str] = []
        self.step_flops: List[float] = []
        self.conv_flops = 0.0
        self.num_convs = 0
        self.keep: list = []          # tensors that must outlive the plan (weights, descriptors)
        self.graph: Optional[ops.Graph] = None
        self.stream = torch.cuda.Stream(device=device)
        # consecutive convolutions walk their tiles in opposite directions, so a layer starts on the part of its input that its
        # producer wrote last (still in the 126 MB L2) instead of the part written first (evicted): +1.5-2.5 % on the whole step
        # (profiles/r02_epilogue_analysis.md); $VCB_TILE_REV=0 turns it off
        self.alternate = os.environ.get("VCB_TILE_REV", "1") == "1"
        self._rev = False

    def add(self, fn: Callable[[object], None], label: str = "op", flops: float = 0.0) -> None:
        self.steps.append(fn)
        self.labels.append(label)
        self.step_flops.append(flops)

    def conv(self, x: TRef, n: int, w: torch.Tensor, b: Optional[torch.Tensor], y: TRef, k: int, s: int, p: int, act: int,
             residual: Optional[TRef] = None, res_mode: int = L.RES_NONE, out_dtype: int = L.F16, a_mode: int = L.A_AUTO,
             flops: Optional[float] = None, wcache: Optional[dict] = None, wkey=None, stats=None) -> bool:
        """`flops`: algorithmic FLOPs of the layer when the launched geometry is a re-expression of it (stems).
        `wcache` / `wkey`: packed weights are a function of the layer, not of the batch: plans of different batch sizes share them.
        `stats` = (seg_of_image tensor, sums tensor): the launch also accumulates train-mode BatchNorm statistics from its epilogue
        (vcb_conv2d_fwd_stats); returns False when the geometry cannot do that (the plain launch is queued, the caller adds the
        separate statistics kernel)."""
        cout, cin = int(w.shape[0]), int(w.shape[1])
        m_out = n * ((x.h + 2 * p - k) // s + 1) * ((x.w + 2 * p - k) // s + 1)
        tb, tp = tuned_choice(k, s, cin, cout, residual is not None, m_out, n, x.h, x.w) if (a_mode == L.A_AUTO and stats is None) else (0, 0)
        if CONV_SHAPE_LOG is not None:
            CONV_SHAPE_LOG.append(dict(n=n, h=x.h, w=x.w, cin=cin, cin_pitch=x.pitch, cout=cout, cout_pitch=y.pitch, k=k, s=s, p=p, act=act,
                                       res=(res_mode if residual is not None else L.RES_NONE), out_dtype=out_dtype, a_mode=a_mode))
        d = ops.make_conv_desc(n, x.h, x.w, cin, cout, k, s, p, cin_pitch=x.pitch, cout_pitch=y.pitch, act=act,
                               res_mode=res_mode if residual is not None else L.RES_NONE,
                               res_pitch=residual.pitch if residual is not None else 0, out_dtype=out_dtype, a_mode=a_mode,
                               tile_rev=self.alternate and self._rev, block_n=tb, cta_pair=tp)
        self._rev = not self._rev
        ho, wo = ops.conv_out_hw(d)
        assert (ho, wo) == (y.h, y.w), ((ho, wo), (y.h, y.w))
        assert x.c == cin or (cin <= 4 and x.pitch == 4), (x.c, cin)

        if wcache is not None and wkey in wcache:
            wp, bp = wcache[wkey]
        else:
            wp, bp = ops.pack_conv_weights(d, w, b)
            if wcache is not None:
                wcache[wkey] = (wp, bp)
        xp, yp, rp = x.ptr, y.ptr, (residual.ptr if residual is not None else 0)
        self.keep += [d, wp, bp, x.buf, y.buf] + ([residual.buf] if residual is not None else [])
        fl = 2.0 * n * ho * wo * cout * cin * k * k if flops is None else flops
        self.conv_flops += fl
        self.num_convs += 1
        label = f"conv{k}x{k}s{s} {cin}->{cout} M={n * ho * wo}" + (" +res" if residual is not None else "")
        if stats is not None:
            seg_t, sums_t = stats
            try:                    # one eager launch decides: geometries without the split epilogue answer VCB_ERR_INVALID
                ops.conv2d_stats(d, xp, wp, bp, yp, seg_t, sums_t, stream=self.stream)
                torch.cuda.synchronize(self.device)
                self.add(lambda st, d=d, xp=xp, wp=wp, bp=bp, yp=yp: ops.conv2d_stats(d, xp, wp, bp, yp, seg_t, sums_t, stream=st),
                         label + " +BN stats", fl)
                return True
            except L.VcbError:
                pass
        self.add(lambda st, d=d, xp=xp, wp=wp, bp=bp, yp=yp, rp=rp: ops.conv2d(d, xp, wp, bp, yp, residual=rp, stream=st), label, fl)
        return stats is None

    def run_eager(self, stream=None) -> None:
        st = stream if stream is not None else self.stream
        for fn in self.steps:
            fn(st)

    def capture(self) -> "ops.Graph":
        torch.cuda.synchronize(self.device)
        ops.graph_begin(self.stream)
        try:
            for fn in self.steps:
                fn(self.stream)
        finally:
            self.graph = ops.graph_end(self.stream)
        return self.graph

    def run(self, use_graph: bool = True) -> None:
        """Enqueue one pass on the plan's stream."""
        if use_graph:
            if self.graph is None:
                self.capture()
            self.graph.launch(self.stream)
        else:
            self.run_eager(self.stream)


_ROWWIN_OK: Optional[bool] = None

# N-tile widths that beat the library's choice (one tile as wide as fits 256 columns), measured per layer shape on B200
# (tests/bringup_conv.py --only bnsweep, profiles/r02_bnsweep.jsonl): a 192- or 256-wide tile leaves room for ONE accumulator stage in
# the 256 TMEM columns of a two-per-SM CTA, so MMA and epilogue serialise; narrower tiles keep two stages at the price of fetching
# the activation tile once per N tile -- a win only while K is small.  (k, stride, cin, cout) -> block_n
TUNED_BLOCK_N = {(1, 1, 192, 192): 64, (1, 1, 384, 384): 128, (3, 2, 384, 768): 128}
# ... and where the CTA-pair kernel wins although the library's rule (two waves of pair tiles) does not pick it: 3x3 384->384 at
# M = 25 600 (one wave and a third of 256-row tiles): 83.0 -> 74.9 us (profiles/r02_bnsweep.jsonl).  (k, stride, cin, cout) -> cta_pair,
# applied from 16 384 output pixels up
TUNED_CTA_PAIR = {(3, 1, 384, 384): 2}

# The complete per-layer table: data/tuned_layers.json, written by tools/autotune_layers.py on a B200 (every convolution shape of the
# BASELINE configurations x {N-tile width, single / pair / patch kernel}, each timed with CUDA events and checked against the
# library's own choice; an entry exists only where a variant was >= 3 % faster).  Key "k,s,cin,cout,res,HxW,round(log2(n))" (input
# image size and batch bucket: the same GEMM shape behaves differently on 7x7 crops and on 46x80 frames) -> [block_n, cta_pair].
# $VCB_TUNED=0 ignores it (and the two small tables above remain).
_TUNED_TABLE: Optional[dict] = None
CONV_SHAPE_LOG: Optional[list] = [] if os.environ.get("VCB_LOG_CONV_SHAPES") == "1" else None


def tuned_choice(k: int, s: int, cin: int, cout: int, has_res: bool, m: int, n: int = 0, h: int = 0, w: int = 0) -> Tuple[int, int]:
    """(block_n, cta_pair) for a convolution of this shape; (0, 0) = the library's own choice"""
    global _TUNED_TABLE
    if _TUNED_TABLE is None:
        _TUNED_TABLE = {}
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "tuned_layers.json")
        if os.environ.get("VCB_TUNED", "1") != "0" and os.path.isfile(path):
            import json
            with open(path) as f:
                _TUNED_TABLE = {kk: tuple(v[:2]) for kk, v in json.load(f).get("layers", {}).items()}
    key = f"{k},{s},{cin},{cout},{int(has_res)},{h}x{w},{int(round(math.log2(max(n, 1))))}"      # input image size + batch bucket
    if key in _TUNED_TABLE:
        return _TUNED_TABLE[key]
    if os.environ.get("VCB_TUNED", "1") == "0":
        return 0, 0
    return TUNED_BLOCK_N.get((k, s, cin, cout), 0), (TUNED_CTA_PAIR.get((k, s, cin, cout), 0) if m >= 16384 else 0)


def SILU_DEFAULT() -> bool:
    return os.environ.get("VCB_SILU", "exp") != "tanh"


class YoloEngine:
    """YOLOv5 v6.0 detector for a fixed batch of `batch` frames at inference size (h, w) (multiples of 32)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], batch: int, h: int, w: int, *, device="cuda:0", conf=0.25, iou=0.45,
                 max_det=300, max_wh=4096.0, max_nms=30000, model_name: Optional[str] = None, fp32_logits: bool = True,
                 a_mode: int = L.A_AUTO, classes: Optional[Sequence[int]] = None):
        assert h % 32 == 0 and w % 32 == 0, "inference shape must be a multiple of the max stride (32)"
        self.device = torch.device(device)
        L.init(self.device.index or 0)
        self.batch, self.h, self.w = batch, h, w
        self.conf, self.iou, self.max_det, self.max_wh, self.max_nms = conf, iou, max_det, max_wh, max_nms
        self.name = model_name or infer_model_name(state_dict)
        # $VCB_FP16_LOGITS=1: Detect logits stored as fp16 (what the reference's autocast path holds) instead of fp32
        self.fp32_logits = fp32_logits and os.environ.get("VCB_FP16_LOGITS", "0") != "1"
        self.classes = None if classes is None else sorted(int(c) for c in classes)     # upstream non_max_suppression(classes=...)
        self.a_mode = a_mode
        self.fuse_c3 = os.environ.get("VCB_C3_FUSE", "1") != "0"      # C3: cv1|cv2 as one GEMM, bottlenecks in place
        self.version = infer_version(state_dict)
        self.layers = LAYERS_V5 if self.version == "v5" else LAYERS_V6
        gd, gw = MODEL_SCALES[self.name]
        det_w = state_dict["model.24.m.0.weight"]
        self.nc = det_w.shape[0] // 3 - 5
        self.no = self.nc + 5
        if "model.24.anchors" in state_dict:      # checkpoint anchors are in grid units
            a = state_dict["model.24.anchors"].float().cpu() * torch.tensor(STRIDES, dtype=torch.float32).view(-1, 1, 1)
            self.anchors_px = [tuple(float(v) for v in a[i].flatten()) for i in range(3)]
        else:
            self.anchors_px = [tuple(float(v) for v in ANCHORS_PX[i]) for i in range(3)]
        self.plan = _Plan(self.device)
        self._build(state_dict, gd, gw)

    # -- buffer helpers ------------------------------------------------------------------------
    def _buf(self, h, w, c, dtype=torch.float16) -> torch.Tensor:
        return torch.zeros(self.batch, h, w, c, dtype=dtype, device=self.device)

    def _probe_rowwin(self) -> bool:
        """One tiny row-window convolution: does the driver encode a tensor map whose windows overlap?"""
        global _ROWWIN_OK
        if _ROWWIN_OK is None:
            try:
                dev = self.device
                x = torch.zeros(1 * 8 * 10 * 16 + 16, dtype=torch.float16, device=dev)
                y = torch.zeros(1, 8, 8, 16, dtype=torch.float16, device=dev)
                d = ops.make_conv_desc(1, 8, 8, 16, 16, 3, 1, 1, cin_pitch=16, cout_pitch=16, act=L.ACT_SILU, a_mode=L.A_ROWWIN)
                wp, bp = ops.pack_conv_weights(d, torch.zeros(16, 16, 3, 3, device=dev), None)
                ops.conv2d(d, x, wp, bp, y)
                torch.cuda.synchronize(dev)
                _ROWWIN_OK = True
            except L.VcbError:
                _ROWWIN_OK = False
        return _ROWWIN_OK

    def _build(self, sd, gd, gw) -> None:
        B, dev, plan = self.batch, self.device, self.plan
        LAYERS = self.layers
        # 1. channel / spatial bookkeeping
        ch: List[int] = []
        hw: List[Tuple[int, int]] = []
        reps: List[int] = []
        for i, (f, n, kind, args) in enumerate(LAYERS):
            def cin(j):
                return (3 if i == 0 else ch[i - 1]) if j == -1 else ch[j]

            def shp(j):
                return ((self.h, self.w) if i == 0 else hw[i - 1]) if j == -1 else hw[j]
            n = max(round(n * gd), 1) if n > 1 else n
            reps.append(n)
            if kind == "Conv":
                c2, k, s, p = _ceil8(args[0] * gw), args[1], args[2], args[3]
                hh, ww = shp(f)
                ch.append(c2); hw.append(((hh + 2 * p - k) // s + 1, (ww + 2 * p - k) // s + 1))
            elif kind == "Focus":
                hh, ww = shp(f)
                ch.append(_ceil8(args[0] * gw)); hw.append((hh // 2, ww // 2))
            elif kind in ("C3", "SPPF", "SPP"):
                ch.append(_ceil8(args[0] * gw)); hw.append(shp(f))
            elif kind == "Up":
                ch.append(cin(f)); hw.append((shp(f)[0] * 2, shp(f)[1] * 2))
            elif kind == "Cat":
                ch.append(sum(cin(j) for j in f)); hw.append(shp(f[0]))
            else:
                ch.append(0); hw.append((0, 0))
        # 2. homes: an output that feeds a Cat lives inside the Cat's buffer (concat-free)
        cat_buf: Dict[int, torch.Tensor] = {}
        home: Dict[int, TRef] = {}
        for i, (f, n, kind, args) in enumerate(LAYERS):
            if kind != "Cat":
                continue
            cat_buf[i] = self._buf(hw[i][0], hw[i][1], ch[i])
            off = 0
            for j in f:
                src = i - 1 if j == -1 else j
                assert src not in home, "an output may feed only one Concat"
                home[src] = TRef(cat_buf[i], off, ch[src])
                off += ch[src]
            home[i] = TRef(cat_buf[i], 0, ch[i])
        for i, (f, n, kind, args) in enumerate(LAYERS):
            if i not in home and kind not in ("Detect",):
                home[i] = TRef(self._buf(hw[i][0], hw[i][1], ch[i]), 0, ch[i])
        self.layer_out = home

        # 3. input ingest
        # uint8 frames -> fp16 space-to-depth NHWC16 (12 used): the 6x6/s2/p2 stem becomes a 3x3/s1/p1 conv over 16
        # channels that the im2col TMA feeds like any other layer (csrc/pointwise.cu frames_to_f16_s2d_kernel)
        self.frames = torch.zeros(B, self.h, self.w, 3, dtype=torch.uint8, device=dev)
        h2, w2 = self.h // 2, self.w // 2
        # Row-window stem ($VCB_STEM_ROWWIN=0 turns it off): the ingest writes a W-padded buffer and the stem fetches one TMA box per
        # filter ROW (3 per tile instead of 9 im2col boxes); falls back to the im2col stem if the driver rejects the overlapping map
        self.stem_rowwin = (os.environ.get("VCB_STEM_ROWWIN", "1") != "0" and self.a_mode == L.A_AUTO and SILU_DEFAULT()
                            and self._probe_rowwin())
        if self.stem_rowwin:
            self._s2d_flat = torch.zeros(B * h2 * (w2 + 2) * 16 + 16, dtype=torch.float16, device=dev)
            self.x_s2d = torch.as_strided(self._s2d_flat, (B, h2, w2, 16), (h2 * (w2 + 2) * 16, (w2 + 2) * 16, 16, 1))
            plan.add(lambda st: ops.frames_to_f16_s2d_wpad(self.frames, self._s2d_flat, stream=st), "ingest u8->f16 s2d (W-padded)")
        else:
            self.x_s2d = torch.zeros(B, h2, w2, 16, dtype=torch.float16, device=dev)
            plan.add(lambda st: ops.frames_to_f16_s2d(self.frames, self.x_s2d, stream=st), "ingest u8->f16 s2d")

        def src_of(i, f) -> TRef:
            return TRef(self.x_s2d, 0, 16) if (i == 0 and f == -1) else home[i - 1 if f == -1 else f]

        # $VCB_SILU=tanh: one-MUFU SiLU (h + h*tanh(h)); default: ex2 + rcp
        SILU = L.ACT_SILU_TANH if os.environ.get("VCB_SILU", "exp") == "tanh" else L.ACT_SILU
        for i, (f, n, kind, args) in enumerate(LAYERS):
            pre = f"model.{i}"
            if kind == "Conv":
                k, s, p = args[1], args[2], args[3]
                w_, b_ = fold_conv_bn(sd, pre, YOLO_BN_EPS, dev)
                if i == 0:
                    assert (k, s, p) == (6, 2, 2) and w_.shape[1] == 3
                    w_, k, s, p = stem_weights_to_s2d(w_), 3, 1, 1
                    stem_flops = 2.0 * B * hw[0][0] * hw[0][1] * w_.shape[0] * 3 * 36      # the 6x6 conv's own count
                    plan.conv(src_of(i, f), B, w_, b_, home[i], k, s, p, SILU, a_mode=L.A_ROWWIN if self.stem_rowwin else self.a_mode,
                              flops=stem_flops)
                    continue
                plan.conv(src_of(i, f), B, w_, b_, home[i], k, s, p, SILU, a_mode=self.a_mode)
            elif kind == "C3":
                x = src_of(i, f)
                c2, shortcut = ch[i], args[1]
                c_ = c2 // 2
                hh, ww = hw[i]
                cat = self._buf(hh, ww, 2 * c_)
                nrep = reps[i]
                tmp = TRef(self._buf(hh, ww, c_), 0, c_)
                if self.fuse_c3:
                    # cv1 and cv2 are 1x1 convolutions of the SAME input: one GEMM with N = 2*c_ (concatenated weights) writes both
                    # halves of the concat buffer, so the input is read once.  The bottleneck chain then runs IN PLACE on the left
                    # half: y <- y + cv2'(cv1'(y)) reads y only through the 1x1 (into tmp) and as the residual of the very tile it
                    # overwrites (the epilogue stages the residual tile before it stores the result).
                    wa, ba = fold_conv_bn(sd, pre + ".cv1", YOLO_BN_EPS, dev)
                    wb, bb = fold_conv_bn(sd, pre + ".cv2", YOLO_BN_EPS, dev)
                    plan.conv(x, B, torch.cat([wa, wb], 0), torch.cat([ba, bb], 0), TRef(cat, 0, 2 * c_), 1, 1, 0, SILU, a_mode=self.a_mode)
                    cur = TRef(cat, 0, c_)
                    for j in range(nrep):
                        w1, b1 = fold_conv_bn(sd, f"{pre}.m.{j}.cv1", YOLO_BN_EPS, dev)
                        plan.conv(cur, B, w1, b1, tmp, 1, 1, 0, SILU, a_mode=self.a_mode)
                        w2, b2 = fold_conv_bn(sd, f"{pre}.m.{j}.cv2", YOLO_BN_EPS, dev)
                        plan.conv(tmp, B, w2, b2, cur, 3, 1, 1, SILU, residual=cur if shortcut else None,
                                  res_mode=L.RES_AFTER_ACT, a_mode=self.a_mode)
                else:
                    # cv2 -> right half of the concat buffer
                    w_, b_ = fold_conv_bn(sd, pre + ".cv2", YOLO_BN_EPS, dev)
                    plan.conv(x, B, w_, b_, TRef(cat, c_, c_), 1, 1, 0, SILU, a_mode=self.a_mode)
                    # cv1 -> chain of bottlenecks -> left half
                    ping = [TRef(self._buf(hh, ww, c_), 0, c_), TRef(self._buf(hh, ww, c_), 0, c_)]
                    cur = ping[0]
                    w_, b_ = fold_conv_bn(sd, pre + ".cv1", YOLO_BN_EPS, dev)
                    plan.conv(x, B, w_, b_, cur, 1, 1, 0, SILU, a_mode=self.a_mode)
                    for j in range(nrep):
                        dst = TRef(cat, 0, c_) if j == nrep - 1 else ping[(j + 1) & 1]
                        w1, b1 = fold_conv_bn(sd, f"{pre}.m.{j}.cv1", YOLO_BN_EPS, dev)
                        plan.conv(cur, B, w1, b1, tmp, 1, 1, 0, SILU, a_mode=self.a_mode)
                        w2, b2 = fold_conv_bn(sd, f"{pre}.m.{j}.cv2", YOLO_BN_EPS, dev)
                        plan.conv(tmp, B, w2, b2, dst, 3, 1, 1, SILU, residual=cur if shortcut else None,
                                  res_mode=L.RES_AFTER_ACT, a_mode=self.a_mode)
                        cur = dst
                w_, b_ = fold_conv_bn(sd, pre + ".cv3", YOLO_BN_EPS, dev)
                plan.conv(TRef(cat, 0, 2 * c_), B, w_, b_, home[i], 1, 1, 0, SILU, a_mode=self.a_mode)
            elif kind == "Focus":
                w_, b_ = fold_conv_bn(sd, pre + ".conv", YOLO_BN_EPS, dev)
                assert tuple(w_.shape[1:]) == (12, 3, 3), "Focus(c1=3, k=3) expected"
                plan.conv(src_of(i, f), B, focus_weights_to_s2d(w_), b_, home[i], 3, 1, 1, SILU,
                          a_mode=L.A_ROWWIN if self.stem_rowwin else self.a_mode, flops=2.0 * B * hw[0][0] * hw[0][1] * w_.shape[0] * 12 * 9)
            elif kind in ("SPPF", "SPP"):
                x = src_of(i, f)
                c_ = x.c // 2
                hh, ww = hw[i]
                cat = self._buf(hh, ww, 4 * c_)
                w_, b_ = fold_conv_bn(sd, pre + ".cv1", YOLO_BN_EPS, dev)
                plan.conv(x, B, w_, b_, TRef(cat, 0, c_), 1, 1, 0, SILU, a_mode=self.a_mode)
                plan.add(lambda st, cat=cat, c_=c_, hh=hh, ww=ww: ops.sppf_pool(cat, 4 * c_, B, hh, ww, c_, stream=st), "sppf 3x maxpool5")
                w_, b_ = fold_conv_bn(sd, pre + ".cv2", YOLO_BN_EPS, dev)
                plan.conv(TRef(cat, 0, 4 * c_), B, w_, b_, home[i], 1, 1, 0, SILU, a_mode=self.a_mode)
            elif kind == "Up":
                x = src_of(i, f)
                y = home[i]
                plan.add(lambda st, x=x, y=y: ops.upsample2x(x.ptr, x.pitch, y.ptr, y.pitch, B, x.h, x.w, x.c, stream=st), "upsample2x")
            elif kind == "Cat":
                pass                                   # producers already wrote their slices
            elif kind == "Detect":
                self._build_head(sd, [home[j] for j in f])

    def _build_head(self, sd, feats: Sequence[TRef]) -> None:
        B, dev, plan = self.batch, self.device, self.plan
        no3 = 3 * self.no
        pitch = (no3 + 7) // 8 * 8
        ldt = torch.float32 if self.fp32_logits else torch.float16
        self.logits = []
        dd = L.DetectDesc()
        dd.n, dd.nc, dd.num_levels = B, self.nc, len(feats)
        dd.logits_dtype = L.F32 if self.fp32_logits else L.F16
        dd.conf_thres = float(self.conf)
        if self.classes is not None:        # dropped in the decode kernel, i.e. before the max_nms / max_det cuts (as upstream)
            dd.use_class_mask = 1
            for c in self.classes:
                if 0 <= c < 256:
                    dd.class_mask[c >> 5] |= (1 << (c & 31))
        P = 0
        for li, ft in enumerate(feats):
            lg = torch.zeros(B, ft.h, ft.w, pitch, dtype=ldt, device=dev)
            self.logits.append(lg)
            w_ = sd[f"model.24.m.{li}.weight"].to(device=dev, dtype=torch.float32)
            b_ = sd[f"model.24.m.{li}.bias"].to(device=dev, dtype=torch.float32)
            plan.conv(ft, B, w_, b_, TRef(lg, 0, no3), 1, 1, 0, L.ACT_NONE, out_dtype=dd.logits_dtype, a_mode=self.a_mode)
            lv = dd.level[li]
            lv.logits, lv.pitch, lv.ny, lv.nx = lg.data_ptr(), pitch, ft.h, ft.w
            lv.stride = float(self.h // ft.h)
            for a in range(3):
                lv.anchor_w[a] = self.anchors_px[li][2 * a]
                lv.anchor_h[a] = self.anchors_px[li][2 * a + 1]
            P += 3 * ft.h * ft.w
        self.num_pred = P
        dd.max_candidates = P
        self.cand_box = torch.zeros(B, P, 4, dtype=torch.float32, device=dev)
        self.cand_score = torch.zeros(B, P, dtype=torch.float32, device=dev)
        self.cand_cls = torch.zeros(B, P, dtype=torch.int32, device=dev)
        self.cand_index = torch.zeros(B, P, dtype=torch.int32, device=dev)
        self.cand_count = torch.zeros(B, dtype=torch.int32, device=dev)
        self.nms_ws = torch.zeros(max(ops.nms_workspace_bytes(B, P) // 8, 1), dtype=torch.int64, device=dev)
        self.det = torch.zeros(B, self.max_det, 6, dtype=torch.float32, device=dev)
        self.det_count = torch.zeros(B, dtype=torch.int32, device=dev)
        # scale_coords parameters per frame: gain, pad_x, pad_y, w0, h0
        self.scale = torch.zeros(5, B, dtype=torch.float32, device=dev)
        self.scale_host = torch.zeros(5, B, dtype=torch.float32).pin_memory()
        self.set_identity_scale()
        nd = L.NmsDesc()
        nd.n, nd.max_candidates, nd.max_det = B, P, self.max_det
        nd.iou_thres, nd.max_wh, nd.max_nms = float(self.iou), float(self.max_wh), int(self.max_nms)
        nd.gain, nd.pad_x, nd.pad_y, nd.w0, nd.h0 = (self.scale[i].data_ptr() for i in range(5))
        self._dd, self._nd = dd, nd
        plan.add(lambda st: ops.detect_decode(dd, self.cand_box, self.cand_score, self.cand_cls, self.cand_index,
                                              self.cand_count, stream=st), "detect decode")
        plan.add(lambda st: ops.nms(nd, self.cand_box, self.cand_score, self.cand_cls, self.cand_index, self.cand_count,
                                    self.nms_ws, self.det, self.det_count, stream=st), "nms")
        # pinned host mirrors of the result
        self.det_host = torch.zeros(B, self.max_det, 6, dtype=torch.float32).pin_memory()
        self.det_count_host = torch.zeros(B, dtype=torch.int32).pin_memory()

    # -- per-call API --------------------------------------------------------------------------
    def set_identity_scale(self) -> None:
        self.scale_host[0].fill_(1.0); self.scale_host[1].zero_(); self.scale_host[2].zero_()
        self.scale_host[3].fill_(float(self.w)); self.scale_host[4].fill_(float(self.h))
        self.scale.copy_(self.scale_host)

    def set_scale(self, shapes0: Sequence[Tuple[int, int]]) -> None:
        """scale_coords parameters for original shapes (h0, w0) letterboxed into (self.h, self.w); a repeated list of shapes (every
        batch of a video) costs one tuple comparison."""
        key = tuple(shapes0)
        if key == getattr(self, "_scale_key", None):
            return
        hw = np.asarray(shapes0, np.float64).reshape(-1, 2)
        gain = np.minimum(self.h / hw[:, 0], self.w / hw[:, 1])
        sh = self.scale_host.numpy()
        n = len(hw)
        self.plan.stream.synchronize()                  # the previous copy out of scale_host has completed
        sh[0, :n] = gain
        sh[1, :n] = (self.w - hw[:, 1] * gain) / 2
        sh[2, :n] = (self.h - hw[:, 0] * gain) / 2
        sh[3, :n] = hw[:, 1]
        sh[4, :n] = hw[:, 0]
        with torch.cuda.stream(self.plan.stream):
            self.scale.copy_(self.scale_host, non_blocking=True)
        self._scale_key = key

    def upload(self, frames_host: torch.Tensor) -> None:
        """H2D of a pinned uint8 [batch, h, w, 3] tensor on the plan's stream."""
        with torch.cuda.stream(self.plan.stream):
            self.frames.copy_(frames_host, non_blocking=True)

    def forward(self, use_graph: bool = True) -> None:
        """frames (device) -> det/det_count (device); asynchronous on the plan's stream."""
        self.plan.run(use_graph)

    def download(self) -> Tuple[np.ndarray, np.ndarray]:
        with torch.cuda.stream(self.plan.stream):
            self.det_host.copy_(self.det, non_blocking=True)
            self.det_count_host.copy_(self.det_count, non_blocking=True)
        self.plan.stream.synchronize()
        return self.det_host.numpy(), self.det_count_host.numpy()

    @property
    def conv_flops_per_frame(self) -> float:
        return self.plan.conv_flops / self.batch


# ---------------------------------------------------------------------------------------------
# ReID
# ---------------------------------------------------------------------------------------------
REID_SIZE = 50
REID_MEAN = (0.485, 0.456, 0.406)
REID_STD = (0.229, 0.224, 0.225)
REID_BN_EPS = 1e-5
REID_BLOCKS = [("layer1.0", 64, 64, False), ("layer1.1", 64, 64, False), ("layer2.0", 64, 128, True),
               ("layer2.1", 128, 128, False), ("layer3.0", 128, 256, True), ("layer3.1", 256, 256, False),
               ("layer4.0", 256, 512, True), ("layer4.1", 512, 512, False)]


def _fold_plain(w, bias, bn, eps, device):
    """conv weight [+bias] and BN tensors (gamma, beta, mean, var) -> folded (w', b') fp32."""
    w = w.to(device=device, dtype=torch.float32)
    g, b, m, v = (t.to(device=device, dtype=torch.float32) for t in bn)
    s = g / torch.sqrt(v + eps)
    b0 = torch.zeros_like(m) if bias is None else bias.to(device=device, dtype=torch.float32)
    return w * s.view(-1, 1, 1, 1), (b0 - m) * s + b


def _pad_cin(w: torch.Tensor, cin: int) -> torch.Tensor:
    out = torch.zeros(w.shape[0], cin, w.shape[2], w.shape[3], dtype=w.dtype, device=w.device)
    out[:, :w.shape[1]] = w
    return out


class ReidEngine:
    """DeepSORT appearance CNN over up to `capacity` crops per call.

    bn_mode="eval": BatchNorm folded into the convolutions (running statistics) -- the fast path, one dense
    batch, results independent of batch composition.  bn_mode="train": the reference as shipped (it never
    calls .eval()): per-call batch statistics; crops are grouped in segments = one reference
    `Extractor.__call__` each (all detections of one class in one frame)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], capacity: int = 64, *, device="cuda:0", bn_mode: str = "eval",
                 a_mode: int = L.A_AUTO, max_segments: int = 128):
        assert bn_mode in ("eval", "train")
        # folded-BN path: fused crop -> stem conv -> ReLU -> max-pool ($VCB_REID_FUSED_STEM=0 keeps the three separate kernels)
        self.fused_stem = os.environ.get("VCB_REID_FUSED_STEM", "1") != "0"
        # ... fed straight from the frames (csrc/reid_stem_direct.cu); $VCB_REID_STEM=patches keeps the round-2 pair of kernels that
        # stage the im2col operand in HBM
        self.stem_direct = os.environ.get("VCB_REID_STEM", "direct") != "patches"
        self.epi_stats = os.environ.get("VCB_EPI_STATS", "0") != "0"      # train-mode BN statistics from the conv epilogue (measured: no gain, off)
        self.device = torch.device(device)
        L.init(self.device.index or 0)
        self.sd = {k: v for k, v in state_dict.items() if not k.startswith("classifier")}
        self.capacity = capacity
        self.bn_mode = bn_mode
        self.a_mode = a_mode
        self._plans: Dict[int, dict] = {}          # crop-count bucket -> {"plan", "graphs"}: nothing here depends on the caller's frames
        self._wcache: dict = {}                    # packed (BN-folded) weights per layer, shared by every bucket
        self._cur = [0, 0, 0]                      # frames pointer, height, width of the call in flight (read by the ROI launch)
        self._frame_bufs: Dict[tuple, Tuple[torch.Tensor, torch.Tensor]] = {}     # (F, H, W) -> (pinned, device) staging, engine-owned
        self.max_graphs = 4                        # CUDA graphs kept per bucket (one per distinct frames pointer / shape), LRU
        self.stream = torch.cuda.Stream(device=self.device)
        self.rois = torch.zeros(capacity, 5, dtype=torch.int32, device=self.device)
        self.rois_host = torch.zeros(capacity, 5, dtype=torch.int32).pin_memory()
        self.features = torch.zeros(capacity, 512, dtype=torch.float32, device=self.device)
        self.features_host = torch.zeros(capacity, 512, dtype=torch.float32).pin_memory()
        # train-mode BatchNorm: segment tables (one segment = one reference Extractor call), filled per call
        self.max_segments = max_segments
        cap_b = self._bucket(capacity)
        self.seg_of_crop = torch.zeros(cap_b, dtype=torch.int32, device=self.device)
        self.seg_host = torch.zeros(cap_b, dtype=torch.int32).pin_memory()
        self.seg_crops = torch.zeros(max_segments + 1, dtype=torch.int32, device=self.device)
        self.seg_crops_host = torch.zeros(max_segments + 1, dtype=torch.int32).pin_memory()
        self.conv_flops_per_crop = 0.0

    def _bucket(self, n: int) -> int:
        b = 8
        while b < n:
            b *= 2
        return min(b, max(self.capacity, n))

    def _bn(self, p):
        sd = self.sd
        return (sd[p + ".weight"], sd[p + ".bias"], sd[p + ".running_mean"], sd[p + ".running_var"])

    # -- engine-owned frame staging -------------------------------------------------------------
    def stage_frames(self, frames_np: np.ndarray) -> torch.Tensor:
        """Host uint8 [F, H, W, 3] (or [H, W, 3]) -> device tensor in an engine-owned buffer keyed by shape (stable pointer, so
        the CUDA graph captured for it is reused); the copy is enqueued on the engine's stream, in order with run()."""
        if frames_np.ndim == 3:
            frames_np = frames_np[None]
        key = tuple(frames_np.shape[:3])
        if key not in self._frame_bufs:
            if len(self._frame_bufs) >= 4:                      # bounded: drop the oldest shape
                self._frame_bufs.pop(next(iter(self._frame_bufs)))
            self._frame_bufs[key] = (torch.empty(key + (3,), dtype=torch.uint8).pin_memory(),
                                     torch.empty(key + (3,), dtype=torch.uint8, device=self.device))
        pinned, dev = self._frame_bufs[key]
        self.stream.synchronize()                               # the previous upload from this pinned buffer has been consumed
        np.copyto(pinned.numpy(), frames_np)
        with torch.cuda.stream(self.stream):
            dev.copy_(pinned, non_blocking=True)
        return dev

    def stage_frame_list(self, frames: Sequence[np.ndarray]) -> torch.Tensor:
        """stage_frames for a list of equally sized HWC uint8 frames (gathered into the pinned buffer by a thread pool)"""
        from .hostcopy import upload_frames
        key = (len(frames),) + tuple(frames[0].shape[:2])
        if key not in self._frame_bufs:
            if len(self._frame_bufs) >= 4:
                self._frame_bufs.pop(next(iter(self._frame_bufs)))
            self._frame_bufs[key] = (torch.empty(key + (3,), dtype=torch.uint8).pin_memory(),
                                     torch.empty(key + (3,), dtype=torch.uint8, device=self.device))
        pinned, dev = self._frame_bufs[key]
        self.stream.synchronize()
        upload_frames(pinned, dev, frames, self.stream)         # gather of chunk i+1 overlaps the H2D of chunk i
        return dev

    def atlas(self, nbytes: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(pinned, device) byte buffers of at least `nbytes`, engine-owned, growing geometrically (Extractor.__call__'s crop atlas)"""
        cur = getattr(self, "_atlas", None)
        if cur is None or cur[0].numel() < nbytes:
            cap = 1 << max(20, int(nbytes - 1).bit_length())
            self.stream.synchronize()
            self._atlas = cur = (torch.empty(cap, dtype=torch.uint8).pin_memory(), torch.empty(cap, dtype=torch.uint8, device=self.device))
        return cur

    # -- eval-mode plan: one per crop-count bucket; the frames pointer / size are launch-time values ------------------
    def _build_eval(self, nb: int) -> dict:
        dev, sd = self.device, self.sd
        plan = _Plan(dev)
        plan.stream = self.stream
        wc = self._wcache

        def buf(h, c, dtype=torch.float16):
            return torch.zeros(nb, h, h, c, dtype=dtype, device=dev)

        rd = L.RoiDesc()
        rd.num_rois, rd.out_size, rd.out_channels = nb, REID_SIZE, 16
        for c in range(3):
            rd.mean[c] = REID_MEAN[c]; rd.inv_std[c] = 1.0 / REID_STD[c]
        plan.keep += [rd]
        cur_ = self._cur
        RELU = L.ACT_RELU
        if "stem" not in wc:
            wc["stem"] = _fold_plain(sd["conv.0.weight"], sd["conv.0.bias"], self._bn("conv.1"), REID_BN_EPS, dev)
        w_, b_ = wc["stem"]
        cur = TRef(buf(25, 64), 0, 64)
        stem_flops = 2.0 * nb * 2500 * 64 * 27
        if self.fused_stem and self.stem_direct:
            # frames + ROIs -> crop / resize / normalise -> conv + bias -> ReLU -> 3x3/s2 max-pool in ONE tcgen05 kernel; neither the
            # crop, nor the im2col operand, nor the 50x50x64 stem map reaches HBM (csrc/reid_stem_direct.cu)
            if "stem_packed_direct" not in wc:
                wc["stem_packed_direct"] = ops.pack_reid_stem_weights_direct(w_, b_)
            wpd = wc["stem_packed_direct"]
            plan.conv_flops += stem_flops
            plan.num_convs += 1
            plan.add(lambda st, cur=cur: ops.reid_stem_direct(rd, cur_[0], cur_[1], cur_[2], self.rois, wpd, cur.buf, stream=st),
                     f"crop+resize+norm + stem conv3x3 3->64 + maxpool3x3s2 (one kernel) M={nb * 2500}", stem_flops)
        elif self.fused_stem:
            # crop -> im2col patches of the stem (K = 27 -> 32), then ONE kernel: tcgen05 GEMM + bias + ReLU + 3x3/s2 max-pool
            # (csrc/reid_stem.cu): the 50x50x64 stem map never reaches HBM
            patches = torch.zeros(nb, 25, 128, 32, dtype=torch.float16, device=dev)
            if "stem_packed" not in wc:
                wc["stem_packed"] = ops.pack_reid_stem_weights(w_, b_)
            wp, bp = wc["stem_packed"]
            plan.keep += [patches]
            plan.add(lambda st: ops.roi_stem_patches(rd, cur_[0], cur_[1], cur_[2], self.rois, patches, stream=st),
                     "roi crop+resize+norm -> stem patches")
            plan.conv_flops += stem_flops
            plan.num_convs += 1
            plan.add(lambda st, cur=cur: ops.reid_stem_pool(patches, wp, bp, cur.buf, nb, stream=st),
                     f"stem conv3x3 3->64 + maxpool3x3s2 (fused) M={nb * 2500}", stem_flops)
        else:
            x0 = buf(REID_SIZE, 16)          # 3 real + 13 zero channels: one 32-byte TMA box per tap (bk = 16)
            plan.add(lambda st: ops.roi_resize_norm(rd, cur_[0], cur_[1], cur_[2], self.rois, x0, stream=st), "roi crop+resize+norm")
            s0 = buf(50, 64)
            plan.conv(TRef(x0, 0, 16), nb, _pad_cin(w_, 16), b_, TRef(s0, 0, 64), 3, 1, 1, RELU, a_mode=self.a_mode, flops=stem_flops,
                      wcache=wc, wkey="stem16")
            plan.add(lambda st, s0=s0, cur=cur: ops.maxpool(s0, 64, cur.buf, 64, nb, 50, 50, 64, 3, 2, 1, stream=st), "maxpool3x3s2")
        size = 25

        def folded(name, wname, bnname):
            if name not in wc:
                wc[name] = _fold_plain(sd[wname], None, self._bn(bnname), REID_BN_EPS, dev)
            return wc[name]

        for prefix, ci, co, down in REID_BLOCKS:
            s = 2 if down else 1
            osz = (size + 2 - 3) // s + 1
            w1, b1 = folded(prefix + ".f1", prefix + ".conv1.weight", prefix + ".bn1")
            t = TRef(buf(osz, co), 0, co)
            plan.conv(cur, nb, w1, b1, t, 3, s, 1, RELU, a_mode=self.a_mode, wcache=wc, wkey=prefix + ".p1")
            if down:
                wd, bd = folded(prefix + ".fd", prefix + ".downsample.0.weight", prefix + ".downsample.1")
                sc = TRef(buf(osz, co), 0, co)
                plan.conv(cur, nb, wd, bd, sc, 1, 2, 0, L.ACT_NONE, a_mode=self.a_mode, wcache=wc, wkey=prefix + ".pd")
            else:
                sc = cur
            w2, b2 = folded(prefix + ".f2", prefix + ".conv2.weight", prefix + ".bn2")
            y = TRef(buf(osz, co), 0, co)
            plan.conv(t, nb, w2, b2, y, 3, 1, 1, RELU, residual=sc, res_mode=L.RES_BEFORE_ACT, a_mode=self.a_mode, wcache=wc, wkey=prefix + ".p2")
            cur, size = y, osz
        assert size == 4
        feats = self.features
        plan.add(lambda st, cur=cur: ops.avgpool_l2norm(cur.buf, 512, nb, 16, 512, feats, stream=st), "avgpool+l2norm")
        self.conv_flops_per_crop = plan.conv_flops / nb
        from collections import OrderedDict
        return {"plan": plan, "graphs": OrderedDict(), "rd": rd}

    # -- train-mode (reference-faithful) plan: conv -> segment statistics -> normalise, graph-captured per bucket ------------
    def _build_train(self, nb: int) -> dict:
        """BatchNorm with the statistics of each reference call (segment).  Every BN layer is three launches over static
        buffers: the convolution writes the pre-BN tensor as fp16, vcb_bn_seg_stats_f16 accumulates per-(segment, channel)
        sums, vcb_bn_seg_apply_f16 normalises (+ residual, ReLU; the stem's also max-pools).  The segment tables are device
        buffers filled per call, so one CUDA graph per bucket serves every segmentation."""
        dev, sd = self.device, self.sd
        plan = _Plan(dev)
        plan.stream = self.stream
        wc = self._wcache
        S = self.max_segments + 1

        def buf(h, c):
            return torch.zeros(nb, h, h, c, dtype=torch.float16, device=dev)

        def f32(name):
            key = "t:" + name
            if key not in wc:
                wc[key] = sd[name].to(device=dev, dtype=torch.float32).contiguous()
            return wc[key]

        rd = L.RoiDesc()
        rd.num_rois, rd.out_size, rd.out_channels = nb, REID_SIZE, 16
        for c in range(3):
            rd.mean[c] = REID_MEAN[c]; rd.inv_std[c] = 1.0 / REID_STD[c]
        plan.keep += [rd]
        cur_ = self._cur
        n_bn = 1 + sum(3 if down else 2 for _, _, _, down in REID_BLOCKS)
        sums = torch.zeros(n_bn, S, 512, 2, dtype=torch.float64, device=dev)
        plan.keep += [sums]

        def zero_sums(st):
            with torch.cuda.stream(st):
                sums.zero_()
        plan.add(zero_sums, "zero BN sums")
        bn_i = [0]

        fuse_apply = os.environ.get("VCB_BN_FUSED_APPLY", "1") != "0"

        def conv_bn(x: TRef, wname, bias_name, bnp, co, k, s, p, act, residual: Optional[TRef], pool=False, flops=None, raw_only=False,
                    res_bn=None):
            """conv -> statistics -> normalise (+ residual) -> act.  `raw_only`: stop after the statistics and return (raw, sums, gamma,
            beta) -- the caller's next conv_bn normalises this tensor as its residual (`res_bn`), so the downsample branch of a block
            costs no apply pass of its own.  $VCB_BN_FUSED_APPLY=0: finalize table + vcb_bn_seg_apply_f16 per BatchNorm (round-2 form)."""
            w = sd[wname].to(device=dev, dtype=torch.float32)
            if w.shape[1] != x.c:
                w = _pad_cin(w, x.c)
            b = None if bias_name is None else sd[bias_name].to(device=dev, dtype=torch.float32)
            osz = (x.h + 2 * p - k) // s + 1
            raw = TRef(buf(osz, co), 0, co)
            sl = sums[bn_i[0]]
            bn_i[0] += 1
            g, be = f32(bnp + ".weight"), f32(bnp + ".bias")
            fused = plan.conv(x, nb, w, b, raw, k, s, p, L.ACT_NONE, a_mode=self.a_mode, wcache=wc, wkey="t:" + wname, flops=flops,
                              stats=(self.seg_of_crop, sl) if self.epi_stats else None)
            if not (self.epi_stats and fused):          # statistics from the convolution's epilogue, else one more pass over the tensor
                plan.add(lambda st, raw=raw, sl=sl: ops.bn_seg_stats_f16(raw.buf, co, osz * osz, nb, self.seg_of_crop, sl, stream=st), f"bn stats c={co}")
            if raw_only:
                return raw, sl, g, be
            out_sz = (osz + 1) // 2 if pool else osz
            y = TRef(buf(out_sz, co), 0, co)
            if fuse_apply and not pool:
                rs, rg, rb = (res_bn[1], res_bn[2], res_bn[3]) if res_bn is not None else (None, None, None)
                res = res_bn[0] if res_bn is not None else residual
                plan.keep += [t for t in (rs, rg, rb) if t is not None]
                plan.add(lambda st, raw=raw, sl=sl, y=y, res=res, rs=rs, rg=rg, rb=rb: ops.bn_seg_apply_fused_f16(
                    raw.buf, co, osz * osz, nb, self.seg_of_crop, self.seg_crops, sl, g, be, REID_BN_EPS,
                    None if res is None else res.ptr, 0 if res is None else res.pitch, act, y.buf, co,
                    res_sums=rs, res_gamma=rg, res_beta=rb, stream=st),
                    f"bn apply c={co}" + (" (+BN of the residual)" if res_bn is not None else ""))
                return y
            assert res_bn is None
            aff = torch.zeros(S, co, 2, dtype=torch.float32, device=dev)
            plan.keep += [aff]
            plan.add(lambda st, sl=sl, aff=aff: ops.bn_seg_finalize(sl, self.seg_crops, S, co, osz * osz, g, be, None, REID_BN_EPS, aff, stream=st),
                     "bn finalize")
            plan.add(lambda st, raw=raw, aff=aff, y=y: ops.bn_seg_apply_f16(
                raw.buf, co, osz, osz, nb, self.seg_of_crop, aff, None if residual is None else residual.ptr,
                0 if residual is None else residual.pitch, act, 1 if pool else 0, y.buf, co, stream=st),
                f"bn apply c={co}" + (" +pool" if pool else ""))
            return y

        stem_flops = 2.0 * nb * 2500 * 64 * 27
        if self.fused_stem and self.stem_direct:
            # the direct stem kernel twice (csrc/reid_stem_direct.cu MODE 1 / MODE 2): statistics of conv + bias per segment, then
            # conv -> per-segment scale / shift -> ReLU -> max-pool; both passes resize the crop in shared memory from the frames
            if "t:stem_packed_direct" not in wc:
                wc["t:stem_packed_direct"] = ops.pack_reid_stem_weights_direct(sd["conv.0.weight"].to(device=dev, dtype=torch.float32),
                                                                               sd["conv.0.bias"].to(device=dev, dtype=torch.float32))
            wpd = wc["t:stem_packed_direct"]
            affine = torch.zeros(S, 64, 2, dtype=torch.float32, device=dev)
            g0, b0 = f32("conv.1.weight"), f32("conv.1.bias")
            cur = TRef(buf(25, 64), 0, 64)
            sl0 = sums[bn_i[0]]
            bn_i[0] += 1
            plan.keep += [affine]
            plan.add(lambda st: ops.reid_stem_direct_stats(rd, cur_[0], cur_[1], cur_[2], self.rois, wpd, self.seg_of_crop, sl0, stream=st),
                     "crop+resize+norm + stem conv: BN statistics only")
            plan.add(lambda st: ops.bn_seg_finalize(sl0, self.seg_crops, S, 64, 2500, g0, b0, None, REID_BN_EPS, affine, stream=st), "bn finalize")
            plan.conv_flops += stem_flops
            plan.num_convs += 1
            plan.add(lambda st, cur=cur: ops.reid_stem_direct_bn(rd, cur_[0], cur_[1], cur_[2], self.rois, wpd, affine, self.seg_of_crop, cur.buf,
                                                                 stream=st),
                     f"crop+resize+norm + stem conv3x3 3->64 + BN(train) + ReLU + maxpool3x3s2 (one kernel) M={nb * 2500}", stem_flops)
        elif self.fused_stem:
            # the fused stem kernel twice over the same im2col patches (csrc/reid_stem.cu MODE 1 / MODE 2): statistics of conv + bias
            # per segment, then conv -> per-segment scale / shift -> ReLU -> max-pool.  The 50x50x64 pre-BN map (1.3 GB per 4096
            # crops, written once and read twice by the three-launch form below) never exists.
            patches = torch.zeros(nb, 25, 128, 32, dtype=torch.float16, device=dev)
            if "t:stem_packed" not in wc:
                wc["t:stem_packed"] = ops.pack_reid_stem_weights(sd["conv.0.weight"].to(device=dev, dtype=torch.float32),
                                                                 sd["conv.0.bias"].to(device=dev, dtype=torch.float32))
            wp, bp = wc["t:stem_packed"]
            affine = torch.zeros(S, 64, 2, dtype=torch.float32, device=dev)
            g0, b0 = f32("conv.1.weight"), f32("conv.1.bias")
            cur = TRef(buf(25, 64), 0, 64)
            sl0 = sums[bn_i[0]]
            bn_i[0] += 1
            plan.keep += [patches, affine]
            plan.add(lambda st: ops.roi_stem_patches(rd, cur_[0], cur_[1], cur_[2], self.rois, patches, stream=st),
                     "roi crop+resize+norm -> stem patches")
            plan.add(lambda st: ops.reid_stem_stats(patches, wp, bp, nb, self.seg_of_crop, sl0, stream=st), "stem conv: BN statistics only")
            plan.add(lambda st: ops.bn_seg_finalize(sl0, self.seg_crops, S, 64, 2500, g0, b0, bp, REID_BN_EPS, affine, stream=st), "bn finalize")
            plan.conv_flops += stem_flops
            plan.num_convs += 1
            plan.add(lambda st, cur=cur: ops.reid_stem_pool_bn(patches, wp, affine, self.seg_of_crop, cur.buf, nb, stream=st),
                     f"stem conv3x3 3->64 + BN(train) + ReLU + maxpool3x3s2 (fused) M={nb * 2500}", stem_flops)
        else:
            x0 = buf(REID_SIZE, 16)
            plan.add(lambda st: ops.roi_resize_norm(rd, cur_[0], cur_[1], cur_[2], self.rois, x0, stream=st), "roi crop+resize+norm")
            cur = conv_bn(TRef(x0, 0, 16), "conv.0.weight", "conv.0.bias", "conv.1", 64, 3, 1, 1, L.ACT_RELU, None, pool=True,
                          flops=stem_flops)
        for prefix, ci, co, down in REID_BLOCKS:
            s_ = 2 if down else 1
            t = conv_bn(cur, prefix + ".conv1.weight", None, prefix + ".bn1", co, 3, s_, 1, L.ACT_RELU, None)
            if down and fuse_apply:       # the downsample branch stays pre-BN; the block's last apply normalises it as its residual
                rb = conv_bn(cur, prefix + ".downsample.0.weight", None, prefix + ".downsample.1", co, 1, 2, 0, L.ACT_NONE, None, raw_only=True)
                cur = conv_bn(t, prefix + ".conv2.weight", None, prefix + ".bn2", co, 3, 1, 1, L.ACT_RELU, None, res_bn=rb)
                continue
            sc = conv_bn(cur, prefix + ".downsample.0.weight", None, prefix + ".downsample.1", co, 1, 2, 0, L.ACT_NONE, None) if down else cur
            cur = conv_bn(t, prefix + ".conv2.weight", None, prefix + ".bn2", co, 3, 1, 1, L.ACT_RELU, sc)
        assert cur.h == 4
        feats = self.features
        plan.add(lambda st, cur=cur: ops.avgpool_l2norm(cur.buf, 512, nb, 16, 512, feats, stream=st), "avgpool+l2norm")
        self.conv_flops_per_crop = plan.conv_flops / nb
        from collections import OrderedDict
        return {"plan": plan, "graphs": OrderedDict(), "rd": rd}

    def _set_segments(self, n: int, nb: int, seg_sizes: Sequence[int]) -> None:
        nseg = len(seg_sizes)
        if nseg > self.max_segments:
            raise ValueError(f"{nseg} BatchNorm segments in one call; this engine was built for at most {self.max_segments}")
        assert sum(seg_sizes) == n, (seg_sizes, n)
        soc = self.seg_host.numpy()
        soc[:n] = np.repeat(np.arange(nseg, dtype=np.int32), np.asarray(seg_sizes, np.int64))
        soc[n:nb] = nseg                                  # padding crops of the bucket: their own (ignored) segment
        cnt = self.seg_crops_host.numpy()
        cnt[:] = 0
        cnt[:nseg] = np.asarray(seg_sizes, np.int32)
        cnt[nseg] = nb - n
        with torch.cuda.stream(self.stream):
            self.seg_of_crop.copy_(self.seg_host, non_blocking=True)
            self.seg_crops.copy_(self.seg_crops_host, non_blocking=True)

    # -- per-call API --------------------------------------------------------------------------
    def run(self, frames: torch.Tensor, rois, n: Optional[int] = None, seg_sizes: Optional[Sequence[int]] = None,
            use_graph: bool = True) -> torch.Tensor:
        """frames: device uint8 [F, H, W, 3]; rois: host int32 [n,5] (frame, x1, y1, x2, y2) or None when
        self.rois was filled on the device.  Returns the device tensor features[:n] (async on self.stream)."""
        if rois is not None:
            rois = np.ascontiguousarray(rois, dtype=np.int32).reshape(-1, 5)
            n = rois.shape[0]
            assert n <= self.capacity, (n, self.capacity)
            self.rois_host.zero_()
            self.rois_host[:n] = torch.from_numpy(rois)
            with torch.cuda.stream(self.stream):
                self.rois.copy_(self.rois_host, non_blocking=True)
        assert n is not None
        if n == 0:
            return self.features[:0]
        nb = self._bucket(n)
        if self.bn_mode == "train":
            segs = tuple(int(k) for k in seg_sizes) if seg_sizes is not None else (n,)
            if (n, nb, segs) != getattr(self, "_seg_key", None):      # same segmentation as the previous call: tables already on the device
                self.stream.synchronize()      # the pinned segment tables of the previous call have been consumed
                self._set_segments(n, nb, list(segs))
                self._seg_key = (n, nb, segs)
        ent = self._plans.get(nb)
        if ent is None:
            ent = self._plans[nb] = self._build_train(nb) if self.bn_mode == "train" else self._build_eval(nb)
        plan = ent["plan"]
        self._cur[0], self._cur[1], self._cur[2] = frames.data_ptr(), int(frames.shape[1]), int(frames.shape[2])
        ent["rd"].num_frames = int(frames.shape[0])          # bounds for the ROI kernels' frame index (part of the graph key below)
        if not use_graph:
            plan.run_eager(self.stream)
            return self.features[:n]
        # graphs bake the frames pointer and size: one per (pointer, shape), least recently used dropped beyond max_graphs.
        # Callers whose frames live in a fresh tensor every call (Extractor.__call__) pass use_graph=False instead.
        gkey = (self._cur[0], self._cur[1], self._cur[2], int(frames.shape[0]))
        g = ent["graphs"].get(gkey)
        if g is None:
            g = plan.capture()
            ent["graphs"][gkey] = g
            while len(ent["graphs"]) > self.max_graphs:
                ent["graphs"].popitem(last=False)
        else:
            ent["graphs"].move_to_end(gkey)
            plan.graph = g
        g.launch(self.stream)
        return self.features[:n]

    def download(self, n: int) -> np.ndarray:
        with torch.cuda.stream(self.stream):
            self.features_host[:n].copy_(self.features[:n], non_blocking=True)
        self.stream.synchronize()
        return self.features_host[:n].numpy().copy()
