"""Tensor-level wrappers over the C ABI (one function per entry point of include/vcb200.h).

Tensors are torch CUDA tensors used purely as device buffers; every function enqueues on the
given (or current) CUDA stream and returns immediately.  NHWC fp16 activations carry an explicit
channel pitch = size of the last dimension of the *buffer* they live in, so a producer can write a
channel slice of a wider concat buffer.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L


def _st(stream) -> int:
    return L.stream_handle(stream)


# ------------------------------------------------------------------------------------------ conv
def make_conv_desc(n, h, w, cin, cout, k, stride, pad, *, cin_pitch=None, cout_pitch=None, act=L.ACT_SILU,
                   res_mode=L.RES_NONE, res_pitch=0, out_dtype=L.F16, a_mode=L.A_AUTO, block_n=0, stages=0,
                   kw=None, epi_direct=False, c4_narrow=False, bk=0, cta_pair=0, a_im2col=False, one_chain=False, dbg1=0,
                   tile_rev=False) -> L.ConvDesc:
    d = L.ConvDesc()
    d.n, d.h, d.w = n, h, w
    d.cin, d.cin_pitch = cin, cin if cin_pitch is None else cin_pitch
    d.cout = cout
    d.cout_pitch = ((cout + 7) // 8 * 8) if cout_pitch is None else cout_pitch
    d.kh, d.kw = k, (k if kw is None else kw)
    d.stride, d.pad = stride, pad
    d.act, d.res_mode, d.res_pitch = act, res_mode, res_pitch
    d.out_dtype, d.a_mode, d.block_n, d.stages = out_dtype, a_mode, block_n, stages
    d.reserved[0] = int(epi_direct)   # debug: 1 = direct global stores; 2 = second TMA producer for B; 3 = skip epilogue; 4 = one accumulator
    d.reserved[1] = dbg1 if dbg1 else 3 if one_chain else (2 if a_im2col else (1 if c4_narrow else 0))   # debug: 1 = 8-byte C4 gather; 2 = im2col TMA even for 1x1; 3 = one accumulation chain
    if tile_rev:
        d.reserved[1] |= 0x100                  # per-launch flag: walk the tiles from the last to the first (L2 reuse between layers)
    d.reserved[2] = bk                          # 0 = auto; 16/32/64 forces the K chunk width of the TMA path
    d.reserved[3] = cta_pair                    # 0 = auto; 1 = single-CTA 128-row kernel; 2 = CTA-pair (cta_group::2) kernel; 3 = 256-row tiles
    return d


def conv_out_hw(d: L.ConvDesc) -> Tuple[int, int]:
    ho, wo = C.c_int32(), C.c_int32()
    L.check(L.load().vcb_conv_out_hw(C.byref(d), C.byref(ho), C.byref(wo)), "vcb_conv_out_hw")
    return ho.value, wo.value


def conv_packed_sizes(d: L.ConvDesc) -> Tuple[int, int]:
    a, b = C.c_int64(), C.c_int64()
    L.check(L.load().vcb_conv_packed_sizes(C.byref(d), C.byref(a), C.byref(b)), "vcb_conv_packed_sizes")
    return a.value, b.value


def pack_conv_weights(d: L.ConvDesc, w_oihw: torch.Tensor, bias: Optional[torch.Tensor], stream=None):
    """fp32 OIHW (BN folded) + bias on the device -> (packed fp16 blob, padded fp32 bias)."""
    lib = L.init(w_oihw.device.index or 0)
    nw, nb = conv_packed_sizes(d)
    w_oihw = w_oihw.contiguous().float()
    assert tuple(w_oihw.shape) == (d.cout, d.cin, d.kh, d.kw), (tuple(w_oihw.shape), (d.cout, d.cin, d.kh, d.kw))
    wp = torch.empty(nw, dtype=torch.float16, device=w_oihw.device)
    bp = torch.empty(nb, dtype=torch.float32, device=w_oihw.device)
    if bias is not None:
        bias = bias.contiguous().float()
    L.check(lib.vcb_conv_pack_weights(C.byref(d), L.ptr(w_oihw), L.ptr(bias), L.ptr(wp), L.ptr(bp), _st(stream)),
            "vcb_conv_pack_weights")
    return wp, bp


def conv2d(d: L.ConvDesc, x, wp, bp, y, residual=None, stream=None) -> None:
    lib = L.load()
    L.check(lib.vcb_conv2d_fwd(C.byref(d), L.ptr(x), L.ptr(wp), L.ptr(bp), L.ptr(residual), L.ptr(y), _st(stream)),
            "vcb_conv2d_fwd")


def conv2d_stats(d: L.ConvDesc, x, wp, bp, y, seg_of_image, sums, stream=None) -> None:
    """conv2d (no activation, fp16 out) + per-(segment, channel) sum / sum of squares of its outputs from the epilogue"""
    L.check(L.load().vcb_conv2d_fwd_stats(C.byref(d), L.ptr(x), L.ptr(wp), L.ptr(bp), L.ptr(y), L.ptr(seg_of_image), L.ptr(sums), _st(stream)),
            "vcb_conv2d_fwd_stats")


# ------------------------------------------------------------------------------------------ data movement
def frames_to_f16c4(frames_u8: torch.Tensor, out: torch.Tensor, stream=None) -> None:
    n, h, w, c = frames_u8.shape
    assert c == 3 and frames_u8.dtype == torch.uint8 and frames_u8.is_contiguous()
    L.check(L.load().vcb_frames_to_f16c4(L.ptr(frames_u8), L.ptr(out), n, h, w, _st(stream)), "vcb_frames_to_f16c4")


def frames_to_f16_s2d(frames_u8: torch.Tensor, out: torch.Tensor, stream=None) -> None:
    n, h, w, c = frames_u8.shape
    assert c == 3 and frames_u8.dtype == torch.uint8 and frames_u8.is_contiguous()
    L.check(L.load().vcb_frames_to_f16_s2d(L.ptr(frames_u8), L.ptr(out), n, h, w, _st(stream)), "vcb_frames_to_f16_s2d")


def frames_to_f16_s2d_wpad(frames_u8: torch.Tensor, out: torch.Tensor, stream=None) -> None:
    """out: fp16 buffer of n*(h/2)*(w/2+2)*16 + 16 elements, zeroed once by the caller (pad columns stay zero)"""
    n, h, w, c = frames_u8.shape
    assert c == 3 and frames_u8.dtype == torch.uint8 and frames_u8.is_contiguous()
    assert out.numel() >= n * (h // 2) * (w // 2 + 2) * 16 + 16
    L.check(L.load().vcb_frames_to_f16_s2d_wpad(L.ptr(frames_u8), L.ptr(out), n, h, w, _st(stream)), "vcb_frames_to_f16_s2d_wpad")


def upsample2x(src, src_pitch, dst, dst_pitch, n, h, w, c, stream=None) -> None:
    L.check(L.load().vcb_upsample2x(L.ptr(src), src_pitch, L.ptr(dst), dst_pitch, n, h, w, c, _st(stream)),
            "vcb_upsample2x")


def sppf_pool(buf, pitch, n, h, w, c, stream=None) -> None:
    L.check(L.load().vcb_sppf_pool(L.ptr(buf), pitch, n, h, w, c, _st(stream)), "vcb_sppf_pool")


def maxpool(src, src_pitch, dst, dst_pitch, n, h, w, c, k, s, p, stream=None) -> None:
    L.check(L.load().vcb_maxpool(L.ptr(src), src_pitch, L.ptr(dst), dst_pitch, n, h, w, c, k, s, p, _st(stream)),
            "vcb_maxpool")


def avgpool_l2norm(x, pitch, n, hw, c, out, stream=None) -> None:
    L.check(L.load().vcb_avgpool_l2norm(L.ptr(x), pitch, n, hw, c, L.ptr(out), _st(stream)), "vcb_avgpool_l2norm")


def bn_train_stats(x, c, seg_row_start, num_seg, gamma, beta, eps, scale, shift, stream=None) -> None:
    L.check(L.load().vcb_bn_train_stats(L.ptr(x), c, L.ptr(seg_row_start), num_seg, L.ptr(gamma), L.ptr(beta),
                                        eps, L.ptr(scale), L.ptr(shift), _st(stream)), "vcb_bn_train_stats")


def bn_apply(x, c, rows, row_seg, scale, shift, residual, res_pitch, act, y, y_pitch, stream=None) -> None:
    L.check(L.load().vcb_bn_apply(L.ptr(x), c, rows, L.ptr(row_seg), L.ptr(scale), L.ptr(shift),
                                  L.ptr(residual), res_pitch, act, L.ptr(y), y_pitch, _st(stream)), "vcb_bn_apply")


def reid_stem_stats(patches, w_packed, bias, num_rois, seg_of_crop, sums, stream=None) -> None:
    L.check(L.load().vcb_reid_stem_stats(L.ptr(patches), L.ptr(w_packed), L.ptr(bias), num_rois, L.ptr(seg_of_crop), L.ptr(sums), _st(stream)),
            "vcb_reid_stem_stats")


def bn_seg_finalize(sums, seg_crops, num_seg_plus1, c, hw, gamma, beta, bias, eps, affine, stream=None) -> None:
    L.check(L.load().vcb_bn_seg_finalize(L.ptr(sums), L.ptr(seg_crops), num_seg_plus1, c, hw, L.ptr(gamma), L.ptr(beta), L.ptr(bias), eps,
                                         L.ptr(affine), _st(stream)), "vcb_bn_seg_finalize")


def bn_seg_apply_fused_f16(x, c, hw, n, seg_of_crop, seg_crops, sums, gamma, beta, eps, residual, res_pitch, act, y, y_pitch, res_sums=None,
                           res_gamma=None, res_beta=None, stream=None) -> None:
    """finalize + apply in one launch; `res_sums` / `res_gamma` / `res_beta`: the residual is a pre-BN tensor normalised on the fly"""
    L.check(L.load().vcb_bn_seg_apply_fused_f16(L.ptr(x), c, hw, n, L.ptr(seg_of_crop), L.ptr(seg_crops), L.ptr(sums), L.ptr(gamma), L.ptr(beta),
                                                eps, L.ptr(residual), res_pitch, L.ptr(res_sums), L.ptr(res_gamma), L.ptr(res_beta), act,
                                                L.ptr(y), y_pitch, _st(stream)), "vcb_bn_seg_apply_fused_f16")


def reid_stem_pool_bn(patches, w_packed, affine, seg_of_crop, out, num_rois, stream=None) -> None:
    L.check(L.load().vcb_reid_stem_pool_bn(L.ptr(patches), L.ptr(w_packed), L.ptr(affine), L.ptr(seg_of_crop), L.ptr(out), num_rois,
                                           _st(stream)), "vcb_reid_stem_pool_bn")


def bn_seg_stats_f16(x, c, hw, n, seg_of_crop, sums, stream=None) -> None:
    L.check(L.load().vcb_bn_seg_stats_f16(L.ptr(x), c, hw, n, L.ptr(seg_of_crop), L.ptr(sums), _st(stream)), "vcb_bn_seg_stats_f16")


def bn_seg_apply_f16(x, c, h, w, n, seg_of_crop, affine, residual, res_pitch, act, pool, y, y_pitch, stream=None) -> None:
    L.check(L.load().vcb_bn_seg_apply_f16(L.ptr(x), c, h, w, n, L.ptr(seg_of_crop), L.ptr(affine), L.ptr(residual), res_pitch, act, pool,
                                          L.ptr(y), y_pitch, _st(stream)), "vcb_bn_seg_apply_f16")


# ------------------------------------------------------------------------------------------ detect / nms
def detect_decode(desc: L.DetectDesc, cand_box, cand_score, cand_cls, cand_index, cand_count, stream=None) -> None:
    L.check(L.load().vcb_detect_decode(C.byref(desc), L.ptr(cand_box), L.ptr(cand_score), L.ptr(cand_cls), L.ptr(cand_index),
                                       L.ptr(cand_count), _st(stream)), "vcb_detect_decode")


def nms_workspace_bytes(n: int, max_candidates: int) -> int:
    return int(L.load().vcb_nms_workspace_bytes(n, max_candidates))


def nms(desc: L.NmsDesc, cand_box, cand_score, cand_cls, cand_index, cand_count, ws, det, det_count, stream=None) -> None:
    L.check(L.load().vcb_nms(C.byref(desc), L.ptr(cand_box), L.ptr(cand_score), L.ptr(cand_cls), L.ptr(cand_index),
                             L.ptr(cand_count), L.ptr(ws), L.ptr(det), L.ptr(det_count), _st(stream)), "vcb_nms")


# ------------------------------------------------------------------------------------------ ROI
def roi_resize_norm(desc: L.RoiDesc, frames_u8, fh, fw, rois, out, stream=None) -> None:
    L.check(L.load().vcb_roi_resize_norm(C.byref(desc), L.ptr(frames_u8), fh, fw, L.ptr(rois), L.ptr(out), _st(stream)),
            "vcb_roi_resize_norm")


def letterbox_half(src_u8, n, h0, w0, dst_u8, h1, w1, top, left, pad=114, stream=None) -> None:
    """exact-2x letterbox on the device (bit-identical to cv2.resize INTER_LINEAR + copyMakeBorder for this ratio)"""
    L.check(L.load().vcb_letterbox_half_u8(L.ptr(src_u8), n, h0, w0, L.ptr(dst_u8), h1, w1, top, left, pad, _st(stream)),
            "vcb_letterbox_half_u8")


def letterbox_bilinear(src_u8, n, h0, w0, dst_u8, h1, w1, top, left, new_h, new_w, xtab, ytab, pad=114, stream=None) -> None:
    """any-ratio letterbox on the device, bit-identical to cv2.resize(INTER_LINEAR) + copyMakeBorder (tables: cv2_linear_table)"""
    L.check(L.load().vcb_letterbox_bilinear_u8(L.ptr(src_u8), n, h0, w0, L.ptr(dst_u8), h1, w1, top, left, new_h, new_w, L.ptr(xtab),
                                               L.ptr(ytab), pad, _st(stream)), "vcb_letterbox_bilinear_u8")


def roi_stem_patches(desc: L.RoiDesc, frames_u8, fh, fw, rois, patches, stream=None) -> None:
    """crop/resize/normalise as roi_resize_norm, written as the fused stem's im2col operand [n][25][128][32] fp16"""
    L.check(L.load().vcb_roi_stem_patches(C.byref(desc), L.ptr(frames_u8), fh, fw, L.ptr(rois), L.ptr(patches), _st(stream)),
            "vcb_roi_stem_patches")


def reid_stem_pool(patches, w_packed, bias, out, num_rois, stream=None) -> None:
    """maxpool3x3s2p1(relu(conv3x3(crop) + bias)) -> out fp16 [n][25][25][64] (tcgen05 GEMM + pooling in shared memory)"""
    L.check(L.load().vcb_reid_stem_pool(L.ptr(patches), L.ptr(w_packed), L.ptr(bias), L.ptr(out), num_rois, _st(stream)),
            "vcb_reid_stem_pool")


def pack_reid_stem_weights(w_oihw: torch.Tensor, bias: torch.Tensor):
    """[64, 3, 3, 3] fp32 (BN folded) -> fp16 [64][32], k = (r*3+s)*3 + c, zero padded; bias fp32 [64]"""
    assert tuple(w_oihw.shape) == (64, 3, 3, 3)
    wk = w_oihw.permute(0, 2, 3, 1).reshape(64, 27)
    wp = torch.zeros(64, 32, dtype=torch.float16, device=w_oihw.device)
    wp[:, :27] = wk.to(torch.float16)
    return wp.contiguous(), bias.to(torch.float32).contiguous()


def pack_reid_stem_weights_direct(w_oihw: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """[64, 3, 3, 3] fp32 + bias -> fp16 [64][32] for the vcb_reid_stem_direct* kernels: k = (r*3+s)*3 + c, then the bias as an fp16
    head (k = 27) and tail (k = 28; head + tail reproduces the fp32 bias to ~2^-22), zeros"""
    assert tuple(w_oihw.shape) == (64, 3, 3, 3)
    wp = torch.zeros(64, 32, dtype=torch.float16, device=w_oihw.device)
    wp[:, :27] = w_oihw.permute(0, 2, 3, 1).reshape(64, 27).to(torch.float16)
    b32 = bias.to(device=w_oihw.device, dtype=torch.float32)
    hi = b32.to(torch.float16)
    wp[:, 27] = hi
    wp[:, 28] = (b32 - hi.float()).to(torch.float16)
    return wp.contiguous()


def reid_stem_direct(desc: L.RoiDesc, frames_u8, fh, fw, rois, w_packed, out, stream=None) -> None:
    """frames + ROIs -> crop/resize/normalise -> conv3x3(3->64) + bias -> ReLU -> maxpool 3/2/1, one kernel, nothing staged in HBM"""
    L.check(L.load().vcb_reid_stem_direct(C.byref(desc), L.ptr(frames_u8), fh, fw, L.ptr(rois), L.ptr(w_packed), L.ptr(out), _st(stream)),
            "vcb_reid_stem_direct")


def reid_stem_direct_stats(desc: L.RoiDesc, frames_u8, fh, fw, rois, w_packed, seg_of_crop, sums, stream=None) -> None:
    L.check(L.load().vcb_reid_stem_direct_stats(C.byref(desc), L.ptr(frames_u8), fh, fw, L.ptr(rois), L.ptr(w_packed), L.ptr(seg_of_crop),
                                                L.ptr(sums), _st(stream)), "vcb_reid_stem_direct_stats")


def reid_stem_direct_bn(desc: L.RoiDesc, frames_u8, fh, fw, rois, w_packed, affine, seg_of_crop, out, stream=None) -> None:
    L.check(L.load().vcb_reid_stem_direct_bn(C.byref(desc), L.ptr(frames_u8), fh, fw, L.ptr(rois), L.ptr(w_packed), L.ptr(affine),
                                             L.ptr(seg_of_crop), L.ptr(out), _st(stream)), "vcb_reid_stem_direct_bn")


def boxes_to_rois(boxes_f64, frame_of, num, fw, fh, rois, stream=None) -> None:
    L.check(L.load().vcb_boxes_to_rois(L.ptr(boxes_f64), L.ptr(frame_of), num, fw, fh, L.ptr(rois), _st(stream)),
            "vcb_boxes_to_rois")


# ------------------------------------------------------------------------------------------ graphs
class Graph:
    """A captured CUDA graph of library calls (vcb_graph_*)."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)

    @property
    def num_kernels(self) -> int:
        return int(L.load().vcb_graph_num_kernels(self._h))

    def launch(self, stream=None) -> None:
        L.check(L.load().vcb_graph_launch(self._h, _st(stream)), "vcb_graph_launch")

    def __del__(self):
        try:
            if self._h:
                L.load().vcb_graph_destroy(self._h)
                self._h = None
        except Exception:
            pass


def graph_begin(stream) -> None:
    L.check(L.load().vcb_graph_begin(_st(stream)), "vcb_graph_begin")


def graph_end(stream) -> Graph:
    h = C.c_void_p()
    L.check(L.load().vcb_graph_end(_st(stream), C.byref(h)), "vcb_graph_end")
    return Graph(h.value)
