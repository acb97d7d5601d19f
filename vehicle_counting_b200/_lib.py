"""ctypes binding of libvcb200.so (include/vcb200.h).

The product path has no CPU or eager-PyTorch fallback: if the shared library is missing or the
device is not sm_100, every entry point raises.  PyTorch is used only for device memory and
streams; all arithmetic on the hot path happens inside the library's CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# $VCB_LIB_PATH: an alternative build of the same library (A/B builds made by tools/ab_build.sh)
LIB_PATH = os.environ.get("VCB_LIB_PATH") or os.path.join(_HERE, "libvcb200.so")

VCB_OK = 0
ACT_NONE, ACT_SILU, ACT_RELU, ACT_SILU_TANH = 0, 1, 2, 3
RES_NONE, RES_AFTER_ACT, RES_BEFORE_ACT = 0, 1, 2
F16, F32 = 0, 1
A_AUTO, A_IM2COL_TMA, A_GATHER, A_C4, A_ROWWIN = 0, 1, 2, 3, 4


class VcbError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("cin", C.c_int32), ("cin_pitch", C.c_int32),
        ("cout", C.c_int32), ("cout_pitch", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("act", C.c_int32), ("res_mode", C.c_int32), ("res_pitch", C.c_int32),
        ("out_dtype", C.c_int32), ("a_mode", C.c_int32),
        ("block_n", C.c_int32), ("stages", C.c_int32),
        ("reserved", C.c_int32 * 4),
    ]


class DetectLevel(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("pitch", C.c_int32), ("ny", C.c_int32), ("nx", C.c_int32),
        ("stride", C.c_float), ("anchor_w", C.c_float * 3), ("anchor_h", C.c_float * 3),
    ]


class DetectDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("nc", C.c_int32), ("num_levels", C.c_int32), ("logits_dtype", C.c_int32),
        ("conf_thres", C.c_float), ("max_candidates", C.c_int32),
        ("level", DetectLevel * 4),
        ("use_class_mask", C.c_int32), ("class_mask", C.c_uint32 * 8),
    ]


class NmsDesc(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("max_candidates", C.c_int32), ("max_det", C.c_int32),
        ("iou_thres", C.c_float), ("max_wh", C.c_float), ("max_nms", C.c_int32),
        ("gain", C.c_void_p), ("pad_x", C.c_void_p), ("pad_y", C.c_void_p), ("w0", C.c_void_p), ("h0", C.c_void_p),
    ]


class RoiDesc(C.Structure):
    _fields_ = [
        ("num_rois", C.c_int32), ("out_size", C.c_int32),
        ("mean", C.c_float * 3), ("inv_std", C.c_float * 3),
        ("out_channels", C.c_int32), ("num_frames", C.c_int32),
    ]


_lib: Optional[C.CDLL] = None
_inited_device: Optional[int] = None

_VP, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float

_SIGNATURES = {
    "vcb_init": ([_I32], _I32),
    "vcb_last_error_string": ([], C.c_char_p),
    "vcb_last_fault": ([C.POINTER(_I32)], _I32),
    "vcb_version": ([], _I32),
    "vcb_set_option": ([C.c_char_p, _I32], _I32),
    "vcb_get_option": ([C.c_char_p], _I32),
    "vcb_read_prof": ([C.POINTER(C.c_uint64)], _I32),
    "vcb_h2d_frames_inplace": ([_VP, _VP, _I32, _I64, _VP], _I32),
    "vcb_conv_packed_sizes": ([C.POINTER(ConvDesc), C.POINTER(_I64), C.POINTER(_I64)], _I32),
    "vcb_conv_pack_weights": ([C.POINTER(ConvDesc), _VP, _VP, _VP, _VP, _VP], _I32),
    "vcb_conv2d_fwd": ([C.POINTER(ConvDesc), _VP, _VP, _VP, _VP, _VP, _VP], _I32),
    "vcb_conv2d_fwd_stats": ([C.POINTER(ConvDesc), _VP, _VP, _VP, _VP, _VP, _VP, _VP], _I32),
    "vcb_conv_out_hw": ([C.POINTER(ConvDesc), C.POINTER(_I32), C.POINTER(_I32)], _I32),
    "vcb_frames_to_f16c4": ([_VP, _VP, _I32, _I32, _I32, _VP], _I32),
    "vcb_frames_to_f16_s2d": ([_VP, _VP, _I32, _I32, _I32, _VP], _I32),
    "vcb_frames_to_f16_s2d_wpad": ([_VP, _VP, _I32, _I32, _I32, _VP], _I32),
    "vcb_letterbox_half_u8": ([_VP, _I32, _I32, _I32, _VP, _I32, _I32, _I32, _I32, _I32, _VP], _I32),
    "vcb_letterbox_bilinear_u8": ([_VP, _I32, _I32, _I32, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _VP, _VP, _I32, _VP], _I32),
    "vcb_upsample2x": ([_VP, _I32, _VP, _I32, _I32, _I32, _I32, _I32, _VP], _I32),
    "vcb_sppf_pool": ([_VP, _I32, _I32, _I32, _I32, _I32, _VP], _I32),
    "vcb_maxpool": ([_VP, _I32, _VP, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _VP], _I32),
    "vcb_detect_decode": ([C.POINTER(DetectDesc), _VP, _VP, _VP, _VP, _VP, _VP], _I32),
    "vcb_nms_workspace_bytes": ([_I32, _I32], _I64),
    "vcb_nms": ([C.POINTER(NmsDesc), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP], _I32),
    "vcb_roi_resize_norm": ([C.POINTER(RoiDesc), _VP, _I32, _I32, _VP, _VP, _VP], _I32),
    "vcb_boxes_to_rois": ([_VP, _VP, _I32, _I32, _I32, _VP, _VP], _I32),
    "vcb_roi_stem_patches": ([C.POINTER(RoiDesc), _VP, _I32, _I32, _VP, _VP, _VP], _I32),
    "vcb_reid_stem_pool": ([_VP, _VP, _VP, _VP, _I32, _VP], _I32),
    "vcb_reid_stem_direct": ([C.POINTER(RoiDesc), _VP, _I32, _I32, _VP, _VP, _VP, _VP], _I32),
    "vcb_reid_stem_direct_stats": ([C.POINTER(RoiDesc), _VP, _I32, _I32, _VP, _VP, _VP, _VP, _VP], _I32),
    "vcb_reid_stem_direct_bn": ([C.POINTER(RoiDesc), _VP, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _VP], _I32),
    "vcb_avgpool_l2norm": ([_VP, _I32, _I32, _I32, _I32, _VP, _VP], _I32),
    "vcb_bn_train_stats": ([_VP, _I32, _VP, _I32, _VP, _VP, _F, _VP, _VP, _VP], _I32),
    "vcb_bn_apply": ([_VP, _I32, _I32, _VP, _VP, _VP, _VP, _I32, _I32, _VP, _I32, _VP], _I32),
    "vcb_reid_stem_stats": ([_VP, _VP, _VP, _I32, _VP, _VP, _VP], _I32),
    "vcb_bn_seg_finalize": ([_VP, _VP, _I32, _I32, _I32, _VP, _VP, _VP, _F, _VP, _VP], _I32),
    "vcb_reid_stem_pool_bn": ([_VP, _VP, _VP, _VP, _VP, _I32, _VP], _I32),
    "vcb_bn_seg_stats_f16": ([_VP, _I32, _I32, _I32, _VP, _VP, _VP], _I32),
    "vcb_bn_seg_apply_fused_f16": ([_VP, _I32, _I32, _I32, _VP, _VP, _VP, _VP, _VP, _F, _VP, _I32, _VP, _VP, _VP, _I32, _VP, _I32, _VP], _I32),
    "vcb_bn_seg_apply_f16": ([_VP, _I32, _I32, _I32, _I32, _VP, _VP, _VP, _I32, _I32, _I32, _VP, _I32, _VP], _I32),
    "vcb_graph_begin": ([_VP], _I32),
    "vcb_graph_end": ([_VP, C.POINTER(_VP)], _I32),
    "vcb_graph_launch": ([_VP, _VP], _I32),
    "vcb_graph_num_kernels": ([_VP], _I32),
    "vcb_graph_destroy": ([_VP], _I32),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES.keys())


def load() -> C.CDLL:
    """dlopen libvcb200.so and declare every prototype (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise VcbError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C vehicle_counting_b200/csrc`). There is no CPU/PyTorch fallback for the hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def last_error() -> str:
    return load().vcb_last_error_string().decode("utf-8", "replace")


def last_fault():
    out = (_I32 * 4)()
    load().vcb_last_fault(out)
    return tuple(out)


def check(rc: int, what: str = "") -> None:
    if rc != VCB_OK:
        raise VcbError(f"{what or 'libvcb200'} failed (rc={rc}): {last_error()}")


def init(device: int = 0) -> C.CDLL:
    """Load the library and bind it to a CUDA device (sm_100 required)."""
    global _inited_device
    lib = load()
    if _inited_device != device:
        check(lib.vcb_init(device), "vcb_init")
        _inited_device = device
    return lib


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL; ints pass through as raw addresses)."""
    if t is None:
        return 0
    if isinstance(t, int):
        return t
    return t.data_ptr()


def stream_handle(stream=None) -> int:
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream
