"""CPU tests of the C-ABI boundary: the shared library loads, exports every symbol include/vcb200.h declares,
host-only entry points work without a GPU, and compute entry points fail loudly (no fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "vcb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vcb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vehicle_counting_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"libvcb200.so does not export {name}"
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS), set(declared) ^ set(_lib.EXPORTED_SYMBOLS)


def test_struct_layouts_match_header_sizes():
    from vehicle_counting_b200 import _lib as L
    assert C.sizeof(L.ConvDesc) == 22 * 4
    assert C.sizeof(L.DetectLevel) == 48
    assert C.sizeof(L.DetectDesc) == 24 + 4 * 48 + 4 + 32 + 4      # + use_class_mask, class_mask[8], tail padding to 8
    assert C.sizeof(L.NmsDesc) == 24 + 5 * 8
    assert C.sizeof(L.RoiDesc) == 8 + 24 + 4 + 4


def test_conv_geometry_host_functions():
    from vehicle_counting_b200 import _lib as L, ops
    d = ops.make_conv_desc(4, 640, 640, 3, 48, 6, 2, 2, cin_pitch=4)           # YOLOv5m stem
    assert ops.conv_out_hw(d) == (320, 320)
    assert ops.conv_packed_sizes(d) == (48 * 192, 48)                           # 36 taps x 4 ch -> 3 K-steps of 64
    d = ops.make_conv_desc(2, 25, 25, 64, 128, 3, 2, 1)                         # ReID layer2.0 conv1, odd size
    assert ops.conv_out_hw(d) == (13, 13)
    d = ops.make_conv_desc(1, 20, 20, 768, 768, 3, 1, 1)
    w, b = ops.conv_packed_sizes(d)
    assert b == 768 and w == 768 * 9 * 768
    d = ops.make_conv_desc(1, 20, 20, 512, 255, 1, 1, 0, cout_pitch=256, out_dtype=L.F32, act=L.ACT_NONE)
    assert ops.conv_packed_sizes(d) == (256 * 512, 256)


def test_invalid_descriptors_are_rejected_with_a_message():
    from vehicle_counting_b200 import _lib as L, ops
    with pytest.raises(L.VcbError, match="multiples of 8"):
        ops.conv_out_hw(ops.make_conv_desc(1, 8, 8, 12, 16, 3, 1, 1))
    with pytest.raises(L.VcbError, match="empty output"):
        ops.conv_out_hw(ops.make_conv_desc(1, 2, 2, 8, 16, 5, 1, 0))
    with pytest.raises(L.VcbError, match="cout_pitch"):
        ops.conv_out_hw(ops.make_conv_desc(1, 8, 8, 8, 16, 3, 1, 1, cout_pitch=8))


def test_compute_entry_points_fail_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from vehicle_counting_b200 import _lib as L
    lib = L.load()
    assert lib.vcb_init(0) != 0                       # no device: error code + message, never a silent fallback
    assert len(L.last_error()) > 0
    d = L.ConvDesc()
    assert lib.vcb_conv2d_fwd(C.byref(d), None, None, None, None, None, None) != 0
    assert "vcb_init" in L.last_error()


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "vehicle_counting_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_stage_wrapper_mirrors_keep_the_reference_signatures():
    """modules/detect.py:8-60 and modules/track.py:8-70: same constructor / run parameters as the reference classes (compared with the
    reference sources when they are mounted, with the recorded parameter lists otherwise)."""
    import inspect
    from vehicle_counting_b200.modules import ImageDetect, VideoTracker
    want = {"ImageDetect.__init__": ["self", "args", "config"], "ImageDetect.run": ["self", "batch"],
            "VideoTracker.__init__": ["self", "num_classes", "cam_config", "video_info", "deepsort_chepoint"],
            "VideoTracker.build_tracker": ["self", "checkpoint", "cam_cfg"],
            "VideoTracker.run": ["self", "image", "boxes", "labels", "scores"]}
    ref_root = "/root/reference/modules"
    if os.path.isdir(ref_root):           # build container: read the parameter lists from the reference sources themselves
        import ast
        for fn, cls in (("detect.py", "ImageDetect"), ("track.py", "VideoTracker")):
            tree = ast.parse(open(os.path.join(ref_root, fn)).read())
            node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls)
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f"{cls}.{f.name}" in want:
                    assert [a.arg for a in f.args.args] == want[f"{cls}.{f.name}"], (cls, f.name)
    for key, params in want.items():
        cls, name = key.split(".")
        got = list(inspect.signature(getattr({"ImageDetect": ImageDetect, "VideoTracker": VideoTracker}[cls], name)).parameters)
        assert got[:len(params)] == params, (key, got)          # extra keyword-only knobs (bn_mode) may follow
