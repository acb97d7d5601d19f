"""ORACLE helper (test infrastructure): import the UNMODIFIED reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference): used by
oracle/make_goldens.py to generate tests/golden/*.npz and by `-m "not gpu"` tests that skip
when the reference is absent.  Recipe from SURVEY §8(c): stub the four import-time
dependencies the container lacks and restore the NumPy aliases removed in 1.24
(/root/reference/networks/deepsort/sort/detection.py:30, sort/preprocessing.py:40,
deep_sort.py:56 use np.float / np.int).
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("VCB_REFERENCE_ROOT", "/root/reference")
REID_CKPT = os.path.join(REF_ROOT, "networks/deepsort/deep/checkpoint/ckpt.t7")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "networks/deepsort/deep_sort.py"))


def install() -> None:
    """Idempotent: after this, `import networks`, `import modules` resolve to the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    import numpy as np
    if not hasattr(np, "float"):
        np.float = float     # type: ignore[attr-defined]
    if not hasattr(np, "int"):
        np.int = int         # type: ignore[attr-defined]

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Any()

        def __getattr__(self, k):
            return _Any()

    A = stub("albumentations", Compose=_Any, Resize=_Any, LongestMaxSize=_Any, PadIfNeeded=_Any,
             Normalize=_Any, BboxParams=_Any)
    Ap = stub("albumentations.pytorch")
    Apt = stub("albumentations.pytorch.transforms", ToTensorV2=_Any, ToTensor=_Any)
    A.pytorch = Ap
    Ap.transforms = Apt
    import collections
    _RGB = collections.namedtuple("IntegerRGB", "red green blue")
    stub("webcolors", name_to_rgb=lambda n: _RGB(0, 0, 0), hex_to_rgb=lambda h: _RGB(0, 0, 0))
    stub("gdown", download=lambda *a, **k: None)
    mpl = stub("matplotlib", use=lambda *a, **k: None)
    mpl.pyplot = stub("matplotlib.pyplot")
    mpl.patches = stub("matplotlib.patches")
    mpl.colors = stub("matplotlib.colors")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def reference_extractor(use_cuda: bool = False):
    install()
    from networks.deepsort.deep.feature_extractor import Extractor  # type: ignore
    return Extractor(REID_CKPT, use_cuda=use_cuda)


def reference_deepsort(**kw):
    install()
    from networks.deepsort import DeepSort  # type: ignore
    return DeepSort(REID_CKPT, use_cuda=False, **kw)
