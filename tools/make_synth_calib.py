"""Measure, with the CPU oracle, the per-layer conv-output variance that
vehicle_counting_b200.weights.synth_yolov5_state_dict uses to keep synthetic activations O(1):
writes vehicle_counting_b200/data/synth_calib.json.  Run once in the build container:
    python tools/make_synth_calib.py
"""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import yolov5 as Y                                            # noqa: E402
from vehicle_counting_b200.weights import synth_yolov5_state_dict        # noqa: E402


def calibrate(name: str, seed: int = 0):
    ones = {}
    model = Y.DetectionModel(name)
    sd = synth_yolov5_state_dict(name, seed, calib=_Ones())
    model.load_state_dict(sd, strict=False)
    model.eval()
    table = {}
    names = {m: n for n, m in model.named_modules()}

    def pre_hook(bn, inp):
        v = inp[0].var().item()
        prefix = names[bn][:-len(".bn")]
        table[prefix] = v
        bn.running_mean.data *= math.sqrt(v)
        bn.running_var.data *= v

    hs = [m.bn.register_forward_pre_hook(pre_hook) for m in model.modules() if isinstance(m, Y.Conv)]
    g = torch.Generator().manual_seed(12345)
    with torch.no_grad():
        model(torch.rand(1, 3, 320, 320, generator=g))
    for h in hs:
        h.remove()
    return table


class _Ones(dict):
    def get(self, k, d=None):
        return 1.0


def main():
    out = {}
    for name in ("yolov5n", "yolov5s", "yolov5m", "yolov5l", "yolov5x"):
        out[name] = calibrate(name)
        print(name, len(out[name]), min(out[name].values()), max(out[name].values()))
    path = os.path.join(ROOT, "vehicle_counting_b200", "data", "synth_calib.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", path)


if __name__ == "__main__":
    main()
