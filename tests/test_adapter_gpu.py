"""The reference's stage-wrapper chain on the GPU path: ImageDetect(args, config).run(batch) -> Detector.inference_step ->
YoloBackbone.detect (/root/reference/modules/detect.py:30-60, networks/detector.py:36-38, networks/yolo.py:68-99): output
contract (types, shapes, order, empty images, class remapping) and agreement with the oracle's restatement of the adapter."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cfg(name="yolov5n"):
    return types.SimpleNamespace(model_name=name, min_iou=0.45, min_conf=0.25, max_det=300)


def test_image_detect_chain_contract_and_oracle_agreement(lib, monkeypatch):
    from oracle import yolov5 as Y
    from vehicle_counting_b200.modules import ImageDetect
    from vehicle_counting_b200.networks import yolo as NY
    z = np.load(os.path.join(GOLD, "yolo_golden.npz"))
    imgs = list(z["imgs"])                                           # 2 RGB uint8 frames of the CPU golden set
    model = Y.build("yolov5n", seed=0, obj_bias=-1.0)
    # the wrapper builds its network through get_model(args, config): hand it the oracle's seeded weights
    monkeypatch.setattr(NY, "load_yolov5_checkpoint", lambda path: (model.state_dict(), None))
    args = types.SimpleNamespace(weight="seeded-yolov5n.pt", mapping=None, mapping_dict=None)
    det = ImageDetect(args, _cfg())
    det.model.model.size = max(imgs[0].shape[:2])                    # AutoShape size= (the goldens were made at the frame size)
    out = det.run({"imgs": imgs, "frames": [1, 2], "ori_imgs": imgs})
    assert set(out) == {"boxes", "labels", "scores"} and all(len(v) == len(imgs) for v in out.values())
    want = Y.yolo_backbone_detect(Y.fp16_storage_twin(model), {"imgs": imgs}, size=max(imgs[0].shape[:2]))
    n_firm = n_hit = 0
    for b in range(len(imgs)):
        boxes, labels, scores = out["boxes"][b], out["labels"][b], out["scores"][b]
        assert boxes.dtype == np.float64 and boxes.ndim == 2 and boxes.shape[1] == 4
        assert labels.dtype == np.int64 and labels.shape == (boxes.shape[0],) and scores.shape == (boxes.shape[0],)
        assert (np.diff(scores) <= 0).all() and boxes.shape[0] <= 300
        assert (boxes[:, 2:] >= 0).all()                             # x, y, w, h with top-left origin
        wb, wl, ws = want[b]["bboxes"], want[b]["classes"], want[b]["scores"]
        for i in range(len(ws)):                                     # every firm oracle row has a twin (same class, IoU > 0.9)
            if abs(ws[i] - 0.25) < 5e-3:
                continue
            n_firm += 1
            same = np.nonzero(labels == wl[i])[0]
            if len(same) == 0:
                continue
            x1 = np.maximum(wb[i, 0], boxes[same, 0]); y1 = np.maximum(wb[i, 1], boxes[same, 1])
            x2 = np.minimum(wb[i, 0] + wb[i, 2], boxes[same, 0] + boxes[same, 2]); y2 = np.minimum(wb[i, 1] + wb[i, 3], boxes[same, 1] + boxes[same, 3])
            inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
            iou = inter / (wb[i, 2] * wb[i, 3] + boxes[same, 2] * boxes[same, 3] - inter)
            n_hit += bool((iou > 0.9).any())
    assert n_firm > 0 and n_hit >= 0.9 * n_firm, (n_hit, n_firm)


def test_image_detect_empty_frames_and_class_mapping(lib, monkeypatch):
    from oracle import yolov5 as Y
    from vehicle_counting_b200.modules import ImageDetect
    from vehicle_counting_b200.networks import yolo as NY
    z = np.load(os.path.join(GOLD, "yolo_golden.npz"))
    imgs = list(z["imgs"])
    quiet = Y.build("yolov5n", seed=0, obj_bias=-30.0)               # nothing passes the confidence threshold
    monkeypatch.setattr(NY, "load_yolov5_checkpoint", lambda path: (quiet.state_dict(), None))
    det = ImageDetect(types.SimpleNamespace(weight="w.pt", mapping=None, mapping_dict=None), _cfg())
    out = det.run({"imgs": imgs})
    for b in range(len(imgs)):                                       # networks/yolo.py:92-97: three empty arrays
        assert out["boxes"][b].shape == (0,) and out["labels"][b].shape == (0,) and out["scores"][b].shape == (0,)
    loud = Y.build("yolov5n", seed=0, obj_bias=-1.0)
    monkeypatch.setattr(NY, "load_yolov5_checkpoint", lambda path: (loud.state_dict(), None))
    plain = ImageDetect(types.SimpleNamespace(weight="w2.pt", mapping=None, mapping_dict=None), _cfg()).run({"imgs": imgs})
    seen = sorted({int(c) for l in plain["labels"] for c in l})
    assert seen, "the seeded network must detect something on the golden frames"
    keep = seen[: max(1, len(seen) // 2)]
    mapping = {c: j for j, c in enumerate(keep)}                      # detect.py:41-46: included_classes = keys, label = mapping[class - 1]
    mapped = ImageDetect(types.SimpleNamespace(weight="w3.pt", mapping=mapping, mapping_dict=None), _cfg())
    assert mapped.class_names == [mapped.model.model.class_names[i] for i in sorted(set(mapping.values()))]
