"""`ImageDetect` with the reference's surface (/root/reference/modules/detect.py:8-60)."""
from __future__ import annotations

import numpy as np
import torch

from ..networks import Detector, get_model


class ImageDetect:
    def __init__(self, args, config):
        if not torch.cuda.is_available():
            raise RuntimeError("vehicle_counting_b200 needs a CUDA device (sm_100); there is no CPU path")
        self.device = torch.device("cuda")
        self.mapping_dict = getattr(args, "mapping", None)
        net = get_model(args, config)
        self.class_names = net.class_names
        if self.mapping_dict is not None:
            self.included_classes = list(self.mapping_dict.keys())
            class_ids = list(self.mapping_dict.values())
            self.class_names = [self.class_names[i] for i in sorted(np.unique(class_ids))]
        self.model = Detector(model=net, device=self.device)
        self.model.eval()
        for p in self.model.parameters():
            p.requires_grad = False

    def run(self, batch):
        boxes_result, labels_result, scores_result = [], [], []
        preds = self.model.inference_step(batch)
        for outputs in preds:
            if self.mapping_dict is not None:      # detect.py:41-46 (off unless --mapping is given)
                keep_idx = [e for e, i in enumerate(outputs["classes"]) if i in self.included_classes]
                labels = [self.mapping_dict[int(i) - 1] for i in outputs["classes"][keep_idx]]
                outputs["classes"] = np.array(labels)
                outputs["scores"] = outputs["scores"][keep_idx]
                outputs["bboxes"] = outputs["bboxes"][keep_idx]
            boxes_result.append(outputs["bboxes"])
            labels_result.append(outputs["classes"])
            scores_result.append(outputs["scores"])
        return {"boxes": boxes_result, "labels": labels_result, "scores": scores_result}
