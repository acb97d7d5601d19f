"""`Detector` shell with the reference's surface (/root/reference/networks/detector.py:11-38,
base_model.py:6-57): an nn.Module that owns the backbone, `.eval()`, `.parameters()`,
`.inference_step(batch)`.  The training-side members of the reference's BaseModel (optimizer, criterion,
scaler, metrics) are not on the inference path and are accepted but unused."""
from __future__ import annotations

from torch import nn


class Detector(nn.Module):
    def __init__(self, model, device=None, optimizer=None, criterion=None, metrics=None, scaler=None, freeze=False, **kwargs):
        super().__init__()
        self.model = model
        self.model_name = "yolov5"
        self.device = device
        self.freeze = freeze
        if self.freeze:
            for p in self.model.parameters():
                p.requires_grad = False

    def forward(self, x):
        return self.model(x)

    def inference_step(self, batch):
        return self.model.detect(batch, self.device)
