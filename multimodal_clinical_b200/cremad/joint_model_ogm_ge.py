"""Crema-D late fusion with OGM-GE (cremad/joint_model_ogm_ge.py of the reference) on the fused step."""
import torch.nn as nn

from ..heads import FusedLateFusionHead
from ..utils.BaseModel import OGMGEBaseModel
from ._pool import pool_features
from .backbone import resnet18


class FusionNet(nn.Module):
    def __init__(self, num_classes, loss_fn):
        super().__init__()
        self.x1_model = resnet18(modality='audio')
        self.x1_classifier = nn.Linear(512, num_classes)
        self.x2_model = resnet18(modality='visual')
        self.x2_classifier = nn.Linear(512, num_classes)
        self.num_classes = num_classes
        self.loss_fn = loss_fn          # nn.CrossEntropyLoss() (mean): what the fused step implements
        self.fused = FusedLateFusionHead(num_classes, mode="jlogits")

    def forward(self, x1_data, x2_data, label):
        """-> (x1_logits, x2_logits, avg_logits, loss); loss = CE((x1+x2)/2, label)."""
        a, v = pool_features(self.x1_model(x1_data), self.x2_model(x2_data))
        return self.fused(a, v, self.x1_classifier, self.x2_classifier, label)


class MultimodalCremadModel(OGMGEBaseModel):
    def __init__(self, args):
        super().__init__(args)
        self.model.fused.ogm_alpha = self.ogm_alpha     # coefficients come out of the same pass as the loss

    def _build_model(self):
        return FusionNet(num_classes=self.args.num_classes, loss_fn=nn.CrossEntropyLoss())
