"""Crema-D plain late fusion (cremad/joint_model.py of the reference) on the fused step."""
import torch.nn as nn

from ..utils.BaseModel import JointLogitsBaseModel
from .joint_model_ogm_ge import FusionNet


class MultimodalCremadModel(JointLogitsBaseModel):
    def __init__(self, args):
        super().__init__(args)

    def _build_model(self):
        return FusionNet(num_classes=self.args.num_classes, loss_fn=nn.CrossEntropyLoss())
