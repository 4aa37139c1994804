"""Crema-D QMF without the joint CE term (cremad/joint_model_qmf_ablate_Ljoint.py of the reference):
loss = sum_m CE(z_m) + L_reg."""
import torch.nn as nn

from ..utils.BaseModel import QMFBaseModel
from ._qmf_variants import LF_LOSS_NO_JOINT, QmfFusionNet


class FusionNet(QmfFusionNet):
    def __init__(self, args, loss_fn):
        super().__init__(args, loss_fn, loss_terms=LF_LOSS_NO_JOINT)


class MultimodalCremadModel(QMFBaseModel):
    def _build_model(self):
        return FusionNet(args=self.args, loss_fn=nn.CrossEntropyLoss())
