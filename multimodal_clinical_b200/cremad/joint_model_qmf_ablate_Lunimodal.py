"""Crema-D QMF without the unimodal CE terms (cremad/joint_model_qmf_ablate_Lunimodal.py of the reference):
loss = CE(z_df) + L_reg; the History is still fed the unimodal batch-mean losses."""
import torch.nn as nn

from ..utils.BaseModel import QMFBaseModel
from ._qmf_variants import LF_LOSS_NO_UNI, QmfFusionNet


class FusionNet(QmfFusionNet):
    def __init__(self, args, loss_fn):
        super().__init__(args, loss_fn, loss_terms=LF_LOSS_NO_UNI)


class MultimodalCremadModel(QMFBaseModel):
    def _build_model(self):
        return FusionNet(args=self.args, loss_fn=nn.CrossEntropyLoss())
