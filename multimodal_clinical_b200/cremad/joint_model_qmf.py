"""Crema-D QMF late fusion (cremad/joint_model_qmf.py of the reference) on the fused step."""
import torch.nn as nn

from ..existing_algos.QMF import QMF
from ..heads import FusedLateFusionHead
from ..utils.BaseModel import QMFBaseModel
from ._pool import pool_features
from .backbone import resnet18


class FusionNet(nn.Module):
    def __init__(self, args, loss_fn):
        super().__init__()
        self.args = args
        self.num_modality = 2
        self.qmf = QMF(self.num_modality, self.args.num_samples)
        self.x1_model = resnet18(modality='audio')
        self.x1_classifier = nn.Linear(512, self.args.num_classes)
        self.x2_model = resnet18(modality='visual')
        self.x2_classifier = nn.Linear(512, self.args.num_classes)
        self.num_classes = self.args.num_classes
        self.loss_fn = loss_fn
        self.fused = FusedLateFusionHead(self.num_classes, mode="qmf", n_data=self.args.num_samples)
        self.fused.bind_qmf(self.qmf)

    def forward(self, x1_data, x2_data, label, idx):
        """-> (x1_logits, x2_logits, avg_logits, loss, logits_df);
        loss = CE(logits_df) + CE(x1) + CE(x2) + ranking regulariser, History updated with idx."""
        a, v = pool_features(self.x1_model(x1_data), self.x2_model(x2_data))
        return self.fused(a, v, self.x1_classifier, self.x2_classifier, label, idx)


class MultimodalCremadModel(QMFBaseModel):
    def __init__(self, args):
        super().__init__(args)

    def _build_model(self):
        return FusionNet(args=self.args, loss_fn=nn.CrossEntropyLoss())
