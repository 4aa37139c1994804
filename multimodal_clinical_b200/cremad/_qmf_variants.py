"""Shared FusionNet for the QMF loss variants of Crema-D: the reference's ablation files differ from
cremad/joint_model_qmf.py in ONE line each (which terms enter ``loss``), so they share the head here and
pass the dropped terms to the fused step as LF_LOSS_* bits (include/lf_fusion.h)."""
import torch.nn as nn

from .._lib import LF_LOSS_NO_JOINT, LF_LOSS_NO_UNI  # noqa: F401  (re-exported to the ablation modules)
from ..existing_algos.QMF import QMF
from ..heads import FusedLateFusionHead
from ._pool import pool_features
from .backbone import resnet18


class QmfFusionNet(nn.Module):
    """cremad/joint_model_qmf.py:13-75 with ``loss_terms`` selecting the ablation:
    0 full loss; LF_LOSS_NO_JOINT: ``loss_joint = 0`` (joint_model_qmf_ablate_Ljoint.py:68);
    LF_LOSS_NO_UNI: the unimodal CE sum is dropped (joint_model_qmf_ablate_Lunimodal.py:70).
    The History still receives the unimodal batch-mean losses in every variant (:63-65)."""

    def __init__(self, args, loss_fn, loss_terms=0):
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
        self.fused = FusedLateFusionHead(self.num_classes, mode="qmf", n_data=self.args.num_samples, loss_terms=loss_terms)
        self.fused.bind_qmf(self.qmf)

    def forward(self, x1_data, x2_data, label, idx):
        a, v = pool_features(self.x1_model(x1_data), self.x2_model(x2_data))
        return self.fused(a, v, self.x1_classifier, self.x2_classifier, label, idx)
