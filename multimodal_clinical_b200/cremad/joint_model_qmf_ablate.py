"""Crema-D ``qmf_ablate`` (cremad/joint_model_qmf_ablate.py of the reference): trains with the plain
late-fusion loss CE((z1+z2)/2) (``istrain=True`` branch, :60-66) and evaluates with the full QMF block
(:68-84), History update included.  Two fused heads share the classifiers and the EMA state."""
import torch
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
        self.fused_train = FusedLateFusionHead(self.num_classes, mode="jlogits")
        self.fused_eval = FusedLateFusionHead(self.num_classes, mode="qmf", n_data=self.args.num_samples)
        self.fused_eval.bind_qmf(self.qmf)
        self.fused = self.fused_train             # the head whose statistics the LightningModule reads

    def forward(self, x1_data, x2_data, label, idx, istrain=None):
        if istrain is None:                       # the reference passes istrain=False from its validation / test steps
            istrain = self.training
        a, v = pool_features(self.x1_model(x1_data), self.x2_model(x2_data))
        if istrain:
            self.fused = self.fused_train
            return self.fused_train(a, v, self.x1_classifier, self.x2_classifier, label)
        self.fused = self.fused_eval
        return self.fused_eval(a, v, self.x1_classifier, self.x2_classifier, label, idx)


class MultimodalCremadModel(QMFBaseModel):
    def __init__(self, args):
        super().__init__(args)
        self.model.fused_eval.bind_ema(self.ema_offset)

    def forward(self, x1, x2, label, idx, istrain=None):
        return self.model(x1, x2, label, idx, istrain)

    def training_step(self, batch, batch_idx):
        x1, x2, label, idx = batch
        x1_logits, x2_logits, avg_logits, loss = self.model(x1, x2, label, idx, True)
        acc = self._step_accuracies()
        acc["df"] = torch.zeros_like(acc["joint"])            # no fused logits on the training branch
        self._log_train_step(loss, acc, with_df=True)
        return loss

    def _build_model(self):
        return FusionNet(args=self.args, loss_fn=nn.CrossEntropyLoss())
