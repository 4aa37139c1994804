"""Crema-D per-modality ensemble with OGM-GE modulation (cremad/ensemble_model_noised.py of the reference,
``model_type: ensemble_ogm_ge``) on the fused step: one cross-entropy per modality, averaged for the backward pass,
encoder gradients modulated by the OGM-GE coefficients that come out of the same pass."""
import torch.nn as nn

from ..existing_algos.OGM_GE import ogm_ge
from ..heads import FusedLateFusionHead
from ..utils.BaseModel import EnsembleBaseModel
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
        self.fused = FusedLateFusionHead(num_classes, mode="ensemble")

    def forward(self, x1_data, x2_data, label):
        """-> (x1_logits, x2_logits, x1_loss, x2_loss) with x_m_loss = CE(x_m_logits, label) (:49-55)."""
        a, v = pool_features(self.x1_model(x1_data), self.x2_model(x2_data))
        return self.fused(a, v, self.x1_classifier, self.x2_classifier, label)


class MultimodalCremadModel(EnsembleBaseModel):
    def __init__(self, args):
        super().__init__(args)
        self.automatic_optimization = False
        self.ogm_modulation = self.args.grad_mod_type
        self.ogm_alpha = self.args.alpha
        self.model.fused.ogm_alpha = self.ogm_alpha     # coefficients come out of the same pass as the losses

    def training_step(self, batch, batch_idx):
        """zero_grad -> backward of (x1_loss + x2_loss) / 2 -> ogm_ge -> step (:93-123)."""
        x1, x2, label = batch
        x1_logits, x2_logits, x1_loss, x2_loss = self.model(x1, x2, label)
        avg_loss = (x1_loss + x2_loss) / 2
        self._record("train", avg_loss, self._accs(), self.train_metrics)
        opt = self.optimizers()
        opt.zero_grad()
        self.manual_backward(avg_loss)
        if self.ogm_modulation:
            ogm_ge(self.model, x1_logits, x2_logits, label, modulation=self.ogm_modulation, alpha=self.ogm_alpha)
        opt.step()
        return avg_loss

    def _build_model(self):
        return FusionNet(num_classes=self.args.num_classes, loss_fn=nn.CrossEntropyLoss())
