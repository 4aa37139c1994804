"""Crema-D QMF loss + OGM-GE modulation of the encoders' conv gradients
(cremad/joint_model_ogm_ge_lreg.py of the reference: FusionNet :13-75 is the QMF one, the training step
:101-161 is manual optimisation with ``ogm_ge`` between backward and the optimiser step)."""
import torch.nn as nn

from ..existing_algos.OGM_GE import ogm_ge
from ..utils.BaseModel import QMFBaseModel
from ._qmf_variants import QmfFusionNet


class FusionNet(QmfFusionNet):
    def __init__(self, args, loss_fn):
        super().__init__(args, loss_fn, loss_terms=0)


class MultimodalCremadModel(QMFBaseModel):
    def __init__(self, args):
        super().__init__(args)
        self.automatic_optimization = False
        self.ogm_modulation = self.args.grad_mod_type
        self.ogm_alpha = self.args.alpha

    def training_step(self, batch, batch_idx):
        x1, x2, label, idx = batch
        x1_logits, x2_logits, avg_logits, loss, logits_df = self.model(x1, x2, label, idx)
        self._log_train_step(loss, self._step_accuracies(), with_df=True)
        opt = self.optimizers()
        opt.zero_grad()
        self.manual_backward(loss)
        if self.ogm_modulation:      # the score sums of these logits were reduced by the fused step: no second pass
            ogm_ge(self.model, x1_logits, x2_logits, label, modulation=self.ogm_modulation, alpha=self.ogm_alpha)
        opt.step()
        return loss

    def _build_model(self):
        return FusionNet(args=self.args, loss_fn=nn.CrossEntropyLoss())
