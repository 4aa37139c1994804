"""LightningModule shared by the reference's stand-alone mean-fusion modules with M modalities or unequal feature widths
(mustard/joint_model.py:86-300 ``MultimodalMustardModel``, avmnist/joint_model.py:139-330 ``MultimodalAVMnistModel``):
same hooks, same logged keys (``train_loss`` / ``train_acc``, ``val_loss`` / ``val_acc`` / ``x{m}_val_acc``,
``avg_test_loss`` / ``avg_test_acc`` / ``x{m}_test_acc``), Adam(lr) as optimizer.  ``self.model`` is a FusionNet whose
heads run on the fused kernel (multi.FusedMeanFusionHeads) and returns ``(z_1, ..., z_M, avg_logits, loss)``."""
from abc import ABC, abstractmethod

import torch

from .lightning_compat import pl


def offset_corrected_accuracies(logits: torch.Tensor, labels: torch.Tensor):
    """Epoch-end unimodal offset correction for M modalities (mustard/joint_model.py:181-195): logits (N, M, C) ->
    per-modality accuracies of ``logits + (mean_m(mean_n logits) - mean_n logits)``.  Runs once per epoch on the collected
    logits, like the reference (the two-modality (N, 2, C) case of the main path has its own device kernels, lf_epoch.cu)."""
    m_out = torch.mean(logits, dim=0)
    offset = torch.mean(m_out, dim=0, keepdim=True) - m_out
    corrected = logits + offset
    return [torch.mean((torch.argmax(corrected[:, m, :], dim=1) == labels).float()) for m in range(logits.shape[1])]


class MeanFusionMultiBaseModel(pl.LightningModule, ABC):
    def __init__(self, args):
        super().__init__()
        self.args = args
        self.model = self._build_model()
        self.val_metrics = {"val_loss": [], "val_acc": [], "val_logits": [], "val_labels": []}
        self.test_metrics = {"test_loss": [], "test_acc": [], "test_logits": [], "test_labels": []}

    def forward(self, *batch):
        return self.model(*batch)

    def _convert_type(self, batch):
        *xs, label = batch
        return tuple(x.to(torch.float32) for x in xs) + (label,)

    def _run(self, batch):
        out = self.model(*self._convert_type(batch))
        *zs, avg_logits, loss = out
        # the joint hit count comes out of the fused step (stats[1]); a 0-d device tensor, no host sync
        step = self.model.fused.last_step
        joint_acc = (step.stats[1] / step.batch).float()
        return zs, avg_logits, loss, joint_acc

    def training_step(self, batch, batch_idx):
        _, _, loss, joint_acc = self._run(batch)
        self.log("train_loss", loss, on_step=True, on_epoch=True, prog_bar=False, logger=True)
        self.log("train_acc", joint_acc, on_step=True, on_epoch=True, prog_bar=False, logger=True)
        return loss

    def _eval_step(self, kind, metrics, batch):
        zs, _, loss, joint_acc = self._run(batch)
        self.log(f"{kind}_loss", loss, on_step=True, on_epoch=True, prog_bar=False, logger=True)
        self.log(f"{kind}_acc", joint_acc, on_step=True, on_epoch=True, prog_bar=False, logger=True)
        metrics[f"{kind}_logits"].append(torch.stack(tuple(zs), dim=1))
        metrics[f"{kind}_labels"].append(batch[-1])
        metrics[f"{kind}_loss"].append(loss.detach())
        metrics[f"{kind}_acc"].append(joint_acc)
        return loss

    def _eval_epoch_end(self, kind, metrics, loss_key, acc_key):
        labels = torch.cat(metrics[f"{kind}_labels"], dim=0).flatten()
        logits = torch.cat(metrics[f"{kind}_logits"], dim=0)
        accs = offset_corrected_accuracies(logits, labels)
        self.log(loss_key, torch.stack(metrics[f"{kind}_loss"]).mean(), on_step=False, on_epoch=True, prog_bar=False, logger=True)
        self.log(acc_key, torch.stack(metrics[f"{kind}_acc"]).mean(), on_step=False, on_epoch=True, prog_bar=False, logger=True)
        for m, a in enumerate(accs):
            self.log(f"x{m + 1}_{kind}_acc", a, on_step=False, on_epoch=True, prog_bar=False, logger=True)
        for v in metrics.values():
            v.clear()

    def validation_step(self, batch, batch_idx):
        return self._eval_step("val", self.val_metrics, batch)

    def on_validation_epoch_end(self) -> None:
        self._eval_epoch_end("val", self.val_metrics, "val_loss", "val_acc")

    def test_step(self, batch, batch_idx):
        return self._eval_step("test", self.test_metrics, batch)

    def on_test_epoch_end(self):
        self._eval_epoch_end("test", self.test_metrics, "avg_test_loss", "avg_test_acc")

    def configure_optimizers(self):
        return torch.optim.Adam(self.parameters(), lr=self.args.learning_rate)

    @abstractmethod
    def _build_model(self):
        ...
