"""Lightning base classes of the late-fusion models, running the per-batch step on the fused CUDA path.

Counterpart of the reference's utils/BaseModel.py for the three classes on the hot path:
``JointLogitsBaseModel`` (:15-289), ``OGMGEBaseModel`` (:797-912, manual optimisation) and
``QMFBaseModel`` (:914-1162).  Same constructor (``args``), abstract ``_build_model()``, hooks, batch
tuples, public attributes (``args, model, num_modality, ema_offset, train_metrics/val_metrics/test_metrics,
ogm_modulation, ogm_alpha, automatic_optimization``) and metric keys (``train_step/*``, ``train_epoch/*``,
``val_epoch/val_avg_acc`` ...).

What changes underneath: ``self.model`` is a FusionNet whose head is a ``FusedLateFusionHead``; one call
of it produces the logits, the loss, the head / feature gradients, the EMA update and every accuracy
count of the step on the device.  ``training_step`` therefore reads its metrics
from the packed statistics of that step instead of re-deriving them with ~20 small kernels and four
``.item()`` syncs (utils/BaseModel.py:78-108); per-step metric lists hold 0-d device tensors and are
reduced once per epoch.  Ensemble / JointProb base classes are outside the path.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch
from torch.optim.lr_scheduler import StepLR

from .._lib import STAT
from ..existing_algos.OGM_GE import ogm_ge
from ..heads import FusedLateFusionHead
from .EMA import EMA
from .fused_sgd import SGDWithFusedHeads
from .lightning_compat import pl


_epoch_ws = {}


def epoch_offset_correction(logits: torch.Tensor, labels: torch.Tensor):
    """Epoch-end unimodal offset correction (utils/BaseModel.py:168-185) on the device: logits (N, 2, C), labels (N) ->
    (offset (2, C), accuracies [x1 uncal, x2 uncal, x1 corrected, x2 corrected] as a 4-element fp32 device tensor)."""
    import ctypes as C
    from .. import _lib
    if not logits.is_cuda:
        raise _lib.LfError("epoch_offset_correction runs on CUDA only; got a CPU tensor")
    lib = _lib.load()
    z = logits.detach().float().contiguous()
    y = labels.to(device=z.device, dtype=torch.int64).contiguous()
    n, m, c = z.shape
    if m != 2:
        raise NotImplementedError("two modalities")
    key = (z.device.index, c)
    if key not in _epoch_ws:
        _epoch_ws[key] = torch.zeros(lib.lf_epoch_workspace_bytes(c), dtype=torch.uint8, device=z.device)
    ws = _epoch_ws[key]
    offset = torch.empty(2, c, device=z.device)
    acc = torch.empty(4, dtype=torch.float64, device=z.device)
    _lib.check(lib.lf_epoch_offset_correction(z.data_ptr(), y.data_ptr(), n, c, offset.data_ptr(), acc.data_ptr(), ws.data_ptr(),
                                              ws.numel(), torch.cuda.current_stream().cuda_stream), "lf_epoch_offset_correction")
    return offset, acc.float()


def _mean(xs):
    return torch.stack([torch.as_tensor(x, dtype=torch.float32).reshape(()) for x in xs]).mean()


class JointLogitsBaseModel(pl.LightningModule, ABC):
    _train_lists = ("train_loss", "train_acc", "train_x1_acc_uncal", "train_x2_acc_uncal", "train_x1_acc", "train_x2_acc")

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.model = self._build_model()
        self.num_modality = 2
        self.ema_offset = EMA(torch.zeros(self.num_modality, self.args.num_classes))
        head = getattr(self.model, "fused", None)
        if not isinstance(head, FusedLateFusionHead):
            raise TypeError("self.model must expose a FusedLateFusionHead as `.fused`: the late-fusion step has no "
                            "eager PyTorch fallback")
        head.bind_ema(self.ema_offset)
        self.train_metrics = {"train_loss": [], "train_acc": [], "train_logits": [], "train_x1_acc_uncal": [],
                              "train_x2_acc_uncal": [], "train_x1_acc": [], "train_x2_acc": []}
        self.val_metrics = {"val_loss": [], "val_acc": [], "val_logits": [], "val_labels": []}
        self.test_metrics = {"test_loss": [], "test_acc": [], "test_logits": [], "test_labels": []}

    def forward(self, x1, x2, label):
        return self.model(x1, x2, label)

    # ------------------------------------------------------------------ shared pieces
    def _step_accuracies(self):
        """Accuracy tensors of the step that just ran, from its packed statistics (no host sync):
        counts / global batch for x1/x2 uncalibrated, joint, df, x1/x2 calibrated (utils/BaseModel.py:78-92)."""
        out = self.model.fused.last_step
        a = (out.stats[STAT["CNT_X1"]:STAT["CNT_X2_CAL"] + 1] / float(out.batch_global)).float()
        return {"x1_uncal": a[0], "x2_uncal": a[1], "joint": a[2], "df": a[3], "x1_cal": a[4], "x2_cal": a[5]}

    def _log_train_step(self, loss, acc, with_df=False):
        kw = dict(on_step=True, on_epoch=True, prog_bar=False, logger=True)
        self.log("train_step/train_loss", loss, **kw)
        self.log("train_step/train_acc", acc["joint"], **kw)
        self.log("train_step/train_x1_acc", acc["x1_cal"], **kw)
        self.log("train_step/train_x2_acc", acc["x2_cal"], **kw)
        self.log("train_step/train_x1_uncal_acc", acc["x1_uncal"], **kw)
        self.log("train_step/train_x2_uncal_acc", acc["x2_uncal"], **kw)
        if with_df:
            self.log("train_step/train_df_acc", acc["df"], **kw)
        m = self.train_metrics
        m["train_acc"].append(acc["joint"]); m["train_loss"].append(loss.detach())
        m["train_x1_acc_uncal"].append(acc["x1_uncal"]); m["train_x2_acc_uncal"].append(acc["x2_uncal"])
        m["train_x1_acc"].append(acc["x1_cal"]); m["train_x2_acc"].append(acc["x2_cal"])
        if with_df:
            m["train_df_acc"].append(acc["df"])

    def _train_epoch_end(self, with_df=False):
        kw = dict(on_step=False, on_epoch=True, prog_bar=False, logger=True)
        m = self.train_metrics
        self.log("train_epoch/train_avg_acc", _mean(m["train_acc"]), **kw)
        self.log("train_epoch/train_avg_loss", _mean(m["train_loss"]), **kw)
        self.log("train_epoch/train_avg_x1_acc_uncal", _mean(m["train_x1_acc_uncal"]), **kw)
        self.log("train_epoch/train_avg_x2_acc_uncal", _mean(m["train_x2_acc_uncal"]), **kw)
        self.log("train_epoch/train_avg_x1_acc", _mean(m["train_x1_acc"]), **kw)
        self.log("train_epoch/train_avg_x2_acc", _mean(m["train_x2_acc"]), **kw)
        if with_df:
            self.log("train_epoch/train_avg_df_acc", _mean(m["train_df_acc"]), **kw)
        for k in self._train_lists + (("train_df_acc",) if with_df else ()):
            m[k].clear()

    def _eval_step(self, kind, batch, with_df=False):
        metrics = self.val_metrics if kind == "val" else self.test_metrics
        if with_df:
            x1, x2, label, idx = batch
            x1_logits, x2_logits, avg_logits, loss, logits_df = self.model(x1, x2, label, idx)
        else:
            x1, x2, label = batch
            x1_logits, x2_logits, avg_logits, loss = self.model(x1, x2, label)
        acc = self._step_accuracies()
        kw = dict(on_step=True, on_epoch=True, prog_bar=False, logger=True)
        self.log(f"{kind}_step/{kind}_acc", acc["joint"], **kw)
        self.log(f"{kind}_step/{kind}_loss", loss, **kw)
        if with_df:
            self.log(f"{kind}_step/logits_df_acc", acc["df"], **kw)          # the reference's key (utils/BaseModel.py:1034, 1110)
            metrics[f"{kind}_df_acc"].append(acc["df"])
        metrics[f"{kind}_logits"].append(torch.stack((x1_logits, x2_logits), dim=1))
        metrics[f"{kind}_labels"].append(label)
        metrics[f"{kind}_loss"].append(loss.detach())
        metrics[f"{kind}_acc"].append(acc["joint"])
        return loss

    def _eval_epoch_end(self, kind, with_df=False):
        """Epoch-end unimodal offset correction over all collected logits (utils/BaseModel.py:168-202):
        once per epoch, outside the per-batch step."""
        metrics = self.val_metrics if kind == "val" else self.test_metrics
        labels = torch.cat(metrics[f"{kind}_labels"], dim=0)
        logits = torch.cat(metrics[f"{kind}_logits"], dim=0)          # (N, M, C)
        offset, acc = epoch_offset_correction(logits, labels)         # device kernel pair, no host sync
        self.last_epoch_offset = offset                                # (M, C): mean_m(mean_n logits) - mean_n logits
        kw = dict(on_step=False, on_epoch=True, prog_bar=False, logger=True)
        self.log(f"{kind}_epoch/{kind}_avg_acc", _mean(metrics[f"{kind}_acc"]), **kw)
        self.log(f"{kind}_epoch/{kind}_avg_loss", _mean(metrics[f"{kind}_loss"]), **kw)
        self.log(f"{kind}_epoch/{kind}_avg_x1_acc_uncal", acc[0], **kw)
        self.log(f"{kind}_epoch/{kind}_avg_x2_acc_uncal", acc[1], **kw)
        self.log(f"{kind}_epoch/{kind}_avg_x1_acc", acc[2], **kw)
        self.log(f"{kind}_epoch/{kind}_avg_x2_acc", acc[3], **kw)
        if with_df:
            self.log(f"{kind}_epoch/{kind}_avg_df_acc", _mean(metrics[f"{kind}_df_acc"]), **kw)
        for k in list(metrics):
            metrics[k].clear()

    # ------------------------------------------------------------------ Lightning hooks
    def training_step(self, batch, batch_idx):
        x1, x2, label = batch
        x1_logits, x2_logits, avg_logits, loss = self.model(x1, x2, label)
        self._log_train_step(loss, self._step_accuracies())
        return loss

    def on_train_epoch_end(self) -> None:
        self._train_epoch_end()

    def validation_step(self, batch, batch_idx):
        return self._eval_step("val", batch)

    def on_validation_epoch_end(self) -> None:
        self._eval_epoch_end("val")

    def test_step(self, batch, batch_idx):
        return self._eval_step("test", batch)

    def on_test_epoch_end(self):
        self._eval_epoch_end("test")

    def _sgd(self):
        """SGD(lr, momentum 0.9, weight decay 1e-4) over all parameters (utils/BaseModel.py:275-285).  The head tensors are
        updated inside the fused step where it can take them (utils/fused_sgd.py; ``args.fused_head_sgd = False`` keeps the
        stock optimizer for everything)."""
        if getattr(self.args, "fused_head_sgd", True):
            return SGDWithFusedHeads(self.parameters(), self.model.fused, lr=self.args.learning_rate, momentum=0.9, weight_decay=1.0e-4)
        return torch.optim.SGD(self.parameters(), lr=self.args.learning_rate, momentum=0.9, weight_decay=1.0e-4)

    def configure_optimizers(self):
        optimizer = self._sgd()
        if self.args.use_scheduler:
            scheduler = {'scheduler': StepLR(optimizer, step_size=70, gamma=0.1), 'interval': 'epoch', 'frequency': 1}
            return [optimizer], [scheduler]
        return optimizer

    @abstractmethod
    def _build_model(self):
        pass


class EnsembleBaseModel(pl.LightningModule, ABC):
    """Per-modality CE ensemble (utils/BaseModel.py:291-562 of the reference): ``self.model(x1, x2, label)`` returns
    ``(x1_logits, x2_logits, x1_loss, x2_loss)``; no EMA calibration and no epoch-end offset correction in this family.
    The accuracies of a step come from the packed statistics of the fused step (no host sync)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.model = self._build_model()
        if not isinstance(getattr(self.model, "fused", None), FusedLateFusionHead):
            raise TypeError("self.model must expose a FusedLateFusionHead as `.fused`: the late-fusion step has no "
                            "eager PyTorch fallback")
        self.train_metrics = {"train_loss": [], "train_acc": [], "train_x1_acc": [], "train_x2_acc": []}
        self.val_metrics = {"val_loss": [], "val_acc": [], "val_x1_acc": [], "val_x2_acc": []}
        self.test_metrics = {"test_loss": [], "test_acc": [], "test_x1_acc": [], "test_x2_acc": []}

    def forward(self, x1, x2, label):
        return self.model(x1, x2, label)

    def _accs(self):
        out = self.model.fused.last_step
        a = (out.stats[STAT["CNT_X1"]:STAT["CNT_JOINT"] + 1] / float(out.batch_global)).float()
        return a[0], a[1], a[2]                      # x1, x2, joint = argmax of (z1 + z2) / 2

    def _record(self, kind, loss, accs, metrics):
        x1_acc, x2_acc, joint_acc = accs
        kw = dict(on_step=True, on_epoch=True, prog_bar=False, logger=True)
        self.log(f"{kind}_step/{kind}_loss", loss, **kw)
        self.log(f"{kind}_step/{kind}_acc", joint_acc, **kw)
        if kind == "train":
            self.log("train_step/train_x1_acc", x1_acc, **kw)
            self.log("train_step/train_x2_acc", x2_acc, **kw)
        metrics[f"{kind}_loss"].append(loss.detach()); metrics[f"{kind}_acc"].append(joint_acc)
        metrics[f"{kind}_x1_acc"].append(x1_acc); metrics[f"{kind}_x2_acc"].append(x2_acc)

    def training_step(self, batch, batch_idx):
        x1, x2, label = batch
        x1_logits, x2_logits, x1_loss, x2_loss = self.model(x1, x2, label)
        avg_loss = (x1_loss + x2_loss)               # (sic: the base class sums, utils/BaseModel.py:361)
        self._record("train", avg_loss, self._accs(), self.train_metrics)
        return avg_loss

    def _eval_step(self, kind, batch):
        x1, x2, label = batch
        x1_logits, x2_logits, x1_loss, x2_loss = self.model(x1, x2, label)
        avg_loss = (x1_loss + x2_loss) / 2
        self._record(kind, avg_loss, self._accs(), self.val_metrics if kind == "val" else self.test_metrics)
        return avg_loss

    def _epoch_end(self, kind, metrics):
        kw = dict(on_step=False, on_epoch=True, prog_bar=False, logger=True)
        self.log(f"{kind}_epoch/{kind}_avg_loss", _mean(metrics[f"{kind}_loss"]), **kw)
        self.log(f"{kind}_epoch/{kind}_avg_acc", _mean(metrics[f"{kind}_acc"]), **kw)
        self.log(f"{kind}_epoch/{kind}_avg_x1_acc", _mean(metrics[f"{kind}_x1_acc"]), **kw)
        self.log(f"{kind}_epoch/{kind}_avg_x2_acc", _mean(metrics[f"{kind}_x2_acc"]), **kw)
        for k in metrics:
            metrics[k].clear()

    def on_train_epoch_end(self) -> None:
        self._epoch_end("train", self.train_metrics)

    def validation_step(self, batch, batch_idx):
        return self._eval_step("val", batch)

    def on_validation_epoch_end(self) -> None:
        self._epoch_end("val", self.val_metrics)

    def test_step(self, batch, batch_idx):
        return self._eval_step("test", batch)

    def on_test_epoch_end(self):
        self._epoch_end("test", self.test_metrics)

    _sgd = JointLogitsBaseModel._sgd
    configure_optimizers = JointLogitsBaseModel.configure_optimizers

    @abstractmethod
    def _build_model(self):
        pass


class OGMGEBaseModel(JointLogitsBaseModel, ABC):
    """Manual optimisation with OGM-GE modulation of the encoders' conv gradients (utils/BaseModel.py:797-912):
    zero_grad -> manual_backward -> ogm_ge -> step, scheduler stepped by hand per epoch."""

    def __init__(self, args):
        super().__init__(args)
        self.automatic_optimization = False
        self.ogm_modulation = self.args.grad_mod_type
        self.ogm_alpha = self.args.alpha

    def training_step(self, batch, batch_idx):
        x1, x2, label = batch
        x1_logits, x2_logits, avg_logits, loss = self.model(x1, x2, label)
        self._log_train_step(loss, self._step_accuracies())
        opt = self.optimizers()
        opt.zero_grad()
        self.manual_backward(loss)
        if self.ogm_modulation:
            ogm_ge(self.model, x1_logits, x2_logits, label, modulation=self.ogm_modulation, alpha=self.ogm_alpha)
        opt.step()
        return loss

    def on_train_epoch_end(self) -> None:
        self._train_epoch_end()
        if self.args.use_scheduler:
            schedulers = self.lr_schedulers()
            if not isinstance(schedulers, list):
                schedulers = [schedulers]
            for scheduler in schedulers:
                scheduler.step()

    @abstractmethod
    def _build_model(self):
        pass


class QMFBaseModel(JointLogitsBaseModel, ABC):
    """Four-tuple batches ``(x1, x2, label, idx)`` and the extra ``df_acc`` metric (utils/BaseModel.py:914-1162).
    Validation / test batches update the QMF History with their indices, as in the reference (:1023-1026)."""

    def __init__(self, args):
        super().__init__(args)
        self.train_metrics.update({"train_df_acc": []})
        self.val_metrics.update({"val_df_acc": []})
        self.test_metrics.update({"test_df_acc": []})

    def forward(self, x1, x2, label, idx):
        return self.model(x1, x2, label, idx)

    def training_step(self, batch, batch_idx):
        x1, x2, label, idx = batch
        x1_logits, x2_logits, avg_logits, loss, logits_df = self.model(x1, x2, label, idx)
        self._log_train_step(loss, self._step_accuracies(), with_df=True)
        return loss

    def on_train_epoch_end(self) -> None:
        self._train_epoch_end(with_df=True)

    def validation_step(self, batch, batch_idx):
        return self._eval_step("val", batch, with_df=True)

    def on_validation_epoch_end(self) -> None:
        self._eval_epoch_end("val", with_df=True)

    def test_step(self, batch, batch_idx):
        return self._eval_step("test", batch, with_df=True)

    def on_test_epoch_end(self):
        self._eval_epoch_end("test", with_df=True)

    @abstractmethod
    def _build_model(self):
        pass
