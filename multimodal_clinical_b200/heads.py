"""The fused late-fusion head behind the reference's ``FusionNet.forward`` contract.

``FusedLateFusionHead`` takes the two pooled feature matrices and the two ``nn.Linear`` classifier
modules (whose parameters keep the reference's state-dict names) and returns exactly what the
reference's ``FusionNet.forward`` returns after the encoders:

    jlogits / ogm_ge : (x1_logits, x2_logits, avg_logits, loss)              cremad/joint_model_ogm_ge.py:50-58
    qmf              : (x1_logits, x2_logits, avg_logits, loss, logits_df)   cremad/joint_model_qmf.py:57-75

Forward AND backward of the head run inside the forward call (one pass of the CUDA step); the autograd
node only hands the stashed ``dfeat`` / ``dW`` / ``db`` back, scaled by the incoming gradient of the
loss.  The logits outputs are marked non-differentiable: in the reference nothing but ``loss`` is ever
back-propagated (utils/BaseModel.py:59-112, 869-875).

There is no eager fallback: CPU tensors or a missing extension raise ``LfError``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .step import LateFusionStep, StepOutput


class _FusedStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, head: "FusedLateFusionHead", f1, f2, w1, b1, w2, b2, label, idx):
        need_dfeat = bool(ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        grad_mode = head._grad_enabled
        out = head.engine.step([f1, f2], [w1, w2], [b1, b2], label, idx=idx, need_dfeat=need_dfeat and grad_mode,
                               update_ema=head.update_ema, ogm_alpha=head.ogm_alpha, backward=grad_mode)
        head.last_step = out
        ctx.out = out if grad_mode else None
        ctx.dtypes = (f1.dtype, f2.dtype)
        res = (out.logits[0], out.logits[1], out.avg_logits, out.loss)
        if out.logits_df is not None:
            res = res + (out.logits_df,)
        ctx.mark_non_differentiable(*(res[:3] + res[4:]))
        return res

    @staticmethod
    def backward(ctx, *grads):
        out: StepOutput = ctx.out
        if out is None:
            raise _lib.LfError("backward through a fused step that ran with gradients disabled")
        g = grads[3]                                    # d(total)/d(loss), a 0-d tensor

        def scaled(t, dtype=None):
            if t is None:
                return None
            t = t * g
            return t if dtype is None or t.dtype == dtype else t.to(dtype)

        need = ctx.needs_input_grad
        return (None,
                scaled(out.dfeat[0], ctx.dtypes[0]) if need[1] else None,
                scaled(out.dfeat[1], ctx.dtypes[1]) if need[2] else None,
                scaled(out.dweight[0]) if need[3] else None, scaled(out.dbias[0]) if need[4] else None,
                scaled(out.dweight[1]) if need[5] else None, scaled(out.dbias[1]) if need[6] else None,
                None, None)


class FusedLateFusionHead(nn.Module):
    """Parameter-free module that owns the device state of the fused step (EMA, QMF History, scratch).

    mode: "jlogits" (mean fusion + CE; also what the OGM-GE models use) or "qmf".
    """

    def __init__(self, num_classes: int, mode: str = "jlogits", n_data: Optional[int] = None,
                 precision: str = "fp32", ema_smoothing: float = 0.05, loss_terms: int = 0):
        super().__init__()
        if mode not in ("jlogits", "ogm_ge", "qmf"):
            raise NotImplementedError(f"fused head mode {mode!r}")
        self.num_classes = int(num_classes)
        self.mode = mode
        self.n_data = n_data
        self.precision = precision
        self.ema_smoothing = ema_smoothing
        self.loss_terms = int(loss_terms)            # QMF loss ablations (LF_LOSS_* bits of include/lf_fusion.h)
        self.ogm_alpha: Optional[float] = None       # set by OGMGEBaseModel: coefficients come out of the same pass
        self.update_ema = True
        self.last_step: Optional[StepOutput] = None
        self._engine: Optional[LateFusionStep] = None
        self._grad_enabled = True
        self._ema = None
        self._qmf_state = None

    def bind_ema(self, ema) -> None:
        """Share the calibration state with the LightningModule's ``utils.EMA.EMA`` (utils/BaseModel.py:30)."""
        self._ema = ema
        self._engine = None

    def bind_qmf(self, qmf) -> None:
        """Share the History arrays with the FusionNet's ``existing_algos.QMF.QMF`` object."""
        self._qmf_state = qmf._state
        self.n_data = qmf._state.n_data
        self._engine = None

    def _get_engine(self, device) -> LateFusionStep:
        if self._engine is None or self._engine.device != device:
            self._engine = LateFusionStep(self.num_classes, mode=self.mode, n_data=self.n_data, device=device,
                                          precision=self.precision, ema_smoothing=self.ema_smoothing,
                                          qmf_state=self._qmf_state, ema=self._ema, loss_terms=self.loss_terms)
            self._engine.fresh_outputs = True
        return self._engine

    @property
    def engine(self) -> LateFusionStep:
        if self._engine is None:
            raise _lib.LfError("the fused head has not run yet")
        return self._engine

    def forward(self, f1: torch.Tensor, f2: torch.Tensor, lin1: nn.Linear, lin2: nn.Linear, label: torch.Tensor,
                idx: Optional[torch.Tensor] = None):
        if not f1.is_cuda:
            raise _lib.LfError("the fused late-fusion head runs on CUDA only (sm_100a); got a CPU tensor")
        if self.mode == "qmf" and idx is None:
            raise ValueError("QMF head needs the dataset indices of the batch (idx)")
        # the reference's eval steps still compute the loss (and, QMF, mutate the History) but never touch
        # the EMA and need no gradients (utils/BaseModel.py:133-160, 1009-1040)
        self._get_engine(f1.device)
        self._grad_enabled = torch.is_grad_enabled()
        self.update_ema = self.training and self._grad_enabled
        return _FusedStep.apply(self, f1, f2, lin1.weight, lin1.bias, lin2.weight, lin2.bias, label,
                                idx.view(-1) if idx is not None else None)
