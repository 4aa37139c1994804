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
from ._lib import STAT
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
        ctx.ensemble = head.mode == "ensemble"
        if ctx.ensemble:
            # one CE per modality (cremad/ensemble_model_noised.py:52-53): batch means of the per-modality CE sums
            ce = (out.stats[STAT["CE_X1"]:STAT["CE_X2"] + 1] / float(out.batch_global)).float()
            res = (out.logits[0], out.logits[1], ce[0], ce[1])
            ctx.mark_non_differentiable(*res[:2])
            return res
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
        # d(total)/d(loss), a 0-d tensor; the ensemble head has one loss per modality, each reaching only its own head
        gm = (grads[2], grads[3]) if ctx.ensemble else (grads[3], grads[3])

        def scaled(t, dtype=None, m=0):
            if t is None:
                return None
            t = t * gm[m]
            return t if dtype is None or t.dtype == dtype else t.to(dtype)

        need = ctx.needs_input_grad
        return (None,
                scaled(out.dfeat[0], ctx.dtypes[0]) if need[1] else None,
                scaled(out.dfeat[1], ctx.dtypes[1], 1) if need[2] else None,
                scaled(out.dweight[0]) if need[3] else None, scaled(out.dbias[0]) if need[4] else None,
                scaled(out.dweight[1], None, 1) if need[5] else None, scaled(out.dbias[1], None, 1) if need[6] else None,
                None, None)


class FusedLateFusionHead(nn.Module):
    """Parameter-free module that owns the device state of the fused step (EMA, QMF History, scratch).

    mode: "jlogits" (mean fusion + CE; also what the OGM-GE models use), "qmf", or "ensemble" (one CE per modality,
    cremad/ensemble_model_noised.py: forward returns (x1_logits, x2_logits, x1_loss, x2_loss)).
    """

    def __init__(self, num_classes: int, mode: str = "jlogits", n_data: Optional[int] = None,
                 precision: str = "auto", ema_smoothing: float = 0.05, loss_terms: int = 0,
                 process_group=None, sharded: bool = False):
        super().__init__()
        if mode not in ("jlogits", "ogm_ge", "qmf", "ensemble"):
            raise NotImplementedError(f"fused head mode {mode!r}")
        if precision not in ("auto", "fp32", "tf32", "bf16"):
            raise ValueError(f"head precision {precision!r}")
        self.num_classes = int(num_classes)
        self.mode = mode
        self.n_data = n_data
        # "auto" follows what the trainer asked PyTorch for (utils/run_trainer.py:47, cremad/run_trainer.py:22), see
        # resolve_precision(); "fp32" / "tf32" / "bf16" pin the arithmetic of the three head GEMMs
        self.precision = precision
        # Batch-sharded (global-batch) semantics are OPT-IN: with sharded=False the head treats the batch it is
        # given as the whole batch even when torch.distributed is initialised, which is what a DDP wrapper
        # (Lightning strategy="auto" on several GPUs) expects -- DDP then averages every gradient, heads included.
        # sharded=True (optionally with a process group) makes the step global: statistics, History, EMA and head
        # gradients are exchanged inside the step and must NOT be reduced again by the caller.
        self.sharded = bool(sharded) or process_group is not None
        self.process_group = process_group
        self.ema_smoothing = ema_smoothing
        self.loss_terms = int(loss_terms)            # QMF loss ablations (LF_LOSS_* bits of include/lf_fusion.h)
        self.ogm_alpha: Optional[float] = None       # set by OGMGEBaseModel: coefficients come out of the same pass
        self.update_ema = True
        self.last_step: Optional[StepOutput] = None
        self._engine: Optional[LateFusionStep] = None
        self._engines = {}                           # (device, precision) -> engine; all share the EMA / History state
        self._grad_enabled = True
        self._ema = None
        self._qmf_state = None
        self._in_step_sgd = None                     # (lr, momentum, weight_decay) once an optimizer asked for it
        self._in_step_done = {}

    def bind_ema(self, ema) -> None:
        """Share the calibration state with the LightningModule's ``utils.EMA.EMA`` (utils/BaseModel.py:30)."""
        self._ema = ema
        self._engine = None
        self._engines = {}

    def bind_qmf(self, qmf) -> None:
        """Share the History arrays with the FusionNet's ``existing_algos.QMF.QMF`` object."""
        self._qmf_state = qmf._state
        self.n_data = qmf._state.n_data
        self._engine = None
        self._engines = {}

    def resolve_precision(self, f1: torch.Tensor) -> str:
        """Arithmetic of the head GEMMs for this call.  Explicit settings win; "auto" takes what the trainer set up:
        bf16 activations or an active bf16 CUDA autocast region (Trainer(precision="bf16-mixed"),
        utils/run_trainer.py:47) -> the bf16 tensor-pipe path; fp32 activations with
        torch.get_float32_matmul_precision() != "highest" (cremad/run_trainer.py:22 sets "medium") -> TF32;
        otherwise exact fp32.  Narrow heads (C < 32) are HBM-bound on the FMA pipe and always run exact fp32."""
        if self.num_classes < 32:
            if self.precision in ("bf16", "tf32"):
                raise _lib.LfError(f"precision {self.precision!r} is a tensor-pipe mode for wide heads (C >= 32); "
                                   f"C = {self.num_classes} runs the exact fp32 FMA path (use 'auto' or 'fp32')")
            return "fp32"
        if self.precision != "auto":
            return self.precision
        autocast_bf16 = torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16
        if (f1.dtype == torch.bfloat16 or autocast_bf16) and f1.shape[1] % 8 == 0:
            return "bf16"
        if f1.dtype == torch.float32 and torch.get_float32_matmul_precision() != "highest":
            return "tf32"
        return "fp32"

    # ------------------------------------------------------------------ SGD inside the step (utils/fused_sgd.py)
    def enable_in_step_sgd(self, lr: float, momentum: float, weight_decay: float, lr_source=None) -> bool:
        """Ask the fused step to apply torch.optim.SGD(lr, momentum, weight_decay) to the head tensors itself.  Returns
        False when that cannot be right: under a DDP wrapper (sharded=False while torch.distributed has several ranks)
        the gradients are only reduced after ``backward()``."""
        import torch.distributed as dist
        if not self.sharded and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return False
        self._in_step_sgd = (float(lr), float(momentum), float(weight_decay))
        self._lr_source = lr_source                  # callable -> current lr (the optimizer's param group, moved by StepLR)
        for eng in self._engines.values():
            self._apply_in_step_sgd(eng)
        return True

    def _apply_in_step_sgd(self, eng) -> None:
        if self._in_step_sgd is None:
            return
        try:
            eng.enable_sgd(*self._in_step_sgd)
        except _lib.LfError:
            pass                                     # exact-fp32 / narrow heads: the optimizer keeps updating them

    def set_in_step_lr(self, lr: float) -> None:
        if self._in_step_sgd is not None and float(lr) != self._in_step_sgd[0]:
            self._in_step_sgd = (float(lr),) + self._in_step_sgd[1:]
            for eng in self._engines.values():
                if eng._sgd is not None:
                    eng.set_lr(lr)

    def in_step_updated(self) -> dict:
        """{id(parameter): momentum buffer} for the head tensors the LAST fused step updated itself."""
        return self._in_step_done

    def _shared_state(self):
        if self._ema is None:
            from .utils.EMA import EMA
            self._ema = EMA(torch.zeros(2, self.num_classes), smoothing=self.ema_smoothing)
        if self.mode == "qmf" and self._qmf_state is None:
            from .existing_algos.QMF import _QmfState
            if self.n_data is None:
                raise ValueError("QMF head needs n_data (args.num_samples)")
            self._qmf_state = _QmfState(2, int(self.n_data))

    def _get_engine(self, device, precision: str) -> LateFusionStep:
        key = (str(device), precision)
        eng = self._engines.get(key)
        if eng is None:
            self._shared_state()
            eng = LateFusionStep(self.num_classes, mode=self.mode, n_data=self.n_data, device=device,
                                 precision=precision, ema_smoothing=self.ema_smoothing, qmf_state=self._qmf_state,
                                 ema=self._ema, loss_terms=self.loss_terms, process_group=self.process_group,
                                 sharded=self.sharded)
            eng.fresh_outputs = True
            self._apply_in_step_sgd(eng)
            self._engines[key] = eng
        self._engine = eng
        return eng

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
        eng = self._get_engine(f1.device, self.resolve_precision(f1))
        if self._in_step_sgd is not None and getattr(self, "_lr_source", None) is not None:
            self.set_in_step_lr(self._lr_source())   # device-resident hyper-parameter: a write only when StepLR moved it
        self._grad_enabled = torch.is_grad_enabled()
        self.update_ema = self.training and self._grad_enabled
        res = _FusedStep.apply(self, f1, f2, lin1.weight, lin1.bias, lin2.weight, lin2.bias, label,
                               idx.view(-1) if idx is not None else None)
        self._in_step_done = {}
        if eng._sgd is not None and self._grad_enabled and eng._sgd.get("mom") is not None:
            params = (lin1.weight, lin1.bias, lin2.weight, lin2.bias)
            self._in_step_done = {id(p): m for p, m in zip(params, eng._sgd["mom"])}
        return res
