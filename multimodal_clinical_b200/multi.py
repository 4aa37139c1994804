"""Mean fusion of M = 2..4 narrow heads with per-modality feature widths on ``lf_multi_heads_step`` (csrc/lf_multi.cu):
what the reference's three-modality (mustard/joint_model.py:72-83) and unequal-width (avmnist/joint_model.py:128-138)
``FusionNet.forward`` do after their encoders, forward and backward in one pass over the features.

``MultiHeadStep`` is the engine (device buffers + the C-ABI call); ``FusedMeanFusionHeads`` is the ``nn.Module`` face:
it takes the feature matrices and the ``nn.Linear`` heads (whose parameters keep the reference's state-dict names)
and returns ``(z_1, ..., z_M, avg_logits, loss)`` like the reference.  No eager fallback: CPU tensors raise ``LfError``."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from ._lib import LF_MAX_MODALITIES, LfMultiHeadsArgs, check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


@dataclass
class MultiStepOutput:
    logits: List[torch.Tensor]
    avg_logits: torch.Tensor
    loss: torch.Tensor            # 0-d
    dweight: List[torch.Tensor]
    dbias: List[torch.Tensor]
    dfeat: List[Optional[torch.Tensor]]
    stats: torch.Tensor           # fp64 [2 + M]: CE sum, joint hits, per-modality hits
    batch: int

    def accuracies(self) -> dict:
        st = self.stats.cpu()
        out = {"joint_acc": float(st[1]) / self.batch}
        for m in range(len(self.logits)):
            out[f"x{m + 1}_acc"] = float(st[2 + m]) / self.batch
        return out


class MultiHeadStep:
    def __init__(self, num_classes: int, device: Optional[torch.device] = None):
        self.lib = _lib.load()
        self.C = int(num_classes)
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise _lib.LfError("MultiHeadStep runs on a CUDA device only (there is no CPU fallback)")
        self._key = None
        self._ws = None

    def step(self, feats: Sequence[torch.Tensor], weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
             label: torch.Tensor, need_dfeat: bool = True) -> MultiStepOutput:
        M = len(feats)
        if not (2 <= M <= LF_MAX_MODALITIES) or len(weights) != M or len(biases) != M:
            raise _lib.LfError(f"mean fusion of {M} heads: 2..{LF_MAX_MODALITIES} modalities, one weight and bias each")
        B, Cn, dev = feats[0].shape[0], self.C, self.device
        dims = []
        for f, w, b in zip(feats, weights, biases):
            for t in (f, w, b):
                if not t.is_cuda or t.dtype != torch.float32:
                    raise _lib.LfError("lf_multi_heads_step takes fp32 CUDA tensors (no CPU / eager fallback)")
            if f.dim() != 2 or f.shape[0] != B or w.shape != (Cn, f.shape[1]) or b.shape != (Cn,):
                raise _lib.LfError(f"bad head shapes: feat {tuple(f.shape)} weight {tuple(w.shape)} bias {tuple(b.shape)}")
            dims.append(f.shape[1])
        feats = [f.contiguous() for f in feats]
        weights = [w.detach().contiguous() for w in weights]
        biases = [b.detach().contiguous() for b in biases]
        label = label.flatten().to(torch.int64).contiguous()
        key = (M, tuple(dims))
        if self._key != key:
            n = self.lib.lf_multi_heads_workspace_bytes(M, Cn, sum(dims))
            self._ws = torch.empty(n, dtype=torch.uint8, device=dev)
            self._key = key
        # fresh outputs every step: they escape to autograd and to metric lists
        logits = [torch.empty(B, Cn, device=dev) for _ in range(M)]
        avg = torch.empty(B, Cn, device=dev)
        dW = [torch.empty_like(w) for w in weights]
        db = [torch.empty_like(b) for b in biases]
        df = [torch.empty_like(f) if need_dfeat else None for f in feats]
        loss = torch.empty((), device=dev)
        stats = torch.empty(2 + M, dtype=torch.float64, device=dev)
        a = LfMultiHeadsArgs()
        a.modalities, a.batch, a.classes, a.need_dfeat = M, B, Cn, int(need_dfeat)
        for m in range(M):
            a.dim[m] = dims[m]
            a.feat[m], a.weight[m], a.bias[m] = feats[m].data_ptr(), weights[m].data_ptr(), biases[m].data_ptr()
            a.logits[m], a.dweight[m], a.dbias[m] = logits[m].data_ptr(), dW[m].data_ptr(), db[m].data_ptr()
            a.dfeat[m] = df[m].data_ptr() if need_dfeat else None
        a.label, a.avg_logits, a.loss_out, a.stats = label.data_ptr(), avg.data_ptr(), loss.data_ptr(), stats.data_ptr()
        a.workspace, a.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        check(self.lib.lf_multi_heads_step(C.byref(a), _stream()), "lf_multi_heads_step")
        return MultiStepOutput(logits, avg, loss, dW, db, df, stats, B)


class _MultiStepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, head: "FusedMeanFusionHeads", label, M, *tensors):
        feats, weights, biases = tensors[:M], tensors[M:2 * M], tensors[2 * M:]
        need_dfeat = any(ctx.needs_input_grad[3 + m] for m in range(M))
        out = head.engine(feats[0].device).step(feats, weights, biases, label, need_dfeat=need_dfeat)
        head.last_step = out
        ctx.out, ctx.M = out, M
        res = tuple(out.logits) + (out.avg_logits, out.loss)
        ctx.mark_non_differentiable(*res[:-1])       # nothing but the loss is ever back-propagated in the reference
        return res

    @staticmethod
    def backward(ctx, *grads):
        out, M, g = ctx.out, ctx.M, grads[-1]
        need = ctx.needs_input_grad
        pick = lambda ts, base: tuple((ts[m] * g if (need[base + m] and ts[m] is not None) else None) for m in range(M))
        return (None, None, None) + pick(out.dfeat, 3) + pick(out.dweight, 3 + M) + pick(out.dbias, 3 + 2 * M)


class FusedMeanFusionHeads(nn.Module):
    """Parameter-free: ``forward(feats, heads, label) -> (z_1, ..., z_M, avg_logits, loss)`` with ``heads`` the ``nn.Linear``
    classifiers of the modalities (mustard: ``x{1,2,3}_model.fc3``; avmnist: ``classifier_x{1,2}``)."""

    def __init__(self, num_classes: int):
        super().__init__()
        self.num_classes = int(num_classes)
        self._engines = {}
        self.last_step: Optional[MultiStepOutput] = None

    def engine(self, device) -> MultiHeadStep:
        key = str(device)
        if key not in self._engines:
            self._engines[key] = MultiHeadStep(self.num_classes, device=device)
        return self._engines[key]

    def forward(self, feats: Sequence[torch.Tensor], heads: Sequence[nn.Linear], label: torch.Tensor):
        if not feats[0].is_cuda:
            raise _lib.LfError("FusedMeanFusionHeads needs CUDA tensors: this path has no CPU / eager fallback")
        M = len(feats)
        feats = [f.float() for f in feats]
        return _MultiStepFn.apply(self, label, M, *feats, *[h.weight for h in heads], *[h.bias for h in heads])
