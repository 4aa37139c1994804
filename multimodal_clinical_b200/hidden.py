"""The hidden layers of the Food101 per-modality MLPs (food101/joint_model_qmf.py:12-26: ``Linear -> ReLU -> Dropout(0.2)``,
twice) on ``lf_hidden_forward`` / ``lf_hidden_backward`` (csrc/lf_hidden.cu): both modalities' layers of one shape per
call, bias + ReLU + dropout in the GEMM epilogue, no stored mask.

``FusedHiddenPair`` is parameter-free; it takes the two ``nn.Linear`` modules (state-dict names unchanged) and the two
inputs.  Precision follows the trainer like the fused head does: bf16 inputs or an active bf16 CUDA autocast region ->
``bf16`` (tensors in HBM are bf16, like autocast's Linear output); fp32 with ``float32_matmul_precision != 'highest'`` ->
``tf32``; else exact fp32 (3xTF32).  The dropout stream is Philox keyed by ``torch.initial_seed()`` with a per-layer, per-call
counter, so a seeded run is reproducible; it is not torch's own dropout stream (statistical, not bit, parity in training
mode -- the same status as the OGM-GE noise).  No eager fallback: CPU tensors raise ``LfError``."""
from __future__ import annotations

import ctypes as C
import itertools

import torch
import torch.nn as nn

from . import _lib
from ._lib import LF_PREC_BF16, LF_PREC_FP32, LF_PREC_TF32, LfHiddenArgs, check

_layer_ids = itertools.count(1)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _HiddenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layer: "FusedHiddenPair", x1, x2, w1, b1, w2, b2):
        prec = layer.resolve_precision(x1)
        dt = torch.bfloat16 if prec == LF_PREC_BF16 else torch.float32
        xs = [x1.detach().to(dt).contiguous(), x2.detach().to(dt).contiguous()]
        B, Din = xs[0].shape
        Dout = w1.shape[0]
        a = layer._args(B, Din, Dout, prec, xs, (w1, w2), (b1, b2))
        hs = [torch.empty(B, Dout, device=x1.device, dtype=dt) for _ in range(2)]
        for m in range(2):
            a.h[m] = hs[m].data_ptr()
        check(layer.lib.lf_hidden_forward(C.byref(a), _stream()), "lf_hidden_forward")
        ctx.layer, ctx.xs, ctx.hs, ctx.rng = layer, xs, hs, (a.training, a.drop_p, a.seed, a.offset, prec)
        ctx.params = (w1, b1, w2, b2)
        ctx.in_dtypes = (x1.dtype, x2.dtype)
        return hs[0], hs[1]

    @staticmethod
    def backward(ctx, dh1, dh2):
        layer, xs, hs = ctx.layer, ctx.xs, ctx.hs
        training, p, seed, offset, prec = ctx.rng
        w1, b1, w2, b2 = ctx.params
        dt = hs[0].dtype
        B, Din = xs[0].shape
        Dout = w1.shape[0]
        a = layer._args(B, Din, Dout, prec, xs, (w1, w2), (b1, b2), rng=(training, p, seed, offset))
        dhs = [g.to(dt).contiguous() for g in (dh1, dh2)]
        dpre = [torch.empty_like(h) for h in hs]
        need_dx = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        dxs = [torch.empty_like(x) for x in xs] if need_dx else [None, None]
        dW = [torch.empty_like(w1, dtype=torch.float32), torch.empty_like(w2, dtype=torch.float32)]
        db = [torch.empty_like(b1, dtype=torch.float32), torch.empty_like(b2, dtype=torch.float32)]
        for m in range(2):
            a.h[m], a.dh[m], a.dpre[m] = hs[m].data_ptr(), dhs[m].data_ptr(), dpre[m].data_ptr()
            a.dx[m] = dxs[m].data_ptr() if need_dx else None
            a.dweight[m], a.dbias[m] = dW[m].data_ptr(), db[m].data_ptr()
        check(layer.lib.lf_hidden_backward(C.byref(a), _stream()), "lf_hidden_backward")
        dx = [dxs[m].to(ctx.in_dtypes[m]) if (need_dx and ctx.needs_input_grad[1 + m]) else None for m in range(2)]
        return None, dx[0], dx[1], dW[0], db[0], dW[1], db[1]


class FusedHiddenPair(nn.Module):
    """``forward(x1, x2, linear1, linear2) -> (h1, h2)`` with ``h_m = dropout_p(relu(linear_m(x_m)))``."""

    def __init__(self, drop_p: float = 0.2, precision: str = "auto"):
        super().__init__()
        if precision not in ("auto", "fp32", "tf32", "bf16"):
            raise ValueError(f"hidden-layer precision {precision!r}")
        self.drop_p = float(drop_p)
        self.precision = precision
        self.layer_id = next(_layer_ids)
        self.calls = 0
        self.lib = None
        self._ws = None
        self._ws_key = None

    def resolve_precision(self, x: torch.Tensor) -> int:
        if self.precision != "auto":
            return {"fp32": LF_PREC_FP32, "tf32": LF_PREC_TF32, "bf16": LF_PREC_BF16}[self.precision]
        if x.dtype == torch.bfloat16 or (torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16):
            return LF_PREC_BF16
        return LF_PREC_TF32 if torch.get_float32_matmul_precision() != "highest" else LF_PREC_FP32

    def _args(self, B, Din, Dout, prec, xs, ws, bs, rng=None) -> LfHiddenArgs:
        if self.lib is None:
            self.lib = _lib.load()
        key = (B, Din, Dout, str(xs[0].device))
        if self._ws_key != key:
            self._ws = torch.empty(self.lib.lf_hidden_workspace_bytes(B, Din, Dout), dtype=torch.uint8, device=xs[0].device)
            self._ws_key = key
        a = LfHiddenArgs()
        a.batch, a.dim_in, a.dim_out, a.precision = B, Din, Dout, prec
        if rng is None:
            self.calls += 1
            rng = (int(self.training), self.drop_p, torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, (self.layer_id << 40) + self.calls)
        a.training, a.drop_p, a.seed, a.offset = rng
        for m in range(2):
            w, b = ws[m].detach(), bs[m].detach()
            if not (w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and b.is_contiguous()):
                raise _lib.LfError("lf_hidden needs contiguous fp32 CUDA weights")
            a.x[m], a.weight[m], a.bias[m] = xs[m].data_ptr(), w.data_ptr(), b.data_ptr()
        a.workspace, a.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        return a

    def forward(self, x1, x2, linear1: nn.Linear, linear2: nn.Linear):
        if not x1.is_cuda:
            raise _lib.LfError("FusedHiddenPair needs CUDA tensors: this path has no CPU / eager fallback")
        if x1.dim() != 2 or x1.shape != x2.shape or linear1.weight.shape != linear2.weight.shape:
            raise _lib.LfError("FusedHiddenPair: the two modalities' layers must have one shape")
        return _HiddenFn.apply(self, x1, x2, linear1.weight, linear1.bias, linear2.weight, linear2.bias)
