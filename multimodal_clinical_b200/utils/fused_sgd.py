"""``FusedHeadSGD``: torch.optim.SGD(momentum, weight_decay) semantics for the head parameters in ONE kernel launch
(the reference's optimiser is utils/BaseModel.py:275-285).  A regular ``torch.optim.Optimizer``
(param groups, ``lr`` visible to ``StepLR``, ``state_dict`` with ``momentum_buffer``), so it can stand in for the
SGD instance of ``configure_optimizers`` for parameter groups that hold only head tensors."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from .._lib import LfSgdArgs, check


class FusedHeadSGD(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, momentum=0.9, weight_decay=1.0e-4):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        lib = _lib.load()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            for i in range(0, len(ps), 8):
                chunk = ps[i:i + 8]
                a = LfSgdArgs()
                a.count = len(chunk)
                a.lr, a.momentum, a.weight_decay = float(group["lr"]), float(group["momentum"]), float(group["weight_decay"])
                firsts = set()
                for k, p in enumerate(chunk):
                    if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                        raise _lib.LfError("FusedHeadSGD needs contiguous fp32 CUDA parameters and gradients")
                    st = self.state[p]
                    firsts.add("momentum_buffer" not in st)
                    if "momentum_buffer" not in st:
                        st["momentum_buffer"] = torch.empty_like(p)
                    a.param[k], a.grad[k], a.momentum_buf[k], a.numel[k] = p.data_ptr(), p.grad.data_ptr(), st["momentum_buffer"].data_ptr(), p.numel()
                if len(firsts) != 1:
                    raise _lib.LfError("FusedHeadSGD: parameters of one launch must share their first-step state")
                a.first_step = int(firsts.pop())
                check(lib.lf_sgd_heads(C.byref(a), torch.cuda.current_stream().cuda_stream), "lf_sgd_heads")
        return loss


class SGDWithFusedHeads(torch.optim.SGD):
    """``torch.optim.SGD(params, lr, momentum, weight_decay)`` of the reference's ``configure_optimizers``
    (utils/BaseModel.py:275-285; enrico/joint_model.py:100-110; food101/joint_model_qmf.py:96-106) whose HEAD parameters are
    updated inside the fused step -- in the tail of the dW kernel, right after their gradients are reduced (and
    all-reduced on several GPUs) -- instead of by foreach kernels after ``backward()``.

    Every parameter stays in the optimizer's groups (so ``StepLR`` sees one ``lr`` and ``state_dict`` has every
    ``momentum_buffer``); ``step()`` skips the tensors the last fused step has already updated and runs the stock SGD
    on the rest (encoders, hidden MLP layers).  Heads fall back to the stock update whenever the step could not take
    them (exact-fp32 / narrow heads: the in-step update is part of the tensor-pipe dW kernel).  Not for training loops
    that modify ``.grad`` of the heads between ``backward()`` and ``step()`` (clipping): pass ``fused_head_sgd=False``."""

    def __init__(self, params, head, lr, momentum=0.9, weight_decay=1.0e-4):
        super().__init__(params, lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.head = head
        # the head reads this optimizer's current lr at every forward (StepLR changes it between steps)
        self.in_step = head.enable_in_step_sgd(lr, momentum, weight_decay, lr_source=lambda: self.param_groups[0]["lr"])

    @torch.no_grad()
    def step(self, closure=None):
        if not self.in_step:
            return super().step(closure)
        head = self.head
        done = head.in_step_updated()                          # {id(param): momentum buffer} of the step that just ran
        hidden = []
        for g in self.param_groups:
            for p in g["params"]:
                buf = done.get(id(p))
                if buf is None:
                    continue
                st = self.state[p]
                mine = st.get("momentum_buffer")
                if mine is not None and mine is not buf:       # loaded from a checkpoint: hand it to the step, then alias
                    buf.copy_(mine)
                st["momentum_buffer"] = buf
                hidden.append((p, p.grad))
                p.grad = None                                  # torch's SGD skips parameters without a gradient
        try:
            return super().step(closure)
        finally:
            for p, gr in hidden:
                p.grad = gr
