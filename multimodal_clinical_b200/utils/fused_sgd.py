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
