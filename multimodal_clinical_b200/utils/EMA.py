"""EMA logit-offset calibration state (utils/EMA.py of the reference), kept on the device.

The reference keeps ``x`` on the CPU and pays one D2H + two H2D copies per step (utils/EMA.py:33,
utils/BaseModel.py:84-85).  Here ``x`` and ``offset`` are (M, C) CUDA tensors updated by ``lf_ema_update``;
the fused head updates them inside the step (``bind``), ``update()`` is the stand-alone call.
"""
from __future__ import annotations

import torch

from .. import _lib
from .._lib import check


class EMA:
    def __init__(self, x0, smoothing=0.05):
        self.x = x0
        self.smoothing = smoothing
        self.counter = 0
        self._offset = None
        self._stats = None

    def to(self, device) -> "EMA":
        device = torch.device(device)
        if self.x.device != device or self.x.dtype != torch.float32 or self._offset is None:
            self.x = self.x.to(device=device, dtype=torch.float32).contiguous()
            self._offset = (torch.mean(self.x, dim=0, keepdim=True) - self.x).contiguous()   # one-time init
        return self

    def update(self, x_new):
        """x <- smoothing * x_new + (1 - smoothing) * x   (utils/EMA.py:29-34); x_new: (M, C) batch-mean logits."""
        if not x_new.is_cuda:
            raise _lib.LfError("EMA.update needs a CUDA tensor; there is no CPU path")
        if self.x.shape[0] != 2:
            raise NotImplementedError("EMA kernels support two modalities")
        self.to(x_new.device)
        Cn = self.x.shape[1]
        if self._stats is None or self._stats.device != x_new.device:
            self._stats = torch.zeros(_lib.LF_STATS_HEADER + 2 * Cn, dtype=torch.float64, device=x_new.device)
        self._stats[_lib.LF_STATS_HEADER:] = x_new.detach().reshape(-1).to(torch.float64)   # "sums" with batch 1
        check(_lib.load().lf_ema_update(self.x.data_ptr(), self._offset.data_ptr(), self._stats.data_ptr(), Cn, 1,
                                        float(self.smoothing), torch.cuda.current_stream().cuda_stream), "lf_ema_update")
        self.counter += 1

    @property
    def offset(self):
        """mean_m(x) - x  (utils/EMA.py:36-38)."""
        if self._offset is None:
            return torch.mean(self.x, dim=0, keepdim=True) - self.x
        return self._offset
