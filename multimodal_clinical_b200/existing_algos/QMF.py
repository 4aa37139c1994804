"""QMF energy-confidence fusion and its ranking-loss History, device-resident.

Mirrors existing_algos/QMF.py of the reference (class / method names and argument meaning):
``History(n_data)`` with ``.correctness`` / ``.confidence`` / ``correctness_update`` /
``get_target_margin``, and ``QMF(n_modality, n_data)`` with ``.history``, ``.df`` and ``.reg_loss``.
The arrays live in HBM as fp64 (the reference keeps them as host numpy fp64 and pays ~14 host round
trips per step); ``.correctness`` / ``.confidence`` give numpy copies on demand.

The training path does not call these methods one by one: ``FusedLateFusionHead`` (heads.py) runs the
update, the min/max normalisation, the ranking loss and its gradient inside the fused step, sharing this
object's state.  The methods below are the stand-alone API for code that uses the algorithm directly;
they are forward-only (no autograd graph) and run on CUDA through the C ABI.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from .._lib import LfQmfArgs, check


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _QmfState:
    """(2, N) fp64 History arrays + scratch shared by the two History views and the fused step."""

    def __init__(self, n_modality: int, n_data: int):
        if n_modality != 2:
            raise NotImplementedError("the fused late-fusion step supports two modalities")
        self.n_data = int(n_data)
        self.device: Optional[torch.device] = None
        self.correctness = self.confidence = self.last_writer = self.ws = self.stats = None

    def to(self, device) -> "_QmfState":
        device = torch.device(device)
        if self.device == device:
            return self
        if device.type != "cuda":
            raise _lib.LfError("QMF History lives on a CUDA device; there is no CPU path")
        lib = _lib.load()
        if self.correctness is not None:                    # state follows the model to another GPU
            self.correctness = self.correctness.to(device); self.confidence = self.confidence.to(device)
        else:
            self.correctness = torch.zeros(2, self.n_data, dtype=torch.float64, device=device)   # QMF.py:13
            self.confidence = torch.zeros(2, self.n_data, dtype=torch.float64, device=device)    # QMF.py:14
        self.last_writer = torch.zeros(self.n_data + 1, dtype=torch.int64, device=device)   # [N] = ticket counter
        self.ws = torch.empty(lib.lf_qmf_workspace_bytes(self.n_data), dtype=torch.uint8, device=device)
        self.mid_ws = None
        self.stats = torch.zeros(_lib.LF_STATS_HEADER, dtype=torch.float64, device=device)
        self.device = device
        return self

    def mid_workspace(self, batch_global: int) -> torch.Tensor:
        need = _lib.load().lf_mid_workspace_bytes(int(batch_global))
        if self.mid_ws is None or self.mid_ws.numel() < need:
            self.mid_ws = torch.zeros(need, dtype=torch.uint8, device=self.device)    # holds a counter the kernel leaves at zero
        return self.mid_ws

    def run(self, idx: torch.Tensor, conf: torch.Tensor, flags: int, loss_uni=(None, None),
            qmf_g: Optional[torch.Tensor] = None, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.to(idx.device)
        Bg = idx.numel()
        q = LfQmfArgs()
        q.batch_global, q.n_data = Bg, self.n_data
        q.idx, q.conf = idx.data_ptr(), conf.data_ptr()
        q.correctness, q.confidence = self.correctness.data_ptr(), self.confidence.data_ptr()
        q.last_writer, q.step_base = self.last_writer.data_ptr(), 0      # device-resident ticket counter
        stats = self.stats if stats is None else stats
        q.stats = stats.data_ptr()
        q.qmf_g = qmf_g.data_ptr() if qmf_g is not None else None
        q.target_out = None
        q.g_begin, q.g_count = 0, Bg
        q.workspace, q.workspace_bytes = self.ws.data_ptr(), self.ws.numel()
        q.flags = flags
        for m in range(2):
            q.loss_uni[m] = loss_uni[m].data_ptr() if loss_uni[m] is not None else None
        check(_lib.load().lf_qmf_history_step(C.byref(q), _stream()), "lf_qmf_history_step")
        return stats


class History(object):
    """One modality's view of the History (existing_algos/QMF.py:5-68)."""

    def __init__(self, n_data, _state: Optional[_QmfState] = None, _m: int = 0):
        self._state = _state if _state is not None else _QmfState(2, n_data)
        self._m = _m
        self.max_correctness = 1
        self.use_ema = True
        self.alpha = 0.1

    def _arr(self, which: str) -> np.ndarray:
        t = getattr(self._state, which)
        if t is None:
            return np.zeros(self._state.n_data)
        return t[self._m].cpu().numpy()

    @property
    def correctness(self) -> np.ndarray:
        return self._arr("correctness")

    @property
    def confidence(self) -> np.ndarray:
        return self._arr("confidence")

    def correctness_update(self, data_idx, correctness, confidence):
        """corr[idx] <- 0.9 corr[idx] + 0.1 * loss (scalar batch-mean CE), conf[idx] <- confidence
        (existing_algos/QMF.py:20-29; duplicates: last writer wins, like numpy fancy assignment)."""
        idx = data_idx.reshape(-1).to(torch.int64).contiguous()
        if not idx.is_cuda:
            raise _lib.LfError("History.correctness_update needs CUDA tensors")
        loss = correctness.detach().reshape(-1).float().contiguous()
        if loss.numel() != 1:
            raise NotImplementedError("the reference hands History a 0-d batch-mean loss (cremad/joint_model_qmf.py:64-65)")
        conf2 = torch.zeros(2, idx.numel(), device=idx.device)
        conf2[self._m] = confidence.detach().reshape(-1).float()
        lu = [None, None]; lu[self._m] = loss
        self._state.run(idx, conf2, _lib.LF_QMF_UPDATE_X1 << self._m, loss_uni=lu)

    def max_correctness_update(self, epoch):
        if epoch > 1:
            self.max_correctness += 1


class QMF:
    """existing_algos/QMF.py:70-141."""

    def __init__(self, n_modality, n_data):
        self._state = _QmfState(n_modality, n_data)
        self.history: List[History] = [History(n_data, self._state, m) for m in range(n_modality)]
        self.n_modality = n_modality

    def df(self, logits: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """logits (M,B,C) -> (logits_df (B,C), conf (M,B)); energy = log(sum(exp z)) un-stabilised like the
        reference (existing_algos/QMF.py:113-117)."""
        if logits.dim() != 3 or logits.shape[0] != 2:
            raise NotImplementedError("QMF.df expects (2, B, C) logits")
        if not logits.is_cuda:
            raise _lib.LfError("QMF.df needs CUDA tensors; there is no CPU path")
        z = logits.detach().float().contiguous()
        _, B, Cn = z.shape
        zdf = torch.empty(B, Cn, device=z.device)
        conf = torch.empty(2, B, device=z.device)
        check(_lib.load().lf_qmf_df(z[0].data_ptr(), z[1].data_ptr(), B, Cn, zdf.data_ptr(), conf.data_ptr(), _stream()),
              "lf_qmf_df")
        return zdf, conf

    def reg_loss(self, confidence: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """Ranking regulariser (existing_algos/QMF.py:119-141), including the flattened roll and the
        ``rank_margin[n]`` indexing of the reference.  Needs a batch of at least two, like the reference."""
        idx = idx.reshape(-1).to(torch.int64).contiguous()
        if idx.numel() < 2:
            raise TypeError("len() of unsized object")       # what the reference raises for a batch of one
        conf = confidence.detach().float().contiguous()
        g = torch.empty(2, idx.numel(), device=idx.device)
        stats = self._state.run(idx, conf, _lib.LF_QMF_REG, qmf_g=g)
        self.last_dconf = g                                  # dL_reg/dconf, for callers that chain gradients by hand
        return (stats[_lib.STAT["REG_SUM"]] / idx.numel()).float()
