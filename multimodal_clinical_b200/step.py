"""Host-side driver of the fused late-fusion step: owns device state (EMA, QMF History), scratch and the
packed statistics buffer, sequences the C-ABI calls on the current CUDA stream, and places the
collectives of the batch-sharded (data-parallel) layout.

One process per GPU.  Rank r holds samples [r*B/G, (r+1)*B/G) of the global batch; heads, EMA state and
the QMF History are replicated and stay bit-identical across ranks because every rank applies the same
update from all-reduced / all-gathered inputs.  Exchanges per step:
  JLOGITS / OGM-GE : all-reduce(stats)                          -> all-reduce([dW1|db1|dW2|db2|cal counts])
  QMF              : all-reduce(stats), all-gather(idx, conf)   -> all-reduce([dW1|db1|dW2|db2|cal counts])
No host synchronisation happens inside a step; metrics are read from ``stats`` lazily.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib, parallel
from ._lib import (LF_MODE_JLOGITS, LF_MODE_QMF, LF_PREC_FP32, LF_PREC_TF32, LF_STATS_HEADER, STAT,
                   LfHeadsArgs, LfQmfArgs, LfTensorList, check)

_MOD = {"OGM_GE": _lib.LF_MOD_OGM_GE, "OGM": _lib.LF_MOD_OGM, "noise": _lib.LF_MOD_NOISE}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.LfError(f"{what} must live on a CUDA device: the fused late-fusion step has no CPU path")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


@dataclass
class StepOutput:
    logits: List[torch.Tensor]            # x1_logits, x2_logits (B,C)
    avg_logits: torch.Tensor              # (B,C)
    logits_df: Optional[torch.Tensor]     # (B,C) QMF
    conf: Optional[torch.Tensor]          # (2,B) QMF
    loss: torch.Tensor                    # 0-d fp32, global-batch loss
    dfeat: List[Optional[torch.Tensor]]   # dL/df_m (B,D)
    dweight: List[torch.Tensor]           # dL/dW_m (C,D), already all-reduced
    dbias: List[torch.Tensor]             # (C)
    stats: torch.Tensor                   # packed fp64 statistics (global sums)
    batch_global: int

    def metric(self, name: str) -> torch.Tensor:
        """Accuracy (count / global batch) or raw statistic as a 0-d device tensor; no sync."""
        return self.stats[STAT[name]]

    def accuracies(self) -> dict:
        """One D2H copy for all step metrics (utils/BaseModel.py:78-92, 961)."""
        h = self.stats[:LF_STATS_HEADER].cpu()
        n = float(self.batch_global)
        return {"x1_acc_uncal": h[STAT["CNT_X1"]].item() / n, "x2_acc_uncal": h[STAT["CNT_X2"]].item() / n,
                "x1_acc_cal": h[STAT["CNT_X1_CAL"]].item() / n, "x2_acc_cal": h[STAT["CNT_X2_CAL"]].item() / n,
                "joint_acc": h[STAT["CNT_JOINT"]].item() / n, "df_acc": h[STAT["CNT_DF"]].item() / n,
                "score1": h[STAT["SCORE_X1"]].item(), "score2": h[STAT["SCORE_X2"]].item()}


class LateFusionStep:
    """Fused heads + fusion + loss + backward + EMA (+ QMF History / OGM-GE coefficients) for two modalities.

    Replaces, for one batch: FusionNet.forward's head/fusion/loss lines (cremad/joint_model_qmf.py:57-75,
    cremad/joint_model_ogm_ge.py:50-58), autograd through them, EMA.update/offset (utils/EMA.py) and the
    metric computations of utils/BaseModel.py training_step.
    """

    def __init__(self, num_classes: int, mode: str = "jlogits", n_data: Optional[int] = None,
                 device: Optional[torch.device] = None, precision: str = "fp32", ema_smoothing: float = 0.05,
                 process_group=None, qmf_state=None, ema=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.LfError("LateFusionStep needs a CUDA device (sm_100a); there is no CPU fallback")
        self.C = int(num_classes)
        self.mode = {"jlogits": LF_MODE_JLOGITS, "ogm_ge": LF_MODE_JLOGITS, "qmf": LF_MODE_QMF}[mode]
        self.precision = {"fp32": LF_PREC_FP32, "tf32": LF_PREC_TF32}[precision]
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.pg = process_group
        self.rank, self.world = parallel.world(process_group)
        self.smoothing = float(ema_smoothing)
        dev = self.device
        self.ema_x = torch.zeros(2, self.C, device=dev)          # EMA.x      (utils/EMA.py:25)
        self.ema_offset = torch.zeros(2, self.C, device=dev)     # EMA.offset (utils/EMA.py:36-38)
        self.ema_counter = 0
        self.stats = torch.zeros(LF_STATS_HEADER + 2 * self.C, dtype=torch.float64, device=dev)
        self.coeff = torch.ones(2, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.n_data = None
        if self.mode == LF_MODE_QMF:
            if n_data is None:
                raise ValueError("QMF mode needs n_data (args.num_samples)")
            self.n_data = int(n_data)
            if qmf_state is None:
                from .existing_algos.QMF import _QmfState
                qmf_state = _QmfState(2, self.n_data)
            if qmf_state.n_data != self.n_data:
                raise ValueError("QMF state length differs from n_data")
            self.qmf_state = qmf_state.to(dev)       # History arrays (2,N) fp64 in HBM, shared with the QMF object
        if ema is not None:                           # share the calibration state with a utils.EMA.EMA object
            ema.to(dev)
            self.ema_x, self.ema_offset, self.smoothing = ema.x, ema._offset, float(ema.smoothing)
        self.ema = ema
        self.fresh_outputs = False
        self._ws = None
        self._ws_key = None
        self._bufs = {}
        self.mod_ws = torch.empty(self.lib.lf_modulate_workspace_bytes(), dtype=torch.uint8, device=dev)

    @property
    def correctness(self) -> torch.Tensor:
        """(2,N) fp64 History.correctness (existing_algos/QMF.py:13), device-resident."""
        return self.qmf_state.correctness

    @property
    def confidence(self) -> torch.Tensor:
        return self.qmf_state.confidence

    # ------------------------------------------------------------------ buffers
    def _buffers(self, B: int, D: int, need_dfeat: bool):
        key = (B, D, need_dfeat)
        if self._ws_key != key:
            dev, Cn = self.device, self.C
            nbytes = self.lib.lf_workspace_bytes(B, D, Cn)
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            qmf = self.mode == LF_MODE_QMF
            b = {}
            b["logits"] = torch.empty(2, B, Cn, device=dev)
            b["avg"] = torch.empty(B, Cn, device=dev)
            b["zdf"] = torch.empty(B, Cn, device=dev) if qmf else None
            b["conf"] = torch.empty(2, B, device=dev) if qmf else None
            b["ldz"] = (Cn + 3) // 4 * 4            # dL/dlogits rows padded to 16 B (TMA row pitch, 128-bit access)
            b["dz"] = torch.zeros(2 if qmf else 1, B, b["ldz"], device=dev)
            b["dfeat"] = torch.empty(2, B, D, device=dev) if need_dfeat else None
            # one flat buffer [dW1 | db1 | dW2 | db2 | cal1 cal2] so the gradient exchange is ONE all-reduce
            n = Cn * D
            b["grad_flat"] = torch.empty(2 * (n + Cn) + 2, device=dev)
            b["qmf_g"] = torch.empty(2, B, device=dev) if qmf else None
            self._bufs = b
            self._ws_key = key
        if self.fresh_outputs:
            # tensors that escape to the caller (autograd, metric lists) get fresh storage every step from
            # torch's caching allocator; scratch (dz, qmf_g, workspace) stays static
            dev, Cn, b = self.device, self.C, dict(self._bufs)
            b["logits"] = torch.empty(2, B, Cn, device=dev)
            b["avg"] = torch.empty(B, Cn, device=dev)
            if b["zdf"] is not None:
                b["zdf"] = torch.empty(B, Cn, device=dev)
                b["conf"] = torch.empty(2, B, device=dev)
            if need_dfeat:
                b["dfeat"] = torch.empty(2, B, D, device=dev)
            b["grad_flat"] = torch.empty(2 * (Cn * D + Cn) + 2, device=dev)
            return b
        return self._bufs

    # ------------------------------------------------------------------ the step
    def step(self, feats: Sequence[torch.Tensor], weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
             label: torch.Tensor, idx: Optional[torch.Tensor] = None, need_dfeat: bool = True,
             update_ema: bool = True, ogm_alpha: Optional[float] = None, backward: bool = True) -> StepOutput:
        """One fused step.  ``backward=False`` is the validation/test forward of the reference (loss, logits
        and -- QMF -- the History update with the batch's indices, utils/BaseModel.py:1023-1026), without
        gradients; ``update_ema=False`` leaves the calibration state alone like the reference's eval steps."""
        lib = self.lib
        f = [_f32c(feats[0], "features"), _f32c(feats[1], "features")]
        W = [_f32c(weights[0], "weight"), _f32c(weights[1], "weight")]
        bb = [_f32c(biases[0], "bias"), _f32c(biases[1], "bias")]
        B, D = f[0].shape
        Cn = self.C
        if f[1].shape != (B, D) or W[0].shape != (Cn, D) or W[1].shape != (Cn, D):
            raise ValueError(f"shape mismatch: feats {tuple(f[0].shape)}/{tuple(f[1].shape)}, weights {tuple(W[0].shape)}")
        label = label.to(device=self.device, dtype=torch.int64).contiguous()
        Bg = B * self.world
        bufs = self._buffers(B, D, need_dfeat)
        if self.fresh_outputs:
            self.stats = torch.zeros_like(self.stats)
            self.loss = torch.empty_like(self.loss)
        n = Cn * D
        gf = bufs["grad_flat"]
        dW = [gf[0:n].view(Cn, D), gf[n + Cn:2 * n + Cn].view(Cn, D)]
        db = [gf[n:n + Cn], gf[2 * n + Cn:2 * n + 2 * Cn]]
        qmf = self.mode == LF_MODE_QMF

        a = LfHeadsArgs()
        a.batch, a.batch_global, a.dim, a.classes = B, Bg, D, Cn
        a.mode, a.precision, a.need_dfeat = self.mode, self.precision, int(need_dfeat)
        a.ld_dlogits = bufs["ldz"]
        for m in range(2):
            a.feat[m] = _ptr(f[m]); a.weight[m] = _ptr(W[m]); a.bias[m] = _ptr(bb[m])
            a.logits[m] = _ptr(bufs["logits"][m])
            a.dfeat[m] = _ptr(bufs["dfeat"][m]) if need_dfeat else None
            a.dweight[m] = _ptr(dW[m]); a.dbias[m] = _ptr(db[m])
        a.label = _ptr(label)
        a.avg_logits = _ptr(bufs["avg"])
        a.logits_df = _ptr(bufs["zdf"]); a.conf = _ptr(bufs["conf"])
        a.dlogits[0] = _ptr(bufs["dz"][0]); a.dlogits[1] = _ptr(bufs["dz"][1]) if qmf else None
        a.qmf_g = _ptr(bufs["qmf_g"]); a.ema_offset = _ptr(self.ema_offset)
        a.stats = _ptr(self.stats)
        a.workspace = _ptr(self._ws); a.workspace_bytes = self._ws.numel()
        st = _stream()

        check(lib.lf_heads_forward(C.byref(a), st), "lf_heads_forward")
        parallel.allreduce_sum_(self.stats, self.pg)              # score sums, CE sums, logit sums, counts
        if update_ema:
            check(lib.lf_ema_update(_ptr(self.ema_x), _ptr(self.ema_offset), _ptr(self.stats), Cn, Bg,
                                    self.smoothing, st), "lf_ema_update")
            self.ema_counter += 1
            if self.ema is not None:
                self.ema.counter += 1
        if ogm_alpha is not None:
            check(lib.lf_ogm_coeff(_ptr(self.stats), float(ogm_alpha), _ptr(self.coeff), st), "lf_ogm_coeff")
        if qmf:
            if idx is None:
                raise ValueError("QMF step needs the dataset indices of the batch (idx)")
            idx = idx.to(device=self.device, dtype=torch.int64).contiguous().view(-1)
            idx_g, conf_g = parallel.gather_batch(idx, bufs["conf"], self.pg)
            q = LfQmfArgs()
            q.batch_global, q.n_data = Bg, self.n_data
            q.idx, q.conf = _ptr(idx_g), _ptr(conf_g)
            qs = self.qmf_state
            q.correctness, q.confidence = _ptr(qs.correctness), _ptr(qs.confidence)
            q.last_writer, q.step_base = _ptr(qs.last_writer), qs.step_base
            q.stats, q.qmf_g, q.target_out = _ptr(self.stats), _ptr(bufs["qmf_g"]), None
            q.g_begin, q.g_count = parallel.shard_range(self.rank, B)
            q.workspace, q.workspace_bytes = _ptr(qs.ws), qs.ws.numel()
            q.flags = _lib.LF_QMF_ALL
            check(lib.lf_qmf_history_step(C.byref(q), st), "lf_qmf_history_step")
            qs.step_base += Bg
        if backward:
            check(lib.lf_heads_backward(C.byref(a), st), "lf_heads_backward")
            # head gradients + calibrated counts: one all-reduce
            parallel.pack_grad_exchange(gf, 2 * (n + Cn), self.stats, STAT["CNT_X1_CAL"], STAT["CNT_X2_CAL"] + 1, self.pg)
        check(lib.lf_loss_finalize(_ptr(self.stats), self.mode, Bg, _ptr(self.loss), st), "lf_loss_finalize")

        return StepOutput(
            logits=[bufs["logits"][0], bufs["logits"][1]], avg_logits=bufs["avg"], logits_df=bufs["zdf"],
            conf=bufs["conf"], loss=self.loss[0],
            dfeat=[bufs["dfeat"][0], bufs["dfeat"][1]] if (need_dfeat and backward) else [None, None],
            dweight=dW, dbias=db, stats=self.stats, batch_global=Bg)

    # ------------------------------------------------------------------ OGM-GE modulation
    def modulate(self, grads: Sequence[torch.Tensor], which: int, modulation: str, seed: int, offset: int) -> None:
        """In-place OGM-GE add_factor over the 4-D gradients of encoder ``which`` (existing_algos/OGM_GE.py:42-54)."""
        sel = [g for g in grads if g is not None and g.dim() == 4]
        for i in range(0, len(sel), _lib.LF_MAX_TENSORS):
            chunk = sel[i:i + _lib.LF_MAX_TENSORS]
            tl = LfTensorList()
            tl.count = len(chunk)
            for k, g in enumerate(chunk):
                if not (g.is_cuda and g.dtype == torch.float32 and g.is_contiguous()):
                    raise _lib.LfError("OGM-GE modulation needs contiguous fp32 CUDA gradients")
                tl.data[k] = g.data_ptr(); tl.numel[k] = g.numel()
            check(self.lib.lf_ogm_modulate(C.byref(tl), _ptr(self.coeff[which:which + 1]), _MOD[modulation], int(seed),
                                           int(offset) + i, _ptr(self.mod_ws), self.mod_ws.numel(), _stream()),
                  "lf_ogm_modulate")
