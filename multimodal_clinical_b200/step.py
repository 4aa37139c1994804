"""Host-side driver of the fused late-fusion step: owns device state (EMA, QMF History), scratch and the
packed statistics buffer, sequences the C-ABI calls on the current CUDA stream, and places the
collectives of the batch-sharded (data-parallel) layout.

One process per GPU.  Rank r holds samples [r*B/G, (r+1)*B/G) of the global batch; heads, EMA state and
the QMF History are replicated and stay bit-identical across ranks because every rank applies the same
update from all-reduced / all-gathered inputs.  Exchanges per step:
  forward -> all-gather([partial stats | idx | conf])  -> lf_step_mid (sums the partials in rank order)
          -> backward -> all-reduce([dW1|db1|dW2|db2|calibrated counts])
i.e. two collectives per step for every head type (idx / conf are only present for QMF).
No host synchronisation happens inside a step; metrics are read from ``stats`` lazily.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib, parallel
from ._lib import (LF_MODE_JLOGITS, LF_MODE_QMF, LF_PREC_BF16, LF_PREC_FP32, LF_PREC_TF32, LF_STATS_HEADER, STAT,
                   LfHeadsArgs, LfMidArgs, LfQmfArgs, LfSgdFused, LfTensorList, check)

_MOD = {"OGM_GE": _lib.LF_MOD_OGM_GE, "OGM": _lib.LF_MOD_OGM, "noise": _lib.LF_MOD_NOISE}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _devc(t: torch.Tensor, dtype, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.LfError(f"{what} must live on a CUDA device: the fused late-fusion step has no CPU path")
    return (t if t.dtype == dtype else t.to(dtype)).contiguous()


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
                 process_group=None, qmf_state=None, ema=None, comm: str = "auto", loss_terms: int = 0,
                 sharded: Optional[bool] = None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.LfError("LateFusionStep needs a CUDA device (sm_100a); there is no CPU fallback")
        self.C = int(num_classes)
        # "ensemble": one CE per modality (cremad/ensemble_model_noised.py:52-53) = the QMF kernels' unimodal terms
        # without the joint term, the ranking regulariser and the History
        self.mode = {"jlogits": LF_MODE_JLOGITS, "ogm_ge": LF_MODE_JLOGITS, "qmf": LF_MODE_QMF, "ensemble": LF_MODE_QMF}[mode]
        if mode == "ensemble":
            loss_terms = int(loss_terms) | _lib.LF_LOSS_NO_JOINT | _lib.LF_LOSS_NO_REG
        self.precision = {"fp32": LF_PREC_FP32, "tf32": LF_PREC_TF32, "bf16": LF_PREC_BF16}[precision]
        self.bf16 = self.precision == LF_PREC_BF16
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.pg = process_group
        # QMF loss-term ablations (LF_LOSS_NO_JOINT = 1, LF_LOSS_NO_UNI = 2; include/lf_fusion.h)
        self.loss_terms = int(loss_terms)
        # sharded=None: join the (default or given) process group when torch.distributed is initialised -- the
        # explicit engine API used by bench.py; False: treat the batch as the whole batch whatever torch.distributed
        # says (what FusedLateFusionHead passes under a DDP wrapper, where DDP reduces the gradients itself)
        self.rank, self.world = parallel.world(process_group) if (sharded is None or sharded) else (0, 1)
        self.smoothing = float(ema_smoothing)
        dev = self.device
        self.ema_x = torch.zeros(2, self.C, device=dev)          # EMA.x      (utils/EMA.py:25)
        self.ema_offset = torch.zeros(2, self.C, device=dev)     # EMA.offset (utils/EMA.py:36-38)
        self.ema_counter = 0
        self.stats = torch.zeros(LF_STATS_HEADER + 2 * self.C, dtype=torch.float64, device=dev)
        self.coeff = torch.ones(2, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.n_data = None
        self.has_history = self.mode == LF_MODE_QMF and not (self.loss_terms & _lib.LF_LOSS_NO_REG)
        if self.has_history:
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
        # "peer": exchanges through NVLink peer memory (csrc/lf_peer.cu); "nccl": torch.distributed collectives;
        # "auto": peer when several CUDA ranks share one node and the IPC mapping succeeds
        self.comm_mode = comm
        self.peer = None
        self._peer_key = None
        self._peers = {}                 # (payload bytes, gradient floats) -> PeerComm or None, built once per size
        self.fresh_outputs = False
        self._pay = None
        self._pay_key = None
        self._ws = None
        self._ws_key = None
        self._bufs = {}
        self.mod_ws = torch.empty(self.lib.lf_modulate_workspace_bytes(), dtype=torch.uint8, device=dev)
        # bf16 copies of the heads (LF_PREC_BF16), re-cast only when a head tensor changes (torch bumps ``_version`` on
        # every in-place update) or kept current by the fused SGD step; ``cache_w16 = False`` casts on every call
        self.cache_w16 = True
        self._w16 = None
        self._w16_key = None
        self._sgd = None
        self._capturing = False
        # mean fusion on wide heads: lf_step_mid + the calibrated-count pass beside the dfeat GEMM (see step());
        # LF_NO_CAL_OVERLAP=1 keeps the single-stream order for A/B runs
        self.cal_overlap = not os.environ.get("LF_NO_CAL_OVERLAP")
        self.cal_overlap_min_classes = 129
        self._side = None

    # ------------------------------------------------------------------ fused SGD on the heads
    def enable_sgd(self, lr: float, momentum: float = 0.9, weight_decay: float = 1.0e-4) -> None:
        """Apply torch.optim.SGD(lr, momentum, weight_decay) (utils/BaseModel.py:275-285) to the head tensors passed
        to ``step`` INSIDE the step: the update runs in the tail of the dW kernel, right after the gradients are
        reduced (after the gradient all-reduce on several GPUs), and also refreshes the bf16 copies of the heads.
        The hyper-parameters live on the device, so a captured graph follows ``set_lr`` (StepLR)."""
        if self.C < 32 or (self.precision == LF_PREC_FP32 and os.environ.get("LF_NO_X3")):
            raise _lib.LfError("the fused SGD step is part of the tensor-pipe backward (wide heads, C >= 32)")
        self._sgd = {"hyper": torch.tensor([lr, momentum, weight_decay, 0.0], device=self.device), "mom": None, "key": None}

    def set_lr(self, lr: float) -> None:
        self._sgd["hyper"][0:1].fill_(float(lr))

    def disable_sgd(self) -> None:
        self._sgd = None

    def _heads_bf16(self, W):
        """(ptr0, ptr1) of current bf16 copies of the heads, or None to let lf_heads_forward cast per call."""
        if not self.bf16 or not self.cache_w16 or (self._capturing and self._sgd is None):
            return None                       # inside a graph only the fused SGD keeps the copies current
        key = (W[0].data_ptr(), W[0]._version, W[1].data_ptr(), W[1]._version)
        if self._w16 is None or self._w16.shape[1:] != W[0].shape:
            self._w16 = torch.empty((2,) + tuple(W[0].shape), dtype=torch.bfloat16, device=self.device)
            self._w16_key = None
        if self._w16_key != key:
            check(self.lib.lf_cast_heads_bf16(W[0].data_ptr(), W[1].data_ptr(), self._w16.data_ptr(), W[0].numel(), _stream()),
                  "lf_cast_heads_bf16")
            self._w16_key = key
        return self._w16[0].data_ptr(), self._w16[1].data_ptr()

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
            self._ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)     # starts with zeroed inter-CTA counters
            qmf = self.mode == LF_MODE_QMF
            b = {}
            # tensor-pipe path: logits rows padded to 16 B so the GEMM epilogue can TMA-store them; callers get views
            # (wide heads in every precision: LF_PREC_FP32 runs the same kernels through the 3xTF32 operand split)
            no_x3 = self.precision == LF_PREC_FP32 and os.environ.get("LF_NO_X3")     # A/B switch: FMA GEMMs, dense rows
            b["ldl"] = (Cn + 3) // 4 * 4 if (Cn >= 32 and not no_x3) else Cn
            b["logits_store"] = torch.empty(2, B, b["ldl"], device=dev)
            b["logits"] = b["logits_store"][:, :, :Cn]
            # avg / z_df share the padded pitch up to 256 classes, where the vector row kernels write them with 128-bit
            # stores; wider heads use the one-warp-per-sample kernels, which measured faster on dense rows
            b["ldf"] = b["ldl"] if Cn <= 256 else Cn
            b["avg_store"] = torch.empty(B, b["ldf"], device=dev)
            b["avg"] = b["avg_store"][:, :Cn]
            b["zdf_store"] = torch.empty(B, b["ldf"], device=dev) if qmf else None
            b["zdf"] = b["zdf_store"][:, :Cn] if qmf else None
            b["conf"] = torch.empty(2, B, device=dev) if qmf else None
            fdt = torch.bfloat16 if self.bf16 else torch.float32
            b["ldz"] = (Cn + 7) // 8 * 8 if self.bf16 else (Cn + 3) // 4 * 4     # dL/dlogits rows padded to 16 B (TMA pitch)
            b["dz"] = torch.zeros(2 if qmf else 1, B, b["ldz"], device=dev, dtype=fdt)
            b["dfeat"] = torch.empty(2, B, D, device=dev, dtype=fdt) if need_dfeat else None
            # one flat buffer [dW1 | db1 | dW2 | db2 | cal1 cal2] so the gradient exchange is ONE all-reduce
            n = Cn * D
            b["grad_flat"] = torch.empty((2 * (n + Cn) + 2 + 3) // 4 * 4, device=dev)[:2 * (n + Cn) + 2]   # 16-B padded slot
            b["qmf_g"] = torch.zeros(2, B, device=dev) if qmf else None      # stays zero for the ensemble loss (no ranking term)
            self._bufs = b
            self._ws_key = key
        if self.fresh_outputs:
            # tensors that escape to the caller (autograd, metric lists) get fresh storage every step from
            # torch's caching allocator; scratch (dz, qmf_g, workspace) stays static
            dev, Cn, b = self.device, self.C, dict(self._bufs)
            b["logits_store"] = torch.empty(2, B, b["ldl"], device=dev)
            b["logits"] = b["logits_store"][:, :, :Cn]
            b["avg_store"] = torch.empty(B, b["ldf"], device=dev)
            b["avg"] = b["avg_store"][:, :Cn]
            if b["zdf"] is not None:
                b["zdf_store"] = torch.empty(B, b["ldf"], device=dev)
                b["zdf"] = b["zdf_store"][:, :Cn]
                b["conf"] = torch.empty(2, B, device=dev)
            if need_dfeat:
                b["dfeat"] = torch.empty(2, B, D, device=dev, dtype=torch.bfloat16 if self.bf16 else torch.float32)
            b["grad_flat"] = torch.empty((2 * (Cn * D + Cn) + 2 + 3) // 4 * 4, device=dev)[:2 * (Cn * D + Cn) + 2]
            return b
        return self._bufs

    def _peer_comm(self, payload_bytes: int, grad_floats: int):
        """The peer-memory communicator for these sizes (built once; collective over the process group)."""
        if self.world == 1 or self.comm_mode == "nccl" or self.device.type != "cuda":
            return None
        key = (payload_bytes, grad_floats)
        if key not in self._peers:
            # one communicator per exchange size, kept for the life of the engine: the short last batch of an epoch
            # and the full batches that follow alternate between two cached communicators instead of re-mapping
            try:
                peer = parallel.PeerComm(payload_bytes, grad_floats, self.pg)
            except _lib.LfError:
                if self.comm_mode == "peer":
                    raise
                peer = None                             # e.g. IPC not permitted: stay on the NCCL collectives
            ok = torch.tensor([1 if peer is not None else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.pg)
            if int(ok.item()) == 0:
                if peer is not None:
                    peer.close()
                peer = None
            self._peers[key] = peer
        self.peer = self._peers[key]
        self._peer_key = key
        return self.peer

    def close(self) -> None:
        """Unmap and free the peer-memory communicators (collective-free; call on every rank)."""
        for peer in self._peers.values():
            if peer is not None:
                peer.close()
        self._peers = {}
        self.peer = None

    def check_peer(self) -> None:
        """Raise LfError if a peer exchange of an earlier step gave up waiting for another rank (one host sync)."""
        for peer in self._peers.values():
            if peer is not None:
                peer.check()

    def _side_stream(self):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def _payload(self, B: int, qmf: bool):
        n_stats = LF_STATS_HEADER + 2 * self.C
        self._off_idx = 8 * n_stats
        self._off_conf = self._off_idx + (8 * B if qmf else 0)
        nbytes = (self._off_conf + (8 * B if qmf else 0) + 15) // 16 * 16
        key = (B, qmf)
        if self.fresh_outputs or self._pay_key != key:
            self._pay = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._pay_key = key
        pay = self._pay
        p_stats = pay[:self._off_idx].view(torch.float64)
        p_idx = pay[self._off_idx:self._off_conf].view(torch.int64) if qmf else None
        p_conf = pay[self._off_conf:].view(torch.float32).view(2, B) if qmf else None
        return pay, p_stats, p_idx, p_conf

    # ------------------------------------------------------------------ the step
    def step(self, feats: Sequence[torch.Tensor], weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
             label: torch.Tensor, idx: Optional[torch.Tensor] = None, need_dfeat: bool = True,
             update_ema: bool = True, ogm_alpha: Optional[float] = None, backward: bool = True) -> StepOutput:
        """One fused step.  ``backward=False`` is the validation/test forward of the reference (loss, logits
        and -- QMF -- the History update with the batch's indices, utils/BaseModel.py:1023-1026), without
        gradients; ``update_ema=False`` leaves the calibration state alone like the reference's eval steps."""
        lib = self.lib
        if self.bf16:            # features arrive in bf16 (what the encoders emit under autocast); fp32 inputs are rounded once
            f = [_devc(feats[0], torch.bfloat16, "features"), _devc(feats[1], torch.bfloat16, "features")]
        else:
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
        # this rank's contribution to the one exchange before the backward pass:
        # [partial statistics (16+2C) f64 | idx (B) i64 | conf (2,B) f32], 8-byte aligned pieces
        pay, p_stats, p_idx, p_conf = self._payload(B, qmf)
        hist = self.has_history
        idx_in_place = True
        if hist:
            if idx is None:
                raise ValueError("QMF step needs the dataset indices of the batch (idx)")
            idx = idx.reshape(-1)
            # step_mid reads the caller's indices in place (one GPU) or pushes them to the peers from where they are
            # (peer exchange): no staging copy kernel in the step.  Only the NCCL all-gather needs them in the payload.
            idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
            self._idx_keep = idx
            idx_in_place = self.world == 1 or self.comm_mode != "nccl"
            if not idx_in_place:
                p_idx.copy_(idx)

        a = LfHeadsArgs()
        a.batch, a.batch_global, a.dim, a.classes = B, Bg, D, Cn
        a.mode, a.precision, a.need_dfeat = self.mode, self.precision, int(need_dfeat)
        a.ld_dlogits = bufs["ldz"]
        a.ld_logits = bufs["ldl"]
        a.ld_fused = bufs["ldf"]
        a.loss_terms = self.loss_terms
        a.fwd_only = int(not backward)
        for m in range(2):
            a.feat[m] = _ptr(f[m]); a.weight[m] = _ptr(W[m]); a.bias[m] = _ptr(bb[m])
            a.logits[m] = _ptr(bufs["logits_store"][m])
            a.dfeat[m] = _ptr(bufs["dfeat"][m]) if need_dfeat else None
            a.dweight[m] = _ptr(dW[m]); a.dbias[m] = _ptr(db[m])
        a.label = _ptr(label)
        a.avg_logits = _ptr(bufs["avg_store"])
        a.logits_df = _ptr(bufs["zdf_store"]); a.conf = _ptr(p_conf)
        a.dlogits[0] = _ptr(bufs["dz"][0]); a.dlogits[1] = _ptr(bufs["dz"][1]) if qmf else None
        a.qmf_g = _ptr(bufs["qmf_g"]); a.ema_offset = _ptr(self.ema_offset)
        a.stats = _ptr(p_stats)                                   # forward writes the LOCAL partial sums
        a.workspace = _ptr(self._ws); a.workspace_bytes = self._ws.numel()
        st = _stream()
        w16 = self._heads_bf16(W)
        if w16 is not None:
            a.weight_bf16[0], a.weight_bf16[1] = w16
        # sharded steps: the peer-memory communicator (None: NCCL collectives) and whether the gradient all-reduce can run
        # inside the dW kernel (tensor-pipe heads), in which case the fused SGD step may follow it there
        peer = self._peer_comm(pay.numel(), max(gf.numel(), int(lib.lf_grad_exchange_floats(D, Cn)))) if self.world > 1 else None
        fuse_ar = peer is not None and backward and bool(lib.lf_heads_backward_fuses_allreduce(C.byref(a)))
        sgd = None
        if self._sgd is not None and backward and self.world > 1 and not fuse_ar:
            raise _lib.LfError("the fused SGD step of a sharded run follows the gradient all-reduce inside the dW kernel: it needs the "
                               "peer-memory communicator (comm='auto' / 'peer') and tensor-pipe heads")
        if self._sgd is not None and backward:
            if W[0] is not weights[0] or W[1] is not weights[1] or bb[0] is not biases[0] or bb[1] is not biases[1]:
                raise _lib.LfError("the fused SGD step updates the head tensors in place: pass contiguous fp32 CUDA tensors")
            key = (W[0].data_ptr(), bb[0].data_ptr(), W[1].data_ptr(), bb[1].data_ptr())
            if self._sgd["key"] != key:          # momentum buffers follow the parameter tensors (zero = torch's first step)
                self._sgd["mom"] = [torch.zeros_like(t) for t in (W[0], bb[0], W[1], bb[1])]
                self._sgd["key"] = key
            sgd = LfSgdFused()
            sgd.hyper = _ptr(self._sgd["hyper"])
            for k, t in enumerate(self._sgd["mom"]):
                sgd.momentum_buf[k] = _ptr(t)
            if w16 is not None:
                sgd.weight_bf16_out[0], sgd.weight_bf16_out[1] = w16
            a.sgd = C.pointer(sgd)

        rows_out = (C.c_uint64 * 2)(0, 0)
        if self.world == 1 or (peer is not None and hist):
            # lf_step_mid sums the forward's per-CTA rows itself (sharded: and exchanges the column sums over peer memory)
            a.stats_rows_out = rows_out
        check(lib.lf_heads_forward(C.byref(a), st), "lf_heads_forward")
        if peer is not None:
            peer.check()                                          # pinned host flag: no synchronisation
        # Mean fusion (cremad/joint_model_ogm_ge.py:54-56): dL/dz is final after the forward pass, so the dfeat GEMM does
        # not depend on lf_step_mid; only the calibrated counts need this step's EMA offsets.  lf_step_mid and the
        # calibrated-count pass (a second read of the logits: 324 MB at the VGGSound shape) then run on a second stream
        # BESIDE the dfeat GEMM, whose CTAs leave 39 K registers per SM and part of the HBM bandwidth unused, and join before
        # the dW GEMM, whose tail consumes the counts (K5: 0.579 -> 0.557 ms; the pair is HBM-bound).
        # (up to 128 classes the dfeat GEMM is epilogue-bound and wants its second group of epilogue warps, which leaves
        # no registers for a co-resident CTA, and the count pass is ~10 us: measured 131 us serial vs 136 us at C = 101)
        # One GPU only: on a sharded step lf_step_mid polls the peers' statistics, and spinning beside the GEMM cost more than
        # the overlap gave (two GPUs, K5: 0.663 ms against 0.614 ms single-stream).
        overlap = (backward and self.cal_overlap and Cn >= self.cal_overlap_min_classes
                   and self.mode == LF_MODE_JLOGITS and self.world == 1
                   and bool(lib.lf_heads_backward_splits_rows(C.byref(a))))
        main_stream = side = None
        if overlap:
            main_stream = torch.cuda.current_stream()
            side = self._side_stream()
            side.wait_stream(main_stream)
            mid_st = side.cuda_stream
        else:
            mid_st = st
        stride = pay.numel()
        mid = LfMidArgs()
        if peer is not None:
            # the exchange is fused into lf_step_mid: push over NVLink, flag barrier, consume the local receive area
            mid.use_peer, mid.payload_local, mid.payload_bytes = 1, _ptr(pay), stride
            mid.off_idx, mid.off_conf = self._off_idx, self._off_conf
            peer.fill(mid.comm)
            gathered = pay
            if hist and idx_in_place:
                mid.payload_idx_src = idx.data_ptr()
        else:
            if hist and idx_in_place and self.world > 1:
                p_idx.copy_(idx)                                  # fell back to NCCL after all (no peer mapping)
            # (world, payload bytes).  An engine that treats its batch as the whole batch (one GPU, or sharded=False under a
            # DDP wrapper) consumes its own payload: it must not look at the process group, whose rank 0 holds other data
            gathered = parallel.gather_payload(pay, self.pg, engine_world=self.world)
        mid.mode, mid.classes, mid.batch_global, mid.n_ranks = self.mode, Cn, Bg, self.world
        mid.batch_local, mid.rank, mid.n_data, mid.update_ema = B, self.rank, self.n_data or 0, int(update_ema)
        base = gathered.data_ptr()
        mid.stats_parts, mid.stats_stride = base, stride // 8
        if rows_out[0]:
            mid.stats_rows, mid.n_stats_rows = rows_out[0], rows_out[1]
        if hist:
            qs = self.qmf_state
            mid.idx_parts, mid.idx_stride = (idx.data_ptr() if self.world == 1 else base + self._off_idx), stride // 8
            mid.conf_parts, mid.conf_stride = base + self._off_conf, stride // 4
            mid.correctness, mid.confidence = _ptr(qs.correctness), _ptr(qs.confidence)
            mid.last_writer, mid.step_base = _ptr(qs.last_writer), 0      # device-resident ticket counter
            mid.qmf_g = _ptr(bufs["qmf_g"]) if backward else None
            mws = qs.mid_workspace(Bg)
            mid.workspace, mid.workspace_bytes = _ptr(mws), mws.numel()
        mid.stats = _ptr(self.stats)
        mid.ema_x, mid.ema_offset, mid.smoothing = _ptr(self.ema_x), _ptr(self.ema_offset), self.smoothing
        if ogm_alpha is not None:
            mid.alpha, mid.coeff_out = float(ogm_alpha), _ptr(self.coeff)
        mid.loss_out = _ptr(self.loss)
        mid.loss_terms = self.loss_terms
        if fuse_ar and hist:
            # ranking terms of this rank's slice only; the partial sums ride in the gradient all-reduce
            if getattr(self, "_reg_partial", None) is None:
                self._reg_partial = torch.zeros(4, device=self.device)
            mid.reg_partial_out = _ptr(self._reg_partial)
        check(lib.lf_step_mid(C.byref(mid), mid_st), "lf_step_mid")
        if update_ema:
            self.ema_counter += 1
            if self.ema is not None:
                self.ema.counter += 1
        if backward:
            a.stats = _ptr(self.stats)                            # calibrated counts join the GLOBAL statistics
            if fuse_ar:
                # ONE launch sequence: the all-reduce of [dW|db|calibrated counts|ranking-loss partial] (and the SGD step)
                # run in the tail of the dW kernel over peer memory
                gcomm = _lib.LfPeerComm()
                peer.fill(gcomm)
                a.grad_comm = C.pointer(gcomm)
                if hist:
                    a.reg_partial, a.loss_out = _ptr(self._reg_partial), _ptr(self.loss)
            if overlap:
                a.bwd_phase = 3                                   # calibrated counts, behind lf_step_mid on the second stream
                check(lib.lf_heads_backward(C.byref(a), mid_st), "lf_heads_backward")
                a.bwd_phase = 2                                   # dfeat GEMM on the main stream, beside them
                check(lib.lf_heads_backward(C.byref(a), st), "lf_heads_backward")
                main_stream.wait_stream(side)
                a.bwd_phase = 4                                   # dW GEMM + tail (reduction, db, counts, [all-reduce], SGD)
                check(lib.lf_heads_backward(C.byref(a), st), "lf_heads_backward")
            elif self.world == 1 or fuse_ar:
                check(lib.lf_heads_backward(C.byref(a), st), "lf_heads_backward")
            else:
                # dW / db first, then their exchange on a side stream while dfeat (which no other rank needs) is
                # still being written on the main stream
                a.bwd_phase = 1
                check(lib.lf_heads_backward(C.byref(a), st), "lf_heads_backward")
                cur = torch.cuda.current_stream()
                side = self._side_stream()
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    sst = side.cuda_stream
                    if peer is not None:
                        gf[2 * (n + Cn):] = self.stats[STAT["CNT_X1_CAL"]:STAT["CNT_X2_CAL"] + 1].to(gf.dtype)
                        ra = _lib.LfPeerReduceArgs()
                        peer.fill(ra.comm)
                        ra.buf, ra.n, ra.n_padded = _ptr(gf), gf.numel(), peer.grad_padded
                        ra.tail_dst, ra.tail_n = self.stats[STAT["CNT_X1_CAL"]:].data_ptr(), 2
                        check(lib.lf_peer_allreduce(C.byref(ra), sst), "lf_peer_allreduce")
                    else:
                        parallel.pack_grad_exchange(gf, 2 * (n + Cn), self.stats, STAT["CNT_X1_CAL"], STAT["CNT_X2_CAL"] + 1, self.pg)
                a.bwd_phase = 2
                check(lib.lf_heads_backward(C.byref(a), st), "lf_heads_backward")
                cur.wait_stream(side)
        bufs = dict(bufs, conf=p_conf if qmf else None)

        return StepOutput(
            logits=[bufs["logits"][0], bufs["logits"][1]], avg_logits=bufs["avg"], logits_df=bufs["zdf"],
            conf=bufs["conf"], loss=self.loss[0],
            dfeat=[bufs["dfeat"][0], bufs["dfeat"][1]] if (need_dfeat and backward) else [None, None],
            dweight=dW, dbias=db, stats=self.stats, batch_global=Bg)

    # ------------------------------------------------------------------ CUDA graph
    def capture(self, feats, weights, biases, label, idx=None, extra=None, warmup: int = 2, **kw):
        """Capture one step (plus ``extra()``, e.g. the OGM-GE modulation calls) reading the GIVEN tensors into a
        CUDA graph.  ``graph.replay()`` then re-runs the step on whatever those tensors hold at that time; the
        returned StepOutput aliases the engine's static buffers.  All device state the step mutates (EMA,
        History, ticket counter) lives in device memory, so replays advance it exactly like eager calls.
        The ``warmup`` eager steps needed before capture also advance that state."""
        if self.fresh_outputs:
            raise _lib.LfError("capture() needs static buffers (fresh_outputs=False)")
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                out = self.step(feats, weights, biases, label, idx=idx, **kw)
                if extra is not None:
                    extra()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        self._capturing = True
        try:
            with torch.cuda.graph(graph):
                out = self.step(feats, weights, biases, label, idx=idx, **kw)
                if extra is not None:
                    extra()
        finally:
            self._capturing = False
        return graph, out

    # ------------------------------------------------------------------ host-fed steps
    def stream_from_host(self, host_batches, weights, biases, extra=None, **kw):
        """Generator over ``host_batches`` (iterable of dicts of PINNED host tensors with keys f1, f2, y[, idx]):
        yields the loss of every step as a Python float.  Each batch is copied host->device on a copy stream into
        one of two staging sets while the previous step computes; the step itself is a CUDA-graph replay on the
        staging set (one graph per set).  The loss is read back (device->host) every step, as the reference's
        training loop does when it logs it."""
        cur = torch.cuda.current_stream()
        copy = torch.cuda.Stream(device=self.device)
        it = iter(host_batches)
        stages, graphs, ready, done = [], [], [], []
        loss_host = torch.empty(1, dtype=torch.float32).pin_memory()

        def upload(slot, hb):
            if done[slot] is not None:
                copy.wait_event(done[slot])              # the step that last read this staging set has finished
            with torch.cuda.stream(copy):
                for k, v in stages[slot].items():
                    v.copy_(hb[k], non_blocking=True)
                ready[slot].record(copy)

        first = next(it, None)
        if first is None:
            return
        for slot in range(2):
            st = {k: torch.empty_like(v, device=self.device) for k, v in first.items()}
            stages.append(st); ready.append(torch.cuda.Event()); done.append(None)
            for k, v in st.items():
                v.copy_(first[k])
            graphs.append(self.capture([st["f1"], st["f2"]], weights, biases, st["y"], idx=st.get("idx"), extra=extra, **kw))
        copy.wait_stream(cur)
        upload(0, first)
        slot, nxt = 0, next(it, None)
        while True:
            cur.wait_event(ready[slot])
            if nxt is not None:
                upload(1 - slot, nxt)                    # overlaps the replay below
            g, out = graphs[slot]
            g.replay()
            done[slot] = torch.cuda.Event(); done[slot].record(cur)
            loss_host.copy_(out.loss.view(1), non_blocking=True)
            cur.synchronize()
            yield float(loss_host[0])
            if nxt is None:
                return
            slot, nxt = 1 - slot, next(it, None)

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
