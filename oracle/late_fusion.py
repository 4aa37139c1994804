"""CPU oracle for the per-batch late-fusion training step.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU with plain torch/numpy ops, the arithmetic that the reference
(Nano1337/multimodal-clinical) performs between "per-modality feature vectors" and "gradients on the
head weights / features plus the algorithm state updates".  It exists to CHECK the CUDA path; nothing
under ``multimodal_clinical_b200/`` may import it.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it.

Parity pinning: the reference ships no tests or golden vectors.  This oracle is
pinned instead against outputs of the UNMODIFIED reference modules imported from /root/reference
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``).

Reference lines followed (paths relative to the reference tree):
  heads + mean fusion + CE .......... cremad/joint_model_ogm_ge.py:50-58, enrico/joint_model.py:49-52,78-86
  QMF energy fusion ................. existing_algos/QMF.py:109-117
  QMF loss assembly ................. cremad/joint_model_qmf.py:57-75, food101/joint_model_qmf.py:61-81
  History update / normalise / pair . existing_algos/QMF.py:20-29, 37-42, 45-68
  ranking regulariser ............... existing_algos/QMF.py:119-141
  OGM-GE scores / coefficients ...... existing_algos/OGM_GE.py:21-40
  OGM-GE gradient modulation ........ existing_algos/OGM_GE.py:42-57
  per-modality ensemble ............. cremad/ensemble_model_noised.py:49-55, 97-120
  epoch-end offset correction ....... utils/BaseModel.py:168-185
  EMA logit offsets ................. utils/EMA.py:29-38, utils/BaseModel.py:82-85
  step metrics ...................... utils/BaseModel.py:78-92, 946-961

Backward passes are obtained with torch autograd over the restated forward, so they are independent of
the analytic backward that the CUDA kernels implement.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

HISTORY_ALPHA = 0.1      # existing_algos/QMF.py:17 (hard-coded, use_ema=True at :16)
EMA_SMOOTHING = 0.05     # utils/EMA.py:24
CONF_DIVISOR = 10.0      # existing_algos/QMF.py:114


# ----------------------------------------------------------------------------------------------
# heads
# ----------------------------------------------------------------------------------------------
def heads_forward(feats: Sequence[torch.Tensor], weights: Sequence[torch.Tensor],
                  biases: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """z_m = f_m W_m^T + b_m  (nn.Linear; cremad/joint_model_qmf.py:57-58)."""
    return [F.linear(f, w, b) for f, w, b in zip(feats, weights, biases)]


def cross_entropy_mean(z: torch.Tensor, y: torch.Tensor, batch_global: Optional[int] = None) -> torch.Tensor:
    """nn.CrossEntropyLoss() with mean reduction (cremad/joint_model_qmf.py:91).

    ``batch_global`` lets a shard compute its contribution to the global-batch mean."""
    if batch_global is None:
        return F.cross_entropy(z, y)
    return F.cross_entropy(z, y, reduction="sum") / batch_global


# ----------------------------------------------------------------------------------------------
# QMF history (numpy fp64, like the reference)
# ----------------------------------------------------------------------------------------------
@dataclass
class HistoryState:
    """Per-modality running 'correctness' / 'confidence' arrays (existing_algos/QMF.py:12-17)."""
    n_data: int
    n_modality: int = 2
    correctness: np.ndarray = field(default=None)   # (M, N) float64
    confidence: np.ndarray = field(default=None)    # (M, N) float64

    def __post_init__(self):
        if self.correctness is None:
            self.correctness = np.zeros((self.n_modality, self.n_data), dtype=np.float64)
        if self.confidence is None:
            self.confidence = np.zeros((self.n_modality, self.n_data), dtype=np.float64)

    def clone(self) -> "HistoryState":
        return HistoryState(self.n_data, self.n_modality, self.correctness.copy(), self.confidence.copy())


def history_update(hist: HistoryState, m: int, idx: np.ndarray, loss_uni: float, conf: np.ndarray) -> None:
    """corr[idx] <- 0.9 corr[idx] + 0.1 L (one scalar for the whole batch); confid[idx] <- conf.

    existing_algos/QMF.py:20-29.  Gather-then-scatter: duplicates in ``idx`` all write the same value;
    for ``confidence`` the last duplicate wins (numpy assignment order).  The product 0.1*L is taken in
    float64 (numpy 1.26.4, pinned by the reference's requirements.txt:71, promotes a python float times
    a 0-d float32 array to float64)."""
    c = hist.correctness[m]
    c[idx] = (1.0 - HISTORY_ALPHA) * c[idx] + HISTORY_ALPHA * float(loss_uni)
    hist.confidence[m][idx] = conf.astype(np.float64)


def history_target_margin(hist: HistoryState, m: int, idx1: np.ndarray, idx2: np.ndarray
                          ) -> Tuple[np.ndarray, np.ndarray]:
    """Min-max normalised correctness pairs -> (target in {-1,0,1}, |margin|) as float32.

    existing_algos/QMF.py:37-42 (global min/max over all N) and :45-68."""
    c = hist.correctness[m]
    lo = c.min()
    hi = float(c.max())
    with np.errstate(invalid="ignore", divide="ignore"):
        a = (c[idx1] - lo) / (hi - lo)
        b = (c[idx2] - lo) / (hi - lo)
    target = (a > b).astype(np.float64) - (a < b).astype(np.float64)
    margin = np.abs(a - b)
    return target.astype(np.float32), margin.astype(np.float32)


# ----------------------------------------------------------------------------------------------
# QMF fusion + regulariser
# ----------------------------------------------------------------------------------------------
def qmf_df(z: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Energy-confidence fusion (existing_algos/QMF.py:109-117).  z: (M,B,C).

    The energy is the NON-stabilised log(sum(exp(z))) exactly as the reference writes it; the fusion
    weights are detached."""
    energy = torch.log(torch.sum(torch.exp(z), dim=-1))
    conf = energy / CONF_DIVISOR
    z_df = (z * conf.unsqueeze(-1).detach()).sum(dim=0)
    return z_df, conf


def _nan_relu(x: torch.Tensor) -> torch.Tensor:
    return torch.clamp_min(x, 0.0)          # clamp propagates NaN like MarginRankingLoss does


def qmf_reg_loss_closed(conf: torch.Tensor, idx: np.ndarray, hist: HistoryState,
                        batch_global: Optional[int] = None) -> torch.Tensor:
    """Closed form of QMF.reg_loss (existing_algos/QMF.py:119-141); O(B) instead of O(B^2).

    The reference builds two (B,B) broadcast matrices but only ever uses row n of the n-th one, and
    rolls the FLATTENED (M,B) confidence, so:
      r0[j] = conf.flatten()[j+1]                      (j = B-1 wraps into modality 1, sample 0)
      s0 = margin0[0]/tnz0[0];  s1 = margin0[0]/tnz0[1] + margin1[1]/tnz1[1]
      L  = mean_j relu(t0[j] (c0[j]-r0[j]-s0)) + mean_j relu(t1[j] (c1[j]-r0[j]-s1))
    Requires M == 2 and B >= 2 (the reference raises for B == 1)."""
    M, B = conf.shape
    assert M == 2 and B >= 2
    idx2 = np.roll(idx, -1)
    t0, m0 = history_target_margin(hist, 0, idx, idx2)
    t1, m1 = history_target_margin(hist, 1, idx, idx2)
    t0 = torch.from_numpy(t0).to(conf.dtype)
    t1 = torch.from_numpy(t1).to(conf.dtype)
    m0 = torch.from_numpy(m0).to(conf.dtype)
    m1 = torch.from_numpy(m1).to(conf.dtype)
    tnz0 = torch.where(t0 == 0, torch.ones_like(t0), t0)
    tnz1 = torch.where(t1 == 0, torch.ones_like(t1), t1)
    r0 = torch.roll(conf, -1)[0]                 # no dim -> flattened roll (QMF.py:125)
    s0 = m0[0] / tnz0[0]
    q0 = m0[0] / tnz0[1]
    q1 = m1[1] / tnz1[1]
    denom = float(B if batch_global is None else batch_global)
    l0 = _nan_relu(t0 * (conf[0] - (r0 + s0))).sum() / denom
    l1 = _nan_relu(t1 * (conf[1] - ((r0 + q0) + q1))).sum() / denom      # same association as the reference
    return l0 + l1


def qmf_reg_loss_literal(conf: torch.Tensor, idx: np.ndarray, hist: HistoryState) -> torch.Tensor:
    """The same loss evaluated the long way round — the (B,B) broadcast and the row pick — for small B.

    Used only to cross-check the closed form; follows existing_algos/QMF.py:124-141 step by step."""
    M, B = conf.shape
    first = conf
    second = torch.roll(conf, -1)
    idx2 = np.roll(idx, -1)
    terms = []
    for n in range(M):
        t, mg = history_target_margin(hist, n, idx, idx2)
        t = torch.from_numpy(t).to(conf.dtype)
        mg = torch.from_numpy(mg).to(conf.dtype)
        tnz = t.clone()
        tnz[tnz == 0] = 1
        second = second[n] + (mg[n] / tnz).reshape(-1, 1)       # (B,) + (B,1) -> (B,B)
        terms.append(F.margin_ranking_loss(first[n], second[n], -t, margin=0.0))
    return torch.stack(terms).sum()


# ----------------------------------------------------------------------------------------------
# OGM-GE
# ----------------------------------------------------------------------------------------------
def ogm_scores(z1: torch.Tensor, z2: torch.Tensor, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """s_m = sum_b softmax(z_m)[b, y_b]  (existing_algos/OGM_GE.py:21-22; the reference loop is O(B^2 C)
    but evaluates exactly this)."""
    s1 = torch.softmax(z1, dim=-1).gather(1, y.view(-1, 1)).sum()
    s2 = torch.softmax(z2, dim=-1).gather(1, y.view(-1, 1)).sum()
    return s1, s2


def ogm_coeffs(score1: float, score2: float, alpha: float) -> Tuple[float, float]:
    """existing_algos/OGM_GE.py:24-40.  Returns (coeff for x1_model, coeff for x2_model)."""
    ratio1 = score1 / score2
    ratio2 = 1.0 / ratio1
    if ratio1 > 1:
        return 1.0 - math.tanh(alpha * max(ratio1, 0.0)), 1.0
    return 1.0, 1.0 - math.tanh(alpha * max(ratio2, 0.0))


def ogm_modulate(grads: Sequence[torch.Tensor], coeff: float, modulation: str,
                 noise: Optional[Sequence[torch.Tensor]] = None) -> List[torch.Tensor]:
    """existing_algos/OGM_GE.py:42-54 for a list of 4-D gradients of one encoder.

    ``noise`` supplies the standard-normal draws xi (same shapes); None -> xi = 0 so that the
    deterministic part (scale, sigma) can be compared exactly.  sigma = unbiased std of the UNSCALED
    gradient + 1e-8."""
    out = []
    for i, g in enumerate(grads):
        if g.dim() != 4:
            out.append(g.clone())
            continue
        sigma = g.std().item() + 1e-8
        xi = noise[i] if noise is not None else torch.zeros_like(g)
        if modulation == "OGM_GE":
            out.append(g * coeff + xi * sigma)
        elif modulation == "OGM":
            out.append(g * coeff)
        elif modulation == "noise":
            out.append(g + xi * sigma)
        else:
            out.append(g.clone())
    return out


# ----------------------------------------------------------------------------------------------
# EMA offsets + metrics
# ----------------------------------------------------------------------------------------------
def ema_update(x: torch.Tensor, z1: torch.Tensor, z2: torch.Tensor,
               batch_global: Optional[int] = None) -> torch.Tensor:
    """x <- 0.05 mean_b(z_m) + 0.95 x  (utils/EMA.py:29-34, call site utils/BaseModel.py:82-83)."""
    denom = float(z1.shape[0] if batch_global is None else batch_global)
    x_new = torch.stack([z1.detach().sum(0) / denom, z2.detach().sum(0) / denom])
    return x_new * EMA_SMOOTHING + x * (1.0 - EMA_SMOOTHING)


def ema_offset(x: torch.Tensor) -> torch.Tensor:
    """offset = mean_m(x) - x  (utils/EMA.py:36-38)."""
    return x.mean(dim=0, keepdim=True) - x


def correct_count(z: torch.Tensor, y: torch.Tensor) -> int:
    return int((torch.argmax(z, dim=1) == y).sum().item())


# ----------------------------------------------------------------------------------------------
# whole steps
# ----------------------------------------------------------------------------------------------
def _leafs(feats, weights, biases, dtype, feat_grad):
    fs = [f.detach().to(dtype).clone().requires_grad_(feat_grad) for f in feats]
    ws = [w.detach().to(dtype).clone().requires_grad_(True) for w in weights]
    bs = [b.detach().to(dtype).clone().requires_grad_(True) for b in biases]
    return fs, ws, bs


def _finish(out: Dict, fs, ws, bs, zs, y, ema_x, feat_grad):
    loss = out["loss"]
    loss.backward()
    out["dW"] = [w.grad for w in ws]
    out["db"] = [b.grad for b in bs]
    out["dfeat"] = [f.grad for f in fs] if feat_grad else [None, None]
    z1, z2 = zs[0].detach(), zs[1].detach()
    B = z1.shape[0]
    if ema_x is not None:
        x = ema_update(ema_x.to(z1.dtype), z1, z2)
        off = ema_offset(x)
        out["ema_x"] = x
        out["ema_offset"] = off
        out["acc_x1_cal"] = correct_count(z1 + off[0], y) / B
        out["acc_x2_cal"] = correct_count(z2 + off[1], y) / B
    out["acc_x1_uncal"] = correct_count(z1, y) / B
    out["acc_x2_uncal"] = correct_count(z2, y) / B
    out["acc_joint"] = correct_count(out["avg_logits"].detach(), y) / B
    s1, s2 = ogm_scores(z1, z2, y)
    out["score1"], out["score2"] = float(s1), float(s2)
    out["loss"] = loss.detach()
    out["logits"] = [z1, z2]
    out["avg_logits"] = out["avg_logits"].detach()
    return out


def jlogits_step(feats, weights, biases, y, ema_x=None, dtype=torch.float32, feat_grad=True) -> Dict:
    """Plain late fusion / OGM-GE head step: avg=(z1+z2)/2, L=CE(avg)
    (cremad/joint_model_ogm_ge.py:50-58; enrico/joint_model.py:78-86) + EMA + metrics
    (utils/BaseModel.py:72-92)."""
    fs, ws, bs = _leafs(feats, weights, biases, dtype, feat_grad)
    zs = heads_forward(fs, ws, bs)
    avg = (zs[0] + zs[1]) / 2
    loss = cross_entropy_mean(avg, y)
    out = {"loss": loss, "avg_logits": avg}
    return _finish(out, fs, ws, bs, zs, y, ema_x, feat_grad)


def mean_fusion_multi_step(feats, weights, biases, y, dtype=torch.float32, feat_grad=True) -> Dict:
    """Mean fusion of M heads with per-modality feature widths: avg = (z_1 + ... + z_M) / M, L = CE(avg)
    (mustard/joint_model.py:72-83 with M = 3; avmnist/joint_model.py:128-138 with D = 48 / 192) and the joint accuracy the
    Lightning modules log (mustard/joint_model.py:134, avmnist/joint_model.py:201)."""
    fs, ws, bs = _leafs(feats, weights, biases, dtype, feat_grad)
    zs = heads_forward(fs, ws, bs)
    acc = zs[0]
    for z in zs[1:]:
        acc = acc + z
    avg = acc / len(zs)
    loss = cross_entropy_mean(avg, y)
    loss.backward()
    B = y.shape[0]
    return {"loss": loss.detach(), "avg_logits": avg.detach(), "logits": [z.detach() for z in zs],
            "dW": [w.grad for w in ws], "db": [b.grad for b in bs], "dfeat": [f.grad for f in fs] if feat_grad else [None] * len(fs),
            "acc_joint": correct_count(avg.detach(), y) / B, "acc_modality": [correct_count(z.detach(), y) / B for z in zs]}


def ensemble_step(feats, weights, biases, y, dtype=torch.float32, feat_grad=True) -> Dict:
    """Per-modality ensemble: x_m_loss = CE(z_m, y), backward of (x1_loss + x2_loss) / 2
    (cremad/ensemble_model_noised.py:49-55, 97-103, 119-120).  No EMA calibration in this family
    (utils/BaseModel.py:291-562); the joint accuracy is that of (z1 + z2) / 2."""
    fs, ws, bs = _leafs(feats, weights, biases, dtype, feat_grad)
    zs = heads_forward(fs, ws, bs)
    l1, l2 = cross_entropy_mean(zs[0], y), cross_entropy_mean(zs[1], y)
    out = {"loss": (l1 + l2) / 2, "avg_logits": (zs[0] + zs[1]) / 2, "loss_x1": l1.detach(), "loss_x2": l2.detach()}
    return _finish(out, fs, ws, bs, zs, y, None, feat_grad)


def epoch_offset_correction(logits: torch.Tensor, labels: torch.Tensor) -> Dict:
    """Epoch-end unimodal offset correction over all collected logits (N, M, C) (utils/BaseModel.py:168-185):
    offset = mean_m(mean_n logits) - mean_n logits; accuracies of the raw and of the corrected unimodal logits."""
    m_out = torch.mean(logits, dim=0)
    offset = torch.mean(m_out, dim=0, keepdim=True) - m_out
    corrected = logits + offset
    N = logits.shape[0]
    return {"offset": offset,
            "x1_acc_uncal": correct_count(logits[:, 0, :], labels) / N, "x2_acc_uncal": correct_count(logits[:, 1, :], labels) / N,
            "x1_acc": correct_count(corrected[:, 0, :], labels) / N, "x2_acc": correct_count(corrected[:, 1, :], labels) / N}


LOSS_NO_JOINT, LOSS_NO_UNI = 1, 2     # the LF_LOSS_* bits of include/lf_fusion.h


def qmf_step(feats, weights, biases, y, idx, hist: HistoryState, ema_x=None, dtype=torch.float32,
             feat_grad=True, literal_reg=False, loss_terms=0) -> Dict:
    """QMF head step (cremad/joint_model_qmf.py:57-75).  Mutates ``hist`` like the reference does:
    the History update precedes the regulariser inside the same forward (:63-67).
    ``loss_terms``: LOSS_NO_JOINT = ``loss_joint = 0`` (cremad/joint_model_qmf_ablate_Ljoint.py:68),
    LOSS_NO_UNI = unimodal sum dropped from the loss (cremad/joint_model_qmf_ablate_Lunimodal.py:70)."""
    fs, ws, bs = _leafs(feats, weights, biases, dtype, feat_grad)
    zs = heads_forward(fs, ws, bs)
    z = torch.stack(zs)
    z_df, conf = qmf_df(z)
    idx_np = idx.detach().cpu().numpy().astype(np.int64)
    loss_uni = []
    for m in range(2):
        lu = cross_entropy_mean(z[m], y)
        loss_uni.append(lu)
        # the reference hands the fp32 loss tensor over (.cpu().numpy()); round through fp32 likewise
        history_update(hist, m, idx_np, float(lu.detach().to(torch.float32)),
                       conf[m].detach().to(torch.float32).numpy())
    reg = (qmf_reg_loss_literal if literal_reg else qmf_reg_loss_closed)(conf, idx_np, hist)
    loss_joint = cross_entropy_mean(z_df, y)
    loss = (0 if loss_terms & LOSS_NO_JOINT else loss_joint) + \
           (0 if loss_terms & LOSS_NO_UNI else torch.stack(loss_uni).sum()) + reg
    avg = (zs[0] + zs[1]) / 2
    out = {"loss": loss, "avg_logits": avg}
    out["logits_df"] = z_df.detach()
    out["conf"] = conf.detach()
    out["loss_uni"] = [float(l.detach()) for l in loss_uni]
    out["loss_joint"] = float(loss_joint.detach())
    out["loss_reg"] = float(reg.detach())
    out["acc_df"] = correct_count(z_df.detach(), y) / z_df.shape[0]
    return _finish(out, fs, ws, bs, zs, y, ema_x, feat_grad)


def make_inputs(B: int, D: int, C: int, seed: int = 5, n_data: Optional[int] = None,
                idx_mode: str = "replacement") -> Dict[str, torch.Tensor]:
    """Synthetic inputs per SURVEY.md §8(d): N(0,1) features, nn.Linear default init, uniform labels."""
    g = torch.Generator().manual_seed(seed)
    bound = 1.0 / math.sqrt(D)
    out = {
        "f1": torch.randn(B, D, generator=g), "f2": torch.randn(B, D, generator=g),
        "W1": (torch.rand(C, D, generator=g) * 2 - 1) * bound, "b1": (torch.rand(C, generator=g) * 2 - 1) * bound,
        "W2": (torch.rand(C, D, generator=g) * 2 - 1) * bound, "b2": (torch.rand(C, generator=g) * 2 - 1) * bound,
        "y": torch.randint(0, C, (B,), generator=g, dtype=torch.int64),
    }
    if n_data is not None:
        if idx_mode == "replacement":
            out["idx"] = torch.randint(0, n_data, (B,), generator=g, dtype=torch.int64)
        else:
            start = int(torch.randint(0, max(n_data - B, 1), (1,), generator=g))
            out["idx"] = (torch.arange(B, dtype=torch.int64) + start) % n_data
    return out
