"""OGM-GE gradient modulation (existing_algos/OGM_GE.py of the reference), without host syncs.

``ogm_ge(model, out_1, out_2, label, alpha, modulation)`` keeps the reference signature and in-place
semantics: it rescales / noises ``.grad`` of every 4-D parameter of ``model.x1_model`` and
``model.x2_model`` and returns None.  The reference spends O(B^2 C) in a Python list comprehension for
the two score sums, one host sync for the ``if ratio_v > 1`` and 40 ``.std().item()`` syncs
(SURVEY.md §0.2, §0.9); here the scores come from the fused step when ``out_1``/``out_2`` are its logits
(or from one small kernel otherwise), the coefficients stay on the device and one multi-tensor pass per
encoder computes the unbiased std and applies ``g*k + sigma*xi`` with Philox noise.
"""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from .._lib import LfTensorList, check

_MOD = {"OGM_GE": _lib.LF_MOD_OGM_GE, "OGM": _lib.LF_MOD_OGM, "noise": _lib.LF_MOD_NOISE}
_state = {}


def _scratch(device):
    key = (device.type, device.index)
    if key not in _state:
        lib = _lib.load()
        _state[key] = dict(
            stats=torch.zeros(_lib.LF_STATS_HEADER, dtype=torch.float64, device=device),
            coeff=torch.ones(2, device=device),
            score_ws=torch.zeros(lib.lf_ogm_scores_workspace_bytes(), dtype=torch.uint8, device=device),
            mod_ws=torch.empty(lib.lf_modulate_workspace_bytes(), dtype=torch.uint8, device=device),
            offset=0)
    return _state[key]


def _fused_stats(model, out_1, out_2):
    """The fused head already reduced the score sums for exactly these logits: reuse them."""
    head = getattr(model, "fused", None)
    last = getattr(head, "last_step", None)
    if last is not None and last.logits[0].data_ptr() == out_1.data_ptr() and last.logits[1].data_ptr() == out_2.data_ptr():
        return last.stats
    return None


def ogm_ge(model, out_1, out_2, label, alpha=0.1, modulation='OGM_GE'):
    if modulation not in _MOD:
        return                                      # the reference silently does nothing for other strings
    if not out_1.is_cuda:
        raise _lib.LfError("ogm_ge needs CUDA tensors; there is no CPU path")
    lib = _lib.load()
    st = _scratch(out_1.device)
    stream = torch.cuda.current_stream().cuda_stream
    stats = _fused_stats(model, out_1, out_2)
    if stats is None:
        z1 = out_1.detach().float().contiguous(); z2 = out_2.detach().float().contiguous()
        y = label.to(torch.int64).contiguous()
        stats = st["stats"]
        check(lib.lf_ogm_scores(z1.data_ptr(), z2.data_ptr(), y.data_ptr(), z1.shape[0], z1.shape[1], stats.data_ptr(),
                                st["score_ws"].data_ptr(), st["score_ws"].numel(), stream), "lf_ogm_scores")
    check(lib.lf_ogm_coeff(stats.data_ptr(), float(alpha), st["coeff"].data_ptr(), stream), "lf_ogm_coeff")
    seed = torch.initial_seed() & ((1 << 64) - 1)

    def add_factor(enc, which):
        grads = []
        for name, parms in enc.named_parameters():
            if len(parms.grad.size()) != 4:          # AttributeError on a frozen parameter, like the reference
                continue
            g = parms.grad
            if not (g.is_cuda and g.dtype == torch.float32 and g.is_contiguous()):
                raise _lib.LfError("OGM-GE modulation needs contiguous fp32 CUDA gradients")
            grads.append(g)
        for i in range(0, len(grads), _lib.LF_MAX_TENSORS):
            chunk = grads[i:i + _lib.LF_MAX_TENSORS]
            tl = LfTensorList()
            tl.count = len(chunk)
            for k, g in enumerate(chunk):
                tl.data[k] = g.data_ptr(); tl.numel[k] = g.numel()
            st["offset"] += 1
            check(lib.lf_ogm_modulate(C.byref(tl), st["coeff"][which:which + 1].data_ptr(), _MOD[modulation], seed,
                                      st["offset"], st["mod_ws"].data_ptr(), st["mod_ws"].numel(), stream),
                  "lf_ogm_modulate")

    add_factor(model.x1_model, 0)
    add_factor(model.x2_model, 1)
