"""Mean fusion on wide heads: lf_step_mid + the calibrated-count pass run on a second stream beside the dfeat GEMM
(LfHeadsArgs.bwd_phase 3 / 2 / 4, multimodal_clinical_b200/step.py).  Same kernels, same arithmetic: every output must
be BIT-identical to the single-stream order, eagerly and replayed from a CUDA graph, and the phases must refuse the
shapes whose dL/dz is formed inside a fused kernel."""
import ctypes as C

import pytest
import torch

from oracle import late_fusion as O

pytestmark = pytest.mark.gpu


def _run(overlap, prec, B, D, Cn, steps=3, graph=False, sgd=False):
    from multimodal_clinical_b200.step import LateFusionStep
    eng = LateFusionStep(Cn, mode="jlogits", device="cuda:0", precision=prec)
    eng.cal_overlap = overlap
    eng.cal_overlap_min_classes = 0        # (by default only heads wider than 128 classes take the two-stream order)
    inp = O.make_inputs(B, D, Cn, seed=11)
    W = [inp["W1"].cuda(), inp["W2"].cuda()]
    b = [inp["b1"].cuda(), inp["b2"].cuda()]
    if sgd:
        eng.enable_sgd(lr=1e-2)
    f = [inp["f1"].cuda(), inp["f2"].cuda()]
    if prec == "bf16":
        f = [x.bfloat16() for x in f]
    y = inp["y"].cuda()
    if graph:
        g, out = eng.capture(f, W, b, y, ogm_alpha=0.7, warmup=1)
        for _ in range(steps - 1):
            g.replay()
    else:
        for _ in range(steps):
            out = eng.step(f, W, b, y, ogm_alpha=0.7)
    torch.cuda.synchronize()
    return {"loss": out.loss.clone(), "stats": out.stats.clone(), "dW1": out.dweight[0].clone(), "db2": out.dbias[1].clone(),
            "df1": out.dfeat[0].clone(), "df2": out.dfeat[1].clone(), "ema": eng.ema_x.clone(), "off": eng.ema_offset.clone(),
            "coeff": eng.coeff.clone(), "W1": W[0].clone()}


@pytest.mark.parametrize("prec,B,D,Cn", [("bf16", 3000, 512, 309), ("tf32", 1500, 512, 309), ("bf16", 2048, 768, 101),
                                         ("fp32", 700, 256, 101), ("fp32", 333, 64, 40)])
def test_overlapped_schedule_is_bit_identical(prec, B, D, Cn):
    a = _run(True, prec, B, D, Cn)
    s = _run(False, prec, B, D, Cn)
    for k in a:
        assert torch.equal(a[k], s[k]), k
    # the calibrated counts did come from this step's offsets: non-zero and at most B
    from multimodal_clinical_b200._lib import STAT
    assert 0 <= float(a["stats"][STAT["CNT_X1_CAL"]]) <= B


def test_overlapped_schedule_in_a_cuda_graph_with_in_step_sgd():
    a = _run(True, "bf16", 4096, 512, 309, steps=4, graph=True, sgd=True)
    s = _run(False, "bf16", 4096, 512, 309, steps=4, graph=False, sgd=True)
    for k in a:
        assert torch.equal(a[k], s[k]), k


def test_split_phases_refused_where_dz_is_formed_in_a_fused_kernel():
    from multimodal_clinical_b200 import _lib
    lib = _lib.load()
    a = _lib.LfHeadsArgs()
    a.batch, a.batch_global, a.dim, a.classes = 256, 256, 512, 6
    a.mode, a.precision = _lib.LF_MODE_JLOGITS, _lib.LF_PREC_FP32
    assert lib.lf_heads_backward_splits_rows(C.byref(a)) == 0          # narrow heads: one fused pass
    a.classes, a.dim, a.mode, a.precision, a.need_dfeat = 101, 768, _lib.LF_MODE_QMF, _lib.LF_PREC_BF16, 1
    a.ld_logits, a.ld_dlogits = 104, 104
    assert lib.lf_heads_backward_splits_rows(C.byref(a)) == 0          # fused QMF backward (tc_backward_qmf)
    a.classes, a.dim, a.mode, a.ld_logits, a.ld_dlogits = 309, 512, _lib.LF_MODE_JLOGITS, 312, 312
    assert lib.lf_heads_backward_splits_rows(C.byref(a)) == 1
