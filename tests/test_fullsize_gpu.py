"""GPU parity AT THE BENCHED SHAPES: the BASELINE.json configurations in the precision bench.py times them in,
against the CPU oracle run in fp64 on the same (bf16-rounded where the mode rounds) inputs.

  K4  Food101 QMF      B = 32768, D = 768, C = 101, N = 65536, bf16, contiguous idx windows (bench.py's make_batches)
  K5  VGGSound shape   B = 131072, D = 512, C = 309, mean fusion + OGM-GE coefficients, bf16 and tf32
  K3  Crema-D OGM-GE   B = 8192, D = 512, C = 6, exact fp32, with the `OGM` modulation of the encoder gradients

Tolerances are BASELINE.json's: 2e-2 (bf16 / tf32 tensor pipe), 1e-5 (fp32).  Integer results (accuracy counts)
are exact functions of the logits the kernel itself returned, so they are re-derived from those on the device.
The oracle is O(B): the K5 comparison costs ~20 s of host time."""
import numpy as np
import pytest
import torch

from oracle import late_fusion as O
from tests.util import assert_close, TOL_FP32, TOL_TENSOR

pytestmark = pytest.mark.gpu


def _eng(**kw):
    from multimodal_clinical_b200.step import LateFusionStep
    return LateFusionStep(device="cuda:0", **kw)


def _counts_from_logits(out, eng, y, qmf):
    """The six accuracy counts recomputed with torch from the logits the step returned (exact integers)."""
    z1, z2 = out.logits[0], out.logits[1]
    cnt = lambda z: int((z.argmax(1) == y).sum())
    want = {"CNT_X1": cnt(z1), "CNT_X2": cnt(z2), "CNT_JOINT": cnt(out.avg_logits),
            "CNT_X1_CAL": cnt(z1 + eng.ema_offset[0]), "CNT_X2_CAL": cnt(z2 + eng.ema_offset[1])}
    if qmf:
        want["CNT_DF"] = cnt(out.logits_df)
    return want


def _check_counts(out, eng, y, qmf, slack=0):
    from multimodal_clinical_b200._lib import STAT
    st = out.stats.cpu()
    for k, v in _counts_from_logits(out, eng, y, qmf).items():
        got = float(st[STAT[k]])
        assert got == round(got)
        # calibrated counts add a fp32 offset to the logits: an exact tie in z + off may round either way on the
        # two sides (torch adds then compares; the kernel does the same adds) -> `slack` samples of B
        assert abs(got - v) <= (slack if "CAL" in k else 0), (k, got, v)


def _bench_idx(B, N, g):
    """bench.py::make_batches: contiguous window (Food101's loaders are not shuffled, food101/run_training.py:39-45)."""
    return (torch.arange(B, dtype=torch.int64) + int(torch.randint(0, N, (1,), generator=g))) % N


def test_k4_bf16_full_size_matches_fp64_oracle():
    B, D, C, N = 32768, 768, 101, 65536
    g = torch.Generator().manual_seed(404)
    base = O.make_inputs(8, D, C, seed=5)
    r16 = lambda x: x.bfloat16().float()
    W = [r16(base["W1"]), r16(base["W2"])]
    b = [base["b1"], base["b2"]]
    eng = _eng(num_classes=C, mode="qmf", n_data=N, precision="bf16")
    hist = O.HistoryState(N)
    ema = torch.zeros(2, C, dtype=torch.float64)
    Wd = [base["W1"].cuda(), base["W2"].cuda()]; bd = [x.cuda() for x in b]
    for s in range(3):
        f = [r16(torch.randn(B, D, generator=g)), r16(torch.randn(B, D, generator=g))]
        y = torch.randint(0, C, (B,), generator=g, dtype=torch.int64)
        idx = _bench_idx(B, N, g)
        ref = O.qmf_step(f, W, b, y, idx, hist, ema_x=ema, dtype=torch.float64)
        ema = ref["ema_x"]
        out = eng.step([x.cuda().bfloat16() for x in f], Wd, bd, y.cuda(), idx=idx.cuda())
        torch.cuda.synchronize()
        assert_close(out.loss, ref["loss"], TOL_TENSOR, f"loss step {s}")
        assert abs(float(out.loss) - float(ref["loss"])) < 2e-3, (float(out.loss), float(ref["loss"]))
        for m in range(2):
            assert_close(out.logits[m], ref["logits"][m], TOL_TENSOR, f"z{m + 1}")
            assert_close(out.dweight[m], ref["dW"][m], TOL_TENSOR, f"dW{m + 1} step {s}")
            assert_close(out.dbias[m], ref["db"][m], TOL_TENSOR, f"db{m + 1} step {s}")
            assert_close(out.dfeat[m].float(), ref["dfeat"][m], TOL_TENSOR, f"df{m + 1} step {s}")
        assert_close(out.avg_logits, ref["avg_logits"], TOL_TENSOR, "avg")
        assert_close(out.logits_df, ref["logits_df"], TOL_TENSOR, "zdf")
        assert_close(out.conf, ref["conf"], TOL_TENSOR, "conf")
        assert_close(eng.ema_x, ref["ema_x"], TOL_TENSOR, "ema_x")
        assert_close(eng.correctness, hist.correctness, 1e-4, "history.correctness")
        assert_close(eng.confidence, hist.confidence, TOL_TENSOR, "history.confidence")
        # every index of the window was updated exactly once; untouched entries are still exactly 0
        touched = torch.zeros(N, dtype=torch.bool); touched[idx] = True
        if s == 0:
            assert torch.all(eng.correctness[:, ~touched.cuda()] == 0)
        _check_counts(out, eng, y.cuda(), qmf=True, slack=2)


@pytest.mark.parametrize("prec", ["bf16", "tf32"])
def test_k5_full_size_matches_fp64_oracle(prec):
    B, D, C = 131072, 512, 309
    g = torch.Generator().manual_seed(505)
    base = O.make_inputs(8, D, C, seed=5)
    rnd = (lambda x: x.bfloat16().float()) if prec == "bf16" else (lambda x: x)
    W = [rnd(base["W1"]), rnd(base["W2"])]
    b = [base["b1"], base["b2"]]
    f = [rnd(torch.randn(B, D, generator=g)), rnd(torch.randn(B, D, generator=g))]
    y = torch.randint(0, C, (B,), generator=g, dtype=torch.int64)
    eng = _eng(num_classes=C, mode="jlogits", precision=prec)
    ref = O.jlogits_step(f, W, b, y, ema_x=torch.zeros(2, C, dtype=torch.float64), dtype=torch.float64)
    fd = [x.cuda().bfloat16() if prec == "bf16" else x.cuda() for x in f]
    out = eng.step(fd, [base["W1"].cuda(), base["W2"].cuda()], [x.cuda() for x in b], y.cuda(), ogm_alpha=0.8)
    torch.cuda.synchronize()
    assert_close(out.loss, ref["loss"], TOL_TENSOR, "loss")
    assert abs(float(out.loss) - float(ref["loss"])) < 2e-3
    for m in range(2):
        assert_close(out.logits[m], ref["logits"][m], TOL_TENSOR, f"z{m + 1}")
        assert_close(out.dweight[m], ref["dW"][m], TOL_TENSOR, f"dW{m + 1}")
        assert_close(out.dbias[m], ref["db"][m], TOL_TENSOR, f"db{m + 1}")
        assert_close(out.dfeat[m].float(), ref["dfeat"][m], TOL_TENSOR, f"df{m + 1}")
    assert_close(out.avg_logits, ref["avg_logits"], TOL_TENSOR, "avg")
    assert_close(eng.ema_x, ref["ema_x"], TOL_TENSOR, "ema_x")
    k = O.ogm_coeffs(ref["score1"], ref["score2"], 0.8)
    assert abs(float(eng.coeff[0]) - k[0]) < 2e-2 and abs(float(eng.coeff[1]) - k[1]) < 2e-2
    _check_counts(out, eng, y.cuda(), qmf=False, slack=4)


def test_k3_full_size_with_ogm_modulation_matches_oracle():
    B, D, C = 8192, 512, 6
    inp = O.make_inputs(B, D, C, seed=33)
    eng = _eng(num_classes=C, mode="jlogits")
    ema = torch.zeros(2, C)
    for s in range(2):
        f = [inp["f1"] + 0.1 * s, inp["f2"] - 0.05 * s]
        ref = O.jlogits_step(f, [inp["W1"], inp["W2"]], [inp["b1"], inp["b2"]], inp["y"], ema_x=ema, dtype=torch.float64)
        ema = ref["ema_x"].float()
        out = eng.step([x.cuda() for x in f], [inp["W1"].cuda(), inp["W2"].cuda()], [inp["b1"].cuda(), inp["b2"].cuda()],
                       inp["y"].cuda(), ogm_alpha=0.8)
        torch.cuda.synchronize()
        assert_close(out.loss, ref["loss"], TOL_FP32, "loss")
        for m in range(2):
            assert_close(out.logits[m], ref["logits"][m], TOL_FP32, f"z{m + 1}")
            assert_close(out.dweight[m], ref["dW"][m], TOL_FP32, f"dW{m + 1}")
            assert_close(out.dbias[m], ref["db"][m], TOL_FP32, f"db{m + 1}")
            assert_close(out.dfeat[m], ref["dfeat"][m], TOL_FP32, f"df{m + 1}")
        assert_close(eng.ema_x, ref["ema_x"], TOL_FP32, "ema_x")
        k = O.ogm_coeffs(ref["score1"], ref["score2"], 0.8)
        assert abs(float(eng.coeff[0]) - k[0]) < 2e-5 and abs(float(eng.coeff[1]) - k[1]) < 2e-5
        _check_counts(out, eng, inp["y"].cuda(), qmf=False, slack=0)
    # `OGM` modulation (noise disabled for the equivalence check, BASELINE.json): ResNet18 conv-shaped gradients
    shapes = [(64, 1, 7, 7), (64, 64, 3, 3), (128, 64, 1, 1), (512, 512, 3, 3)]
    for which in range(2):
        gr = [torch.randn(sh, device="cuda") * 1e-3 for sh in shapes]
        g0 = [x.clone() for x in gr]
        eng.modulate(gr, which=which, modulation="OGM", seed=5, offset=0)
        torch.cuda.synchronize()
        for a, b0 in zip(gr, g0):
            assert_close(a, b0.double() * float(eng.coeff[which]), 1e-6, "OGM scale")


def test_k4_fp32_full_size_3xtf32_matches_fp64_oracle():
    """K4 in the exact tier: LF_PREC_FP32 on a 101-way head runs the tensor-pipe kernels through the 3xTF32 operand split
    (csrc/lf_tc.cu), with the split-K dW GEMM's accumulation chain bounded by chunked TMA reduce-adds.  Contract: 1e-5."""
    B, D, C, N = 32768, 768, 101, 65536
    g = torch.Generator().manual_seed(414)
    base = O.make_inputs(8, D, C, seed=5)
    W = [base["W1"], base["W2"]]; b = [base["b1"], base["b2"]]
    eng = _eng(num_classes=C, mode="qmf", n_data=N, precision="fp32")
    hist = O.HistoryState(N)
    ema = torch.zeros(2, C, dtype=torch.float64)
    Wd = [x.cuda() for x in W]; bd = [x.cuda() for x in b]
    from multimodal_clinical_b200 import _lib
    lib = _lib.load()
    for s in range(2):
        f = [torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)]
        y = torch.randint(0, C, (B,), generator=g, dtype=torch.int64)
        idx = _bench_idx(B, N, g)
        ref = O.qmf_step(f, W, b, y, idx, hist, ema_x=ema, dtype=torch.float64)
        ema = ref["ema_x"]
        if s == 0:
            lib.lf_profile_enable(1)
        out = eng.step([x.cuda() for x in f], Wd, bd, y.cuda(), idx=idx.cuda())
        torch.cuda.synchronize()
        if s == 0:
            names = set(_lib.profile_report()); lib.lf_profile_enable(0)
            assert {"tc_logits", "tc_dfeat", "tc_dweight"} <= names and not any(n.startswith("sgemm") for n in names), names
        assert_close(out.loss, ref["loss"], TOL_FP32, f"loss step {s}")
        for m in range(2):
            assert_close(out.logits[m], ref["logits"][m], TOL_FP32, f"z{m + 1}")
            assert_close(out.dweight[m], ref["dW"][m], TOL_FP32, f"dW{m + 1} step {s}")
            assert_close(out.dbias[m], ref["db"][m], TOL_FP32, f"db{m + 1} step {s}")
            assert_close(out.dfeat[m], ref["dfeat"][m], TOL_FP32, f"df{m + 1} step {s}")
        assert_close(out.logits_df, ref["logits_df"], TOL_FP32, "zdf")
        assert_close(eng.ema_x, ref["ema_x"], TOL_FP32, "ema_x")
        assert_close(eng.correctness, hist.correctness, 1e-6, "history.correctness")
        _check_counts(out, eng, y.cuda(), qmf=True, slack=2)


def test_k5_fp32_full_size_3xtf32_matches_fp64_oracle():
    B, D, C = 131072, 512, 309
    g = torch.Generator().manual_seed(515)
    base = O.make_inputs(8, D, C, seed=5)
    W = [base["W1"], base["W2"]]; b = [base["b1"], base["b2"]]
    f = [torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)]
    y = torch.randint(0, C, (B,), generator=g, dtype=torch.int64)
    eng = _eng(num_classes=C, mode="jlogits", precision="fp32")
    ref = O.jlogits_step(f, W, b, y, ema_x=torch.zeros(2, C, dtype=torch.float64), dtype=torch.float64)
    out = eng.step([x.cuda() for x in f], [x.cuda() for x in W], [x.cuda() for x in b], y.cuda(), ogm_alpha=0.8)
    torch.cuda.synchronize()
    assert_close(out.loss, ref["loss"], TOL_FP32, "loss")
    for m in range(2):
        assert_close(out.logits[m], ref["logits"][m], TOL_FP32, f"z{m + 1}")
        assert_close(out.dweight[m], ref["dW"][m], TOL_FP32, f"dW{m + 1}")
        assert_close(out.dbias[m], ref["db"][m], TOL_FP32, f"db{m + 1}")
        assert_close(out.dfeat[m], ref["dfeat"][m], TOL_FP32, f"df{m + 1}")
    assert_close(eng.ema_x, ref["ema_x"], TOL_FP32, "ema_x")
    _check_counts(out, eng, y.cuda(), qmf=False, slack=4)
