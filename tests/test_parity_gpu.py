"""GPU parity: the CUDA path (through the C-ABI via LateFusionStep) against the golden vectors produced by
the unmodified reference, and against the CPU oracle on seeded inputs.  Tolerances are BASELINE.json's:
1e-5 relative (fp32 path), 2e-2 (tensor-pipe path)."""
import numpy as np
import pytest
import torch

from oracle import late_fusion as O
from tests.util import load_golden, assert_close, relerr, t, TOL_FP32, TOL_TENSOR, loss_terms_of

pytestmark = pytest.mark.gpu

QMF_CASES = ["qmf_cremad_b64", "qmf_small_c11", "qmf_food_c101", "qmf_d768_c7", "qmf_b2",
             "qmf_ablate_ljoint_b48", "qmf_ablate_lunimodal_b48", "qmf_ablate_ljoint_c101", "qmf_ogm_ge_lreg_b48"]
OGM_CASES = ["ogm_cremad_b48", "ogm_wide_c309", "jlogits_enrico_b32"]


def _step(**kw):
    from multimodal_clinical_b200.step import LateFusionStep
    return LateFusionStep(device="cuda:0", **kw)


def cu(x):
    return t(x).cuda()


@pytest.mark.parametrize("name", QMF_CASES)
def test_qmf_cuda_matches_reference_golden(name):
    g = load_golden(name)
    B, D, C, N, steps = [int(v) for v in g["meta"]]
    eng = _step(num_classes=C, mode="qmf", n_data=N, loss_terms=loss_terms_of(name))
    W = [cu(g["W1"]), cu(g["W2"])]
    b = [cu(g["b1"]), cu(g["b2"])]
    for s in range(steps):
        p = f"s{s}_"
        out = eng.step([cu(g[p + "f1"]), cu(g[p + "f2"])], W, b, cu(g[p + "y"]), idx=cu(g[p + "idx"]))
        torch.cuda.synchronize()
        assert_close(out.logits[0], g[p + "z1"], TOL_FP32, "z1")
        assert_close(out.logits[1], g[p + "z2"], TOL_FP32, "z2")
        assert_close(out.avg_logits, g[p + "avg"], TOL_FP32, "avg")
        assert_close(out.logits_df, g[p + "zdf"], TOL_FP32, "zdf")
        assert_close(out.loss, g[p + "loss"], TOL_FP32, f"loss step {s}")
        for m in range(2):
            assert_close(out.dweight[m], g[p + f"dW{m+1}"], TOL_FP32, f"dW{m+1} step {s}")
            assert_close(out.dbias[m], g[p + f"db{m+1}"], TOL_FP32, f"db{m+1} step {s}")
            assert_close(out.dfeat[m], g[p + f"df{m+1}"], TOL_FP32, f"df{m+1} step {s}")
        assert_close(eng.ema_x, g[p + "ema_x"], TOL_FP32, "ema_x")
        assert_close(eng.ema_offset, g[p + "ema_off"], 1e-4, "ema_off")
        # 1e-6: the golden ran under numpy 2.x, whose 0.1*loss product is fp32; the device (like the
        # reference's pinned numpy 1.26.4) takes it in fp64 -> ~1e-7 relative difference
        assert_close(eng.correctness, g[p + "corr"], 1e-6, "history.correctness")
        assert_close(eng.confidence, g[p + "confid"], 1e-6, "history.confidence")
        acc = out.accuracies()
        for k, gk in (("x1_acc_uncal", "acc_x1_uncal"), ("x2_acc_uncal", "acc_x2_uncal"), ("x1_acc_cal", "acc_x1_cal"),
                      ("x2_acc_cal", "acc_x2_cal"), ("joint_acc", "acc_joint"), ("df_acc", "acc_df")):
            assert abs(acc[k] - float(g[p + gk])) < 1e-6, (k, acc[k], float(g[p + gk]))


@pytest.mark.parametrize("name", OGM_CASES)
def test_jlogits_cuda_matches_reference_golden(name):
    g = load_golden(name)
    B, D, C, _, steps = [int(v) for v in g["meta"]]
    alpha = float(g["alpha"])
    eng = _step(num_classes=C, mode="jlogits")
    W = [cu(g["W1"]), cu(g["W2"])]
    b = [cu(g["b1"]), cu(g["b2"])]
    for s in range(steps):
        p = f"s{s}_"
        has_df = (p + "df1") in g
        out = eng.step([cu(g[p + "f1"]), cu(g[p + "f2"])], W, b, cu(g[p + "y"]), need_dfeat=has_df, ogm_alpha=alpha)
        torch.cuda.synchronize()
        assert_close(out.logits[0], g[p + "z1"], TOL_FP32, "z1")
        assert_close(out.logits[1], g[p + "z2"], TOL_FP32, "z2")
        assert_close(out.avg_logits, g[p + "avg"], TOL_FP32, "avg")
        assert_close(out.loss, g[p + "loss"], TOL_FP32, "loss")
        for m in range(2):
            assert_close(out.dweight[m], g[p + f"dW{m+1}"], TOL_FP32, f"dW{m+1}")
            assert_close(out.dbias[m], g[p + f"db{m+1}"], TOL_FP32, f"db{m+1}")
            if has_df:
                assert_close(out.dfeat[m], g[p + f"df{m+1}"], TOL_FP32, f"df{m+1}")
        assert_close(eng.ema_x, g[p + "ema_x"], TOL_FP32, "ema_x")
        acc = out.accuracies()
        for k, gk in (("x1_acc_uncal", "acc_x1_uncal"), ("x2_acc_uncal", "acc_x2_uncal"), ("x1_acc_cal", "acc_x1_cal"),
                      ("x2_acc_cal", "acc_x2_cal"), ("joint_acc", "acc_joint")):
            assert abs(acc[k] - float(g[p + gk])) < 1e-6, k
        if (p + "coeff") in g:
            assert_close(eng.coeff, g[p + "coeff"], 2e-5, "OGM-GE coefficients")


@pytest.mark.parametrize("B,D,C,N", [(257, 512, 6, 1000), (1000, 768, 101, 5000), (130, 128, 309, 300),
                                     (33, 36, 2, 50), (4096, 512, 6, 6698)])
def test_qmf_cuda_matches_oracle_seeded(B, D, C, N):
    """Seeded random inputs, three steps of History evolution, checked against the CPU oracle in fp64
    (the oracle's own fp32 summation noise would otherwise eat the 1e-5 budget at large B)."""
    inp = O.make_inputs(B, D, C, seed=B + C, n_data=N)
    eng = _step(num_classes=C, mode="qmf", n_data=N)
    hist = O.HistoryState(N)
    ema = torch.zeros(2, C, dtype=torch.float64)
    W = [inp["W1"], inp["W2"]]; b = [inp["b1"], inp["b2"]]
    for s in range(3):
        step_in = O.make_inputs(B, D, C, seed=100 * s + B, n_data=N)
        f = [step_in["f1"], step_in["f2"]]
        ref = O.qmf_step(f, W, b, step_in["y"], step_in["idx"], hist, ema_x=ema, dtype=torch.float64)
        ema = ref["ema_x"]
        out = eng.step([x.cuda() for x in f], [x.cuda() for x in W], [x.cuda() for x in b], step_in["y"].cuda(),
                       idx=step_in["idx"].cuda())
        torch.cuda.synchronize()
        assert_close(out.loss, ref["loss"], TOL_FP32, "loss")
        assert_close(out.logits_df, ref["logits_df"], TOL_FP32, "zdf")
        for m in range(2):
            assert_close(out.logits[m], ref["logits"][m], TOL_FP32, "logits")
            assert_close(out.dweight[m], ref["dW"][m], TOL_FP32, "dW")
            assert_close(out.dbias[m], ref["db"][m], TOL_FP32, "db")
            assert_close(out.dfeat[m], ref["dfeat"][m], TOL_FP32, "dfeat")
        assert_close(eng.correctness, hist.correctness, 1e-6, "correctness")
        assert_close(eng.ema_x, ref["ema_x"], TOL_FP32, "ema")


@pytest.mark.parametrize("B,D,C", [(8192, 512, 6), (513, 512, 20), (2048, 768, 101), (777, 512, 309), (5, 8, 2)])
def test_jlogits_cuda_matches_oracle_seeded(B, D, C):
    inp = O.make_inputs(B, D, C, seed=B + C)
    eng = _step(num_classes=C, mode="jlogits")
    ref = O.jlogits_step([inp["f1"], inp["f2"]], [inp["W1"], inp["W2"]], [inp["b1"], inp["b2"]], inp["y"],
                         ema_x=torch.zeros(2, C, dtype=torch.float64), dtype=torch.float64)
    out = eng.step([inp["f1"].cuda(), inp["f2"].cuda()], [inp["W1"].cuda(), inp["W2"].cuda()],
                   [inp["b1"].cuda(), inp["b2"].cuda()], inp["y"].cuda(), ogm_alpha=0.8)
    torch.cuda.synchronize()
    assert_close(out.loss, ref["loss"], TOL_FP32, "loss")
    for m in range(2):
        assert_close(out.logits[m], ref["logits"][m], TOL_FP32, "logits")
        assert_close(out.dweight[m], ref["dW"][m], TOL_FP32, "dW")
        assert_close(out.dbias[m], ref["db"][m], TOL_FP32, "db")
        assert_close(out.dfeat[m], ref["dfeat"][m], TOL_FP32, "dfeat")
    k = O.ogm_coeffs(ref["score1"], ref["score2"], 0.8)
    assert_close(eng.coeff, np.array(k), 2e-5, "coeff")
    acc = out.accuracies()
    assert abs(acc["joint_acc"] - ref["acc_joint"]) < 2.0 / B
    assert abs(acc["x1_acc_cal"] - ref["acc_x1_cal"]) < 2.0 / B


@pytest.mark.parametrize("B,D,C,N", [(1000, 768, 101, 5000), (4096, 768, 101, 65536), (700, 512, 309, 900)])
def test_qmf_tensor_pipe_matches_oracle(B, D, C, N):
    """LF_PREC_TF32 (tcgen05 kind::tf32 GEMMs) against the fp64 oracle at the reduced-precision tolerance
    BASELINE.json states for the reference's bf16-mixed mode (2e-2)."""
    inp = O.make_inputs(B, D, C, seed=B + C, n_data=N)
    eng = _step(num_classes=C, mode="qmf", n_data=N, precision="tf32")
    hist = O.HistoryState(N)
    ema = torch.zeros(2, C, dtype=torch.float64)
    W = [inp["W1"], inp["W2"]]; b = [inp["b1"], inp["b2"]]
    for s in range(2):
        step_in = O.make_inputs(B, D, C, seed=100 * s + B, n_data=N)
        f = [step_in["f1"], step_in["f2"]]
        ref = O.qmf_step(f, W, b, step_in["y"], step_in["idx"], hist, ema_x=ema, dtype=torch.float64)
        ema = ref["ema_x"]
        out = eng.step([x.cuda() for x in f], [x.cuda() for x in W], [x.cuda() for x in b], step_in["y"].cuda(),
                       idx=step_in["idx"].cuda())
        torch.cuda.synchronize()
        assert_close(out.loss, ref["loss"], TOL_TENSOR, "loss")
        assert_close(out.logits_df, ref["logits_df"], TOL_TENSOR, "zdf")
        for m in range(2):
            assert_close(out.logits[m], ref["logits"][m], TOL_TENSOR, "logits")
            assert_close(out.dweight[m], ref["dW"][m], TOL_TENSOR, "dW")
            assert_close(out.dbias[m], ref["db"][m], TOL_TENSOR, "db")
            assert_close(out.dfeat[m], ref["dfeat"][m], TOL_TENSOR, "dfeat")
        assert_close(eng.ema_x, ref["ema_x"], TOL_TENSOR, "ema")


@pytest.mark.parametrize("B,D,C,N,prec", [(300, 128, 32, 700, "tf32"), (257, 96, 33, 500, "tf32"), (515, 256, 48, 900, "bf16"),
                                          (260, 64, 64, 300, "tf32"), (333, 128, 80, 600, "bf16"), (270, 192, 96, 400, "tf32"),
                                          (1100, 128, 112, 2000, "tf32"), (520, 256, 128, 1500, "bf16"),
                                          (20000, 128, 101, 30000, "tf32"), (19, 64, 57, 40, "tf32")])
def test_fused_forward_kernel_all_widths_and_tilings(B, D, C, N, prec):
    """The fused QMF forward (lf_tc_fwd.cu) is instantiated per 16-column chunk count (C = 32..128 -> 2..8 chunks, two or
    four column shares per TMEM lane quarter) and sizes its M tiles from the batch (B = 20000 -> 296 tiles of 72 samples,
    half-empty lane quarters): every instantiation against the fp64 oracle, forward outputs AND the statistics that feed
    the History / EMA / loss, plus the integer accuracy counts of the fp32-rounded logits."""
    inp = O.make_inputs(B, D, C, seed=B + C, n_data=N)
    eng = _step(num_classes=C, mode="qmf", n_data=N, precision=prec)
    hist = O.HistoryState(N)
    ema = torch.zeros(2, C, dtype=torch.float64)
    W = [inp["W1"], inp["W2"]]; b = [inp["b1"], inp["b2"]]
    if prec == "bf16":
        W = [w.bfloat16().float() for w in W]
    for s in range(2):
        step_in = O.make_inputs(B, D, C, seed=100 * s + B, n_data=N)
        f = [step_in["f1"], step_in["f2"]]
        if prec == "bf16":
            f = [x.bfloat16().float() for x in f]
        ref = O.qmf_step(f, W, b, step_in["y"], step_in["idx"], hist, ema_x=ema, dtype=torch.float64)
        ema = ref["ema_x"]
        out = eng.step([x.cuda() for x in f], [x.cuda() for x in W], [x.cuda() for x in b], step_in["y"].cuda(),
                       idx=step_in["idx"].cuda())
        torch.cuda.synchronize()
        assert_close(out.loss, ref["loss"], TOL_TENSOR, "loss")
        assert_close(out.logits_df, ref["logits_df"], TOL_TENSOR, "zdf")
        assert_close(out.avg_logits, ref["avg_logits"], TOL_TENSOR, "avg")
        for m in range(2):
            assert_close(out.logits[m], ref["logits"][m], TOL_TENSOR, "logits")
            assert_close(out.dweight[m], ref["dW"][m], TOL_TENSOR, "dW")
            assert_close(out.dfeat[m].float(), ref["dfeat"][m], TOL_TENSOR, "dfeat")      # bf16 tensors in LF_PREC_BF16
        assert_close(eng.ema_x, ref["ema_x"], TOL_TENSOR, "ema")
        assert_close(eng.correctness, hist.correctness, 2e-3, "correctness")
        # counts are exact functions of the logits the kernel itself produced
        z1, z2, zdf, y = out.logits[0], out.logits[1], out.logits_df, step_in["y"].cuda()
        acc = out.accuracies()
        assert abs(acc["x1_acc_uncal"] - float((z1.argmax(1) == y).float().mean())) < 1e-6
        assert abs(acc["x2_acc_uncal"] - float((z2.argmax(1) == y).float().mean())) < 1e-6
        assert abs(acc["df_acc"] - float((zdf.argmax(1) == y).float().mean())) < 1e-6
        assert abs(acc["joint_acc"] - float((out.avg_logits.argmax(1) == y).float().mean())) < 1e-6


@pytest.mark.parametrize("B,D,C", [(2048, 768, 101), (777, 512, 309), (4096, 512, 309)])
def test_jlogits_tensor_pipe_matches_oracle(B, D, C):
    inp = O.make_inputs(B, D, C, seed=B + C)
    eng = _step(num_classes=C, mode="jlogits", precision="tf32")
    ref = O.jlogits_step([inp["f1"], inp["f2"]], [inp["W1"], inp["W2"]], [inp["b1"], inp["b2"]], inp["y"],
                         ema_x=torch.zeros(2, C, dtype=torch.float64), dtype=torch.float64)
    out = eng.step([inp["f1"].cuda(), inp["f2"].cuda()], [inp["W1"].cuda(), inp["W2"].cuda()],
                   [inp["b1"].cuda(), inp["b2"].cuda()], inp["y"].cuda(), ogm_alpha=0.8)
    torch.cuda.synchronize()
    assert_close(out.loss, ref["loss"], TOL_TENSOR, "loss")
    for m in range(2):
        assert_close(out.logits[m], ref["logits"][m], TOL_TENSOR, "logits")
        assert_close(out.dweight[m], ref["dW"][m], TOL_TENSOR, "dW")
        assert_close(out.dbias[m], ref["db"][m], TOL_TENSOR, "db")
        assert_close(out.dfeat[m], ref["dfeat"][m], TOL_TENSOR, "dfeat")
    k = O.ogm_coeffs(ref["score1"], ref["score2"], 0.8)
    assert_close(eng.coeff, np.array(k), 1e-2, "coeff")


def test_tensor_pipe_matches_reference_golden_wide():
    """Golden vectors from the unmodified reference (C = 101 QMF, C = 309 jlogits) through LF_PREC_TF32."""
    g = load_golden("qmf_food_c101")
    B, D, C, N, steps = [int(v) for v in g["meta"]]
    eng = _step(num_classes=C, mode="qmf", n_data=N, precision="tf32")
    for s in range(steps):
        p = f"s{s}_"
        out = eng.step([cu(g[p + "f1"]), cu(g[p + "f2"])], [cu(g["W1"]), cu(g["W2"])], [cu(g["b1"]), cu(g["b2"])],
                       cu(g[p + "y"]), idx=cu(g[p + "idx"]))
        torch.cuda.synchronize()
        assert_close(out.loss, g[p + "loss"], TOL_TENSOR, "loss")
        for m in range(2):
            assert_close(out.logits[m], g[p + f"z{m+1}"], TOL_TENSOR, "z")
            assert_close(out.dweight[m], g[p + f"dW{m+1}"], TOL_TENSOR, "dW")
            assert_close(out.dfeat[m], g[p + f"df{m+1}"], TOL_TENSOR, "df")
    g = load_golden("ogm_wide_c309")
    B, D, C, _, steps = [int(v) for v in g["meta"]]
    eng = _step(num_classes=C, mode="jlogits", precision="tf32")
    for s in range(steps):
        p = f"s{s}_"
        out = eng.step([cu(g[p + "f1"]), cu(g[p + "f2"])], [cu(g["W1"]), cu(g["W2"])], [cu(g["b1"]), cu(g["b2"])],
                       cu(g[p + "y"]))
        torch.cuda.synchronize()
        assert_close(out.loss, g[p + "loss"], TOL_TENSOR, "loss")
        for m in range(2):
            assert_close(out.dweight[m], g[p + f"dW{m+1}"], TOL_TENSOR, "dW")
            assert_close(out.dfeat[m], g[p + f"df{m+1}"], TOL_TENSOR, "df")


def test_step_is_deterministic_run_to_run():
    inp = O.make_inputs(3000, 512, 6, seed=3, n_data=4000)
    outs = []
    for _ in range(2):
        eng = _step(num_classes=6, mode="qmf", n_data=4000)
        o = eng.step([inp["f1"].cuda(), inp["f2"].cuda()], [inp["W1"].cuda(), inp["W2"].cuda()],
                     [inp["b1"].cuda(), inp["b2"].cuda()], inp["y"].cuda(), idx=inp["idx"].cuda())
        torch.cuda.synchronize()
        outs.append([o.loss.clone(), o.dweight[0].clone(), o.dfeat[1].clone(), eng.correctness.clone()])
    for a, b in zip(*outs):
        assert torch.equal(a, b), "fused step must be bit-reproducible (reference runs deterministic=True)"


def test_qmf_batch_of_one_is_rejected():
    """The reference raises for B == 1; the C-ABI returns LF_ERR_BAD_ARG."""
    from multimodal_clinical_b200._lib import LfError
    eng = _step(num_classes=3, mode="qmf", n_data=10)
    inp = O.make_inputs(1, 16, 3, seed=1, n_data=10)
    with pytest.raises(LfError):
        eng.step([inp["f1"].cuda(), inp["f2"].cuda()], [inp["W1"].cuda(), inp["W2"].cuda()],
                 [inp["b1"].cuda(), inp["b2"].cuda()], inp["y"].cuda(), idx=inp["idx"].cuda())


def test_qmf_degenerate_history_gives_nan_loss():
    """All N indices covered by one batch on step 1 -> max == min -> NaN margins."""
    N = 16
    inp = O.make_inputs(N, 32, 4, seed=2)
    eng = _step(num_classes=4, mode="qmf", n_data=N)
    out = eng.step([inp["f1"].cuda(), inp["f2"].cuda()], [inp["W1"].cuda(), inp["W2"].cuda()],
                   [inp["b1"].cuda(), inp["b2"].cuda()], inp["y"].cuda(), idx=torch.arange(N).cuda())
    assert torch.isnan(out.loss).item()


@pytest.mark.parametrize("mode", ["OGM", "OGM_GE", "noise"])
def test_modulate_scale_and_noise_moments(mode):
    """existing_algos/OGM_GE.py:42-54: exact scale when noise is off; with noise on, the residual
    (g' - k g)/sigma must look N(0,1) (mean, std, kurtosis) and non-4-D grads stay untouched."""
    torch.manual_seed(0)
    eng = _step(num_classes=6, mode="jlogits")
    eng.coeff.copy_(torch.tensor([0.3096, 1.0]))
    shapes = [(64, 3, 7, 7), (64, 64, 3, 3), (128, 64, 3, 3), (512, 512, 3, 3), (7, 5, 3, 1)]
    grads = [torch.randn(s, device="cuda") * (1e-3 * (i + 1)) + 1e-4 for i, s in enumerate(shapes)]
    bn = torch.randn(64, device="cuda")
    before = [g.clone() for g in grads]
    bn0 = bn.clone()
    eng.modulate(grads + [bn], which=0, modulation=mode, seed=1234, offset=0)
    torch.cuda.synchronize()
    assert torch.equal(bn, bn0)
    k = 1.0 if mode == "noise" else 0.3096
    for g, g0 in zip(grads, before):
        if mode == "OGM":
            assert torch.equal(g, g0 * torch.tensor(k, device="cuda")) or relerr(g, g0 * k) < 1e-7
            continue
        sigma = g0.std().item() + 1e-8
        z = ((g - g0 * k) / sigma).double().flatten()
        n = z.numel()
        tol = 6.0 / np.sqrt(n)
        assert abs(z.mean().item()) < tol, "noise mean"
        assert abs(z.std().item() - 1.0) < tol + 2e-3, "noise std vs std(g)+1e-8"
        if n > 10000:
            assert abs((z ** 4).mean().item() - 3.0) < 0.15, "noise kurtosis"
    if mode != "OGM":
        # Philox stream is addressable: same (seed, offset) -> same draws; different offset -> different
        g2 = [b.clone() for b in before]
        eng.modulate(g2, which=0, modulation=mode, seed=1234, offset=0)
        assert all(torch.equal(a, b) for a, b in zip(grads, g2))
        g3 = [b.clone() for b in before]
        eng.modulate(g3, which=0, modulation=mode, seed=1234, offset=7)
        assert not torch.equal(grads[1], g3[1])


def test_modulate_matches_reference_golden():
    g = load_golden("ogm_modulate_small")
    eng = _step(num_classes=6, mode="jlogits")
    z1, z2, y = t(g["z1"]), t(g["z2"]), t(g["y"])
    s1, s2 = O.ogm_scores(z1, z2, y)
    k = O.ogm_coeffs(float(s1), float(s2), float(g["alpha"]))
    eng.coeff.copy_(torch.tensor(k))
    names = sorted(n[len("before/"):] for n in g if n.startswith("before/"))
    for which, enc in enumerate(("x1_model", "x2_model")):
        sel = [n for n in names if n.startswith(enc)]
        grads = [cu(g["before/" + n]).clone() for n in sel]
        eng.modulate(grads, which=which, modulation="OGM", seed=0, offset=0)
        for n, gr in zip(sel, grads):
            assert_close(gr, g["after_OGM/" + n], 1e-6, n)


# ---------------------------------------------------------------- LF_PREC_BF16: the reference's bf16-mixed mode
@pytest.mark.parametrize("mode,B,D,C,N", [("qmf", 1000, 768, 101, 5000), ("jlogits", 515, 512, 309, None), ("qmf", 130, 256, 40, 300)])
def test_bf16_mode_matches_oracle_on_bf16_rounded_inputs(mode, B, D, C, N):
    """bf16 features / heads (rounded once, like autocast), fp32 accumulation and fp32 row math.  Checked against the
    oracle run in fp64 on the same rounded inputs; 2e-2 is BASELINE.json's bf16 tolerance (dz is stored in bf16)."""
    inp = O.make_inputs(B, D, C, seed=B + C, n_data=N)
    r16 = lambda x: x.bfloat16().float()
    f = [r16(inp["f1"]), r16(inp["f2"])]; W = [r16(inp["W1"]), r16(inp["W2"])]; b = [inp["b1"], inp["b2"]]
    eng = _step(num_classes=C, mode=mode, n_data=N, precision="bf16")
    hist = O.HistoryState(N) if N else None
    ema = torch.zeros(2, C, dtype=torch.float64)
    for s in range(2):
        if mode == "qmf":
            ref = O.qmf_step(f, W, b, inp["y"], inp["idx"], hist, ema_x=ema, dtype=torch.float64)
        else:
            ref = O.jlogits_step(f, W, b, inp["y"], ema_x=ema, dtype=torch.float64)
        ema = ref["ema_x"]
        out = eng.step([x.cuda().bfloat16() for x in f], [inp["W1"].cuda(), inp["W2"].cuda()], [x.cuda() for x in b], inp["y"].cuda(),
                       idx=inp["idx"].cuda() if N else None)
        torch.cuda.synchronize()
        assert out.dfeat[0].dtype == torch.bfloat16
        assert_close(out.loss, ref["loss"], TOL_TENSOR, "loss")
        assert_close(out.logits[0], ref["logits"][0], TOL_TENSOR, "z1")
        for m in range(2):
            assert_close(out.dweight[m], ref["dW"][m], TOL_TENSOR, f"dW{m+1}")
            assert_close(out.dbias[m], ref["db"][m], TOL_TENSOR, f"db{m+1}")
            assert_close(out.dfeat[m].float(), ref["dfeat"][m], TOL_TENSOR, f"df{m+1}")
        assert_close(eng.ema_x, ref["ema_x"], TOL_TENSOR, "ema_x")


def test_bf16_mode_rejects_narrow_heads():
    from multimodal_clinical_b200 import _lib
    inp = O.make_inputs(64, 512, 6, seed=1)
    eng = _step(num_classes=6, mode="jlogits", precision="bf16")
    with pytest.raises(_lib.LfError):
        eng.step([inp["f1"].cuda(), inp["f2"].cuda()], [inp["W1"].cuda(), inp["W2"].cuda()], [inp["b1"].cuda(), inp["b2"].cuda()], inp["y"].cuda())
