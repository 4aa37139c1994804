"""GPU parity of the round-2 additions against fixtures generated from the UNMODIFIED reference (tests/golden/make_golden.py):
  * the per-modality ensemble with OGM-GE (cremad/ensemble_model_noised.py, model_type ensemble_ogm_ge)
  * the epoch-end unimodal offset correction (utils/BaseModel.py:161-202) -- device kernel pair and the LightningModule hook
  * the Food101 module itself (food101/joint_model_qmf.py: MLP heads on SigLIP embeddings, QMF loss)
  * feature-side pooling (cremad/joint_model_qmf.py:48-55) against torch's adaptive_avg_pool2d/3d
  * SGD inside the fused step, wired through configure_optimizers, against the stock optimizer."""
import argparse

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import late_fusion as O
from tests.util import assert_close, cu, load_golden, t, TOL_FP32, TOL_TENSOR

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ensemble_ogm_b48", "ensemble_wide_c101"])
def test_ensemble_step_matches_reference_golden(name):
    from multimodal_clinical_b200.step import LateFusionStep
    g = load_golden(name)
    B, D, C, _, steps = [int(v) for v in g["meta"]]
    eng = LateFusionStep(C, mode="ensemble", device="cuda:0", precision="fp32")
    W = [cu(g["W1"]), cu(g["W2"])]
    b = [cu(g["b1"]), cu(g["b2"])]
    for s in range(steps):
        p = f"s{s}_"
        out = eng.step([cu(g[p + "f1"]), cu(g[p + "f2"])], W, b, cu(g[p + "y"]), ogm_alpha=float(g["alpha"]))
        torch.cuda.synchronize()
        from multimodal_clinical_b200._lib import STAT
        st = out.stats.cpu()
        assert abs(float(st[STAT["CE_X1"]]) / B - float(g[p + "loss_x1"])) < 1e-5 * max(1.0, float(g[p + "loss_x1"]))
        assert abs(float(st[STAT["CE_X2"]]) / B - float(g[p + "loss_x2"])) < 1e-5 * max(1.0, float(g[p + "loss_x2"]))
        assert_close(out.loss, g[p + "loss_x1"] + g[p + "loss_x2"], TOL_FP32, "x1_loss + x2_loss")
        for m in range(2):
            # the engine differentiates x1_loss + x2_loss; the reference's backward is of their mean (x 1/2)
            assert_close(out.logits[m], g[p + f"z{m + 1}"], TOL_FP32, "logits")
            assert_close(out.dweight[m] * 0.5, g[p + f"dW{m + 1}"], TOL_FP32, "dW")
            assert_close(out.dbias[m] * 0.5, g[p + f"db{m + 1}"], TOL_FP32, "db")
            assert_close(out.dfeat[m] * 0.5, g[p + f"df{m + 1}"], TOL_FP32, "dfeat")
        assert_close(eng.coeff, g[p + "coeff"], 2e-5, "OGM-GE coefficients")
        assert abs(float(st[STAT["CNT_X1"]]) / B - float(g[p + "acc_x1"])) < 1e-6
        assert abs(float(st[STAT["CNT_JOINT"]]) / B - float(g[p + "acc_joint"])) < 1e-6


def test_ensemble_lightning_module_trains_like_the_reference():
    """model_type ensemble_ogm_ge through the factory: (x1_logits, x2_logits, x1_loss, x2_loss), manual optimisation,
    gradients of (x1_loss + x2_loss) / 2 on heads and features equal the reference fixture."""
    from multimodal_clinical_b200.cremad import get_model
    import multimodal_clinical_b200.cremad.ensemble_model_noised as me
    g = load_golden("ensemble_ogm_b48")
    B, D, C, _, steps = [int(v) for v in g["meta"]]

    class Enc(nn.Module):                                   # identity encoder that still owns a 4-D parameter for ogm_ge
        def __init__(self):
            super().__init__()
            self.conv = nn.Sequential(nn.Conv2d(D, D, 1, bias=False))
            with torch.no_grad():
                nn.init.dirac_(self.conv[0].weight)

        def forward(self, x):
            return self.conv(x)
    orig = me.resnet18
    me.resnet18 = lambda modality: Enc()
    tf32_was = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                 # the stub encoder is a (Dirac) convolution: keep it exact
    try:
        args = argparse.Namespace(model_type="ensemble_ogm_ge", num_classes=C, learning_rate=0.0, use_scheduler=False, grad_mod_type="OGM",
                                  alpha=float(g["alpha"]), fused_head_sgd=False)
        model = get_model(args).cuda()
    finally:
        me.resnet18 = orig
    net = model.model
    with torch.no_grad():
        net.x1_classifier.weight.copy_(cu(g["W1"])); net.x1_classifier.bias.copy_(cu(g["b1"]))
        net.x2_classifier.weight.copy_(cu(g["W2"])); net.x2_classifier.bias.copy_(cu(g["b2"]))
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    model._opt = opt
    model.optimizers = lambda: opt
    model.train()
    p = "s0_"
    a = cu(g[p + "f1"]).view(B, D, 1, 1).requires_grad_(True)
    v = cu(g[p + "f2"]).view(B, D, 1, 1).requires_grad_(True)
    z1, z2, l1, l2 = net(a, v, cu(g[p + "y"]))
    assert_close(l1, g[p + "loss_x1"], TOL_FP32, "x1_loss"); assert_close(l2, g[p + "loss_x2"], TOL_FP32, "x2_loss")
    ((l1 + l2) / 2).backward()
    assert_close(net.x1_classifier.weight.grad, g[p + "dW1"], TOL_FP32, "dW1")
    assert_close(net.x2_classifier.bias.grad, g[p + "db2"], TOL_FP32, "db2")
    assert_close(a.grad.view(B, D), g[p + "df1"], TOL_FP32, "df1")
    # and one full training_step (zero_grad -> backward -> ogm_ge -> step) runs and scales the encoder gradient
    loss = model.training_step((a.detach(), v.detach(), cu(g[p + "y"])), 0)
    torch.cuda.synchronize()
    assert_close(loss, (g[p + "loss_x1"] + g[p + "loss_x2"]) / 2, TOL_FP32, "avg_loss")
    torch.backends.cudnn.allow_tf32 = tf32_was


@pytest.mark.parametrize("name", ["epoch_end_c6", "epoch_end_c101"])
def test_epoch_offset_correction_kernel_matches_reference_golden(name):
    from multimodal_clinical_b200.utils.BaseModel import epoch_offset_correction
    g = load_golden(name)
    nb = int(g["meta"][0])
    logits = torch.cat([t(g[f"b{i}_logits"]) for i in range(nb)])
    labels = torch.cat([t(g[f"b{i}_labels"]) for i in range(nb)])
    offset, acc = epoch_offset_correction(logits.cuda(), labels.cuda())
    torch.cuda.synchronize()
    ref = O.epoch_offset_correction(logits, labels)
    assert_close(offset, ref["offset"], 1e-5, "offset")
    want = [float(g["log/val_epoch/val_avg_x1_acc_uncal"]), float(g["log/val_epoch/val_avg_x2_acc_uncal"]),
            float(g["log/val_epoch/val_avg_x1_acc"]), float(g["log/val_epoch/val_avg_x2_acc"])]
    assert np.allclose(acc.cpu().numpy(), want, atol=1e-6), (acc, want)


def test_lightning_epoch_end_hook_logs_the_reference_values():
    """JointLogitsBaseModel.on_validation_epoch_end fed with the fixture's per-batch logits logs what the reference logs."""
    from multimodal_clinical_b200.utils.BaseModel import JointLogitsBaseModel
    from multimodal_clinical_b200.heads import FusedLateFusionHead
    g = load_golden("epoch_end_c6")
    nb, B, C = [int(v) for v in g["meta"]]

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.fused = FusedLateFusionHead(C, mode="jlogits")

    class Model(JointLogitsBaseModel):
        def _build_model(self):
            return Net()

    m = Model(argparse.Namespace(num_classes=C, learning_rate=0.1, use_scheduler=False)).cuda()
    logged = {}
    m.log = lambda key, val, **kw: logged.__setitem__(key, float(val))
    for i in range(nb):
        m.val_metrics["val_logits"].append(cu(g[f"b{i}_logits"])); m.val_metrics["val_labels"].append(cu(g[f"b{i}_labels"]))
        m.val_metrics["val_loss"].append(cu(g[f"b{i}_loss"])); m.val_metrics["val_acc"].append(cu(g[f"b{i}_acc"]))
    m.on_validation_epoch_end()
    for k, v in g.items():
        if k.startswith("log/"):
            assert abs(logged[k[4:]] - float(v)) < 1e-6, (k, logged[k[4:]], float(v))
    assert all(len(v) == 0 for v in m.val_metrics.values())


def _fill_food101(net, seed):
    """tests/golden/make_golden.py::food101_fill, same draw order."""
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in sorted(net.named_parameters()):
            if not (n.startswith("x1_model.") or n.startswith("x2_model.")):
                continue
            bound = 1.0 / np.sqrt(p.shape[-1] if p.dim() > 1 else {0: 768, 3: 512, 6: 512}[int(n.split(".")[2])])
            p.copy_(((torch.rand(p.shape, generator=gen) * 2 - 1) * bound).to(p.device))


def test_food101_module_matches_reference_golden():
    """food101/joint_model_qmf.FusionNet of the reference vs ours (same state-dict names; the MLP's hidden layers on the fused
    hidden-layer kernels, the last Linear + QMF loss in the fused step), everything in the exact tier: outputs, gradients on
    every MLP parameter and on the embeddings.  The logits come out of THREE stacked 3xTF32 GEMMs (768 -> 512 -> 512 -> 101),
    each of which truncates its TMEM accumulator toward zero (~3e-6 at K = 768): 2e-5 for the stack, 1e-5 per layer
    (tests/test_hidden_gpu.py, tests/test_parity_gpu.py)."""
    TOL_STACK = 2e-5
    from multimodal_clinical_b200.food101.joint_model_qmf import FusionNet
    g = load_golden("food101_module_b32")
    B, D, C, N, steps, seed = [int(v) for v in g["meta"]]
    args = argparse.Namespace(num_classes=C, num_samples=N, encoder="precomputed", head_precision="fp32")
    net = FusionNet(args, nn.CrossEntropyLoss()).cuda()
    net.eval()                                   # Dropout off, as in the fixture; the fused head still trains (grad enabled)
    net.fused.train()
    _fill_food101(net, seed)
    names = dict(net.named_parameters())
    for s in range(steps):
        p = f"s{s}_"
        e1 = cu(g[p + "e1"]).requires_grad_(True); e2 = cu(g[p + "e2"]).requires_grad_(True)
        net.zero_grad()
        z1, z2, avg, loss, zdf = net(e1, e2, cu(g[p + "y"]), cu(g[p + "idx"]))
        loss.backward()
        torch.cuda.synchronize()
        assert_close(z1, g[p + "z1"], TOL_STACK, "z1"); assert_close(z2, g[p + "z2"], TOL_STACK, "z2")
        assert_close(avg, g[p + "avg"], TOL_STACK, "avg"); assert_close(zdf, g[p + "zdf"], TOL_STACK, "zdf")
        assert_close(loss, g[p + "loss"], TOL_FP32, "loss")
        assert_close(e1.grad, g[p + "de1"], 4e-5, "d embeddings 1"); assert_close(e2.grad, g[p + "de2"], 4e-5, "d embeddings 2")
        for k, ref in g.items():
            if k.startswith(p + "grad/"):
                assert_close(names[k[len(p) + 5:]].grad, ref, 4e-5, k)
            elif k.startswith(p + "gradR/"):
                gr = torch.Generator().manual_seed(1000 + s)
                gp = names[k[len(p) + 6:]].grad.cpu()
                R = torch.randn(gp.shape[1], 4, generator=gr)
                L = torch.randn(4, gp.shape[0], generator=gr)
                assert_close(gp @ R, ref, 4e-5, k)
                assert_close(L @ gp, g[k.replace("gradR/", "gradL/")], 4e-5, k + " (left)")
        assert_close(net.qmf.history[0].correctness, g[p + "corr"][0], 1e-6, "history")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,C,H,W", [(5, 3, 512, 7, 7), (3, 1, 130, 9, 10), (2, 4, 64, 1, 1)])
def test_pool_features_matches_torch(dtype, B, T, C, H, W):
    from multimodal_clinical_b200.cremad._pool import pool_features
    gen = torch.Generator().manual_seed(3)
    a = torch.randn(B, C, H, W, generator=gen).to(dtype).cuda().requires_grad_(True)
    v = torch.randn(B * T, C, H, W, generator=gen).to(dtype).cuda().requires_grad_(True)
    pa, pv = pool_features(a, v)
    (pa.float().square().sum() + (pv.float() * 3).sum()).backward()
    a2 = a.detach().float().requires_grad_(True); v2 = v.detach().float().requires_grad_(True)
    ra = torch.flatten(F.adaptive_avg_pool2d(a2, 1), 1)
    rv = torch.flatten(F.adaptive_avg_pool3d(v2.view(B, -1, C, H, W).permute(0, 2, 1, 3, 4), 1), 1)
    (ra.square().sum() + (rv * 3).sum()).backward()
    tol = 1e-6 if dtype == torch.float32 else 8e-3
    assert pa.dtype == dtype and pa.shape == (B, C) and pv.shape == (B, C)
    assert_close(pa.float(), ra, tol, "audio pool"); assert_close(pv.float(), rv, tol, "visual pool")
    assert_close(a.grad.float(), a2.grad, tol * 2, "d audio maps"); assert_close(v.grad.float(), v2.grad, tol * 2, "d visual maps")


def test_in_step_sgd_through_configure_optimizers_matches_the_stock_optimizer():
    """Food101 QMF module under bf16 autocast: heads updated inside the fused step (SGDWithFusedHeads) vs torch.optim.SGD on every
    parameter -- same weights after a few steps, StepLR followed, momentum buffers visible in the optimizer state."""
    from multimodal_clinical_b200.food101.joint_model_qmf import MultimodalFoodModel
    from multimodal_clinical_b200.utils.fused_sgd import SGDWithFusedHeads
    C, N, B = 101, 4096, 512
    results = {}
    for fused in (True, False):
        torch.manual_seed(11)
        args = argparse.Namespace(num_classes=C, num_samples=N, encoder="precomputed", learning_rate=0.05, use_scheduler=True,
                                  fused_head_sgd=fused)
        model = MultimodalFoodModel(args).cuda()
        for mod in model.modules():
            if isinstance(mod, nn.Dropout):
                mod.p = 0.0
        (opt,), (sched,) = model.configure_optimizers()
        assert isinstance(opt, SGDWithFusedHeads) == fused
        sched = sched["scheduler"]
        sched.step_size = 2                                  # StepLR(50, 0.5) -> every 2 "epochs" for the test
        gen = torch.Generator().manual_seed(5)
        model.train()
        for step in range(5):
            e1 = torch.randn(B, 768, generator=gen).cuda(); e2 = torch.randn(B, 768, generator=gen).cuda()
            y = torch.randint(0, C, (B,), generator=gen).cuda()
            idx = ((torch.arange(B) + step * B) % N).cuda()
            opt.zero_grad()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                loss = model.training_step((e1, e2, y, idx), step)
            loss.backward()
            opt.step()
            sched.step()
        torch.cuda.synchronize()
        if fused:
            assert model.model.fused.in_step_updated(), "the fused step did not take the head update"
            assert "momentum_buffer" in opt.state[model.model.x1_model.mlp[6].weight]
        results[fused] = {n: p.detach().float().clone() for n, p in model.named_parameters()}
    for n in results[True]:
        assert_close(results[True][n], results[False][n], 2e-4, n)


def test_step_mid_loops_when_the_batch_exceeds_its_grid():
    """A QMF batch larger than step_mid's grid (148 x 512 threads): the rolled loops for tickets, History.confidence winners
    and the pair terms beyond the register-held first position, with heavy duplication of indices (3000 distinct << B).
    The History is longer than the touched range: the reference adds the batch-MEAN correctness to every touched entry, so
    a first step that touches all of it leaves a uniform History and a NaN normalisation (covered by
    test_qmf_degenerate_history_gives_nan_loss); the untouched tail keeps min != max here."""
    from multimodal_clinical_b200.step import LateFusionStep
    B, D, C, N, touched = 80000, 32, 6, 5000, 3000
    eng = LateFusionStep(C, mode="qmf", n_data=N, device="cuda:0")
    hist = O.HistoryState(N)
    ema = torch.zeros(2, C, dtype=torch.float64)
    base = O.make_inputs(8, D, C, seed=5)
    W = [base["W1"], base["W2"]]; b = [base["b1"], base["b2"]]
    gen = torch.Generator().manual_seed(77)
    for s in range(2):
        f = [torch.randn(B, D, generator=gen), torch.randn(B, D, generator=gen)]
        y = torch.randint(0, C, (B,), generator=gen)
        idx = torch.randint(0, touched, (B,), generator=gen)
        ref = O.qmf_step(f, W, b, y, idx, hist, ema_x=ema, dtype=torch.float64)
        ema = ref["ema_x"]
        out = eng.step([x.cuda() for x in f], [x.cuda() for x in W], [x.cuda() for x in b], y.cuda(), idx=idx.cuda())
        torch.cuda.synchronize()
        assert torch.isfinite(ref["loss"]).item(), "degenerate test input"
        assert_close(out.loss, ref["loss"], TOL_FP32, f"loss step {s}")
        assert_close(out.dweight[0], ref["dW"][0], TOL_FP32, "dW1"); assert_close(out.dfeat[1], ref["dfeat"][1], TOL_FP32, "df2")
        assert_close(eng.correctness, hist.correctness, 1e-6, "history.correctness")
        assert_close(eng.confidence, hist.confidence, 1e-6, "history.confidence")
