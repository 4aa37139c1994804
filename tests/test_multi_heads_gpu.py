"""lf_multi_heads_step (csrc/lf_multi.cu): mean fusion of M = 2..4 narrow heads with per-modality feature widths
 against fixtures from the reference's own three-modality / unequal-width FusionNets
(mustard/joint_model.py, avmnist/joint_model.py) and against the CPU oracle in fp64 on seeded inputs.  Exact tier: 1e-5."""
import pytest
import torch
import torch.nn as nn

from oracle import late_fusion as O
from tests.util import assert_close, load_golden, t, TOL_FP32

pytestmark = pytest.mark.gpu


def _eng(C):
    from multimodal_clinical_b200.multi import MultiHeadStep
    return MultiHeadStep(C, device="cuda:0")


def _check(out, ref, M, y, B):
    assert_close(out.loss, ref["loss"], TOL_FP32, "loss")
    assert_close(out.avg_logits, ref["avg_logits"], TOL_FP32, "avg")
    for m in range(M):
        assert_close(out.logits[m], ref["logits"][m], TOL_FP32, f"z{m}")
        assert_close(out.dweight[m], ref["dW"][m], TOL_FP32, f"dW{m}")
        assert_close(out.dbias[m], ref["db"][m], TOL_FP32, f"db{m}")
        if out.dfeat[m] is not None:
            assert_close(out.dfeat[m], ref["dfeat"][m], TOL_FP32, f"df{m}")
    # counts are exact functions of the logits the kernel returned
    st = out.stats.cpu()
    assert int(st[1]) == int((out.avg_logits.argmax(1) == y).sum())
    for m in range(M):
        assert int(st[2 + m]) == int((out.logits[m].argmax(1) == y).sum())


@pytest.mark.parametrize("name", ["multi_mustard_b24", "multi_avmnist_b40"])
def test_multi_heads_match_reference_golden(name):
    g = load_golden(name)
    B, C, M = [int(v) for v in g["meta"]]
    out = _eng(C).step([t(g[f"f{m}"]).cuda() for m in range(M)], [t(g[f"W{m}"]).cuda() for m in range(M)],
                       [t(g[f"b{m}"]).cuda() for m in range(M)], t(g["y"]).cuda())
    torch.cuda.synchronize()
    assert_close(out.loss, g["loss"], TOL_FP32, "loss")
    assert_close(out.avg_logits, g["avg"], TOL_FP32, "avg")
    for m in range(M):
        assert_close(out.logits[m], g[f"z{m}"], TOL_FP32, f"z{m}")
        assert_close(out.dweight[m], g[f"dW{m}"], TOL_FP32, f"dW{m}")
        assert_close(out.dbias[m], g[f"db{m}"], TOL_FP32, f"db{m}")
        assert_close(out.dfeat[m], g[f"df{m}"], TOL_FP32, f"df{m}")
    assert abs(out.accuracies()["joint_acc"] - float(g["acc_joint"])) < 1e-7


# (B, C, dims): three modalities of one width (mustard), unequal widths (avmnist), widths that are not multiples of 4
# (scalar tile loads), four modalities, the widest head the kernel takes, a batch of one, more tiles than CTAs
@pytest.mark.parametrize("B,C,dims", [(24, 2, (100, 100, 100)), (4096, 10, (48, 192)), (37, 5, (81, 371, 300)), (1, 3, (8, 12)),
                                      (1000, 32, (64, 128, 32, 96)), (100000, 10, (48, 192)), (70000, 2, (100, 100, 100)),
                                      (513, 20, (512, 256))])
def test_multi_heads_match_fp64_oracle(B, C, dims):
    M = len(dims)
    g = torch.Generator().manual_seed(B + C + M)
    f = [torch.randn(B, d, generator=g) for d in dims]
    W = [(torch.rand(C, d, generator=g) * 2 - 1) / d ** 0.5 for d in dims]
    b = [(torch.rand(C, generator=g) * 2 - 1) / d ** 0.5 for d in dims]
    y = torch.randint(0, C, (B,), generator=g)
    ref = O.mean_fusion_multi_step(f, W, b, y, dtype=torch.float64)
    out = _eng(C).step([x.cuda() for x in f], [x.cuda() for x in W], [x.cuda() for x in b], y.cuda())
    torch.cuda.synchronize()
    _check(out, ref, M, y.cuda(), B)


def test_multi_heads_without_feature_gradients_and_run_to_run_determinism():
    B, C, dims = 3000, 10, (48, 192)
    g = torch.Generator().manual_seed(9)
    f = [torch.randn(B, d, generator=g).cuda() for d in dims]
    W = [torch.randn(C, d, generator=g).cuda() / d ** 0.5 for d in dims]
    b = [torch.randn(C, generator=g).cuda() for d in dims]
    y = torch.randint(0, C, (B,), generator=g).cuda()
    eng = _eng(C)
    a = eng.step(f, W, b, y, need_dfeat=False)
    assert a.dfeat == [None, None]
    c = eng.step(f, W, b, y); d = eng.step(f, W, b, y)
    torch.cuda.synchronize()
    for m in range(2):
        assert torch.equal(c.dweight[m], d.dweight[m]) and torch.equal(c.dfeat[m], d.dfeat[m]) and torch.equal(a.dweight[m], c.dweight[m])
    assert torch.equal(c.loss, d.loss)


def test_fused_mean_fusion_heads_module_matches_torch_autograd():
    """The nn.Module face under autograd, with encoders in front (three modalities): gradients reach the encoders and the
    heads exactly as through torch's own Linear + CrossEntropyLoss."""
    from multimodal_clinical_b200.multi import FusedMeanFusionHeads
    torch.manual_seed(3)
    B, C, dims = 200, 2, (100, 100, 100)
    enc = [nn.Linear(16, d).cuda() for d in dims]
    heads = [nn.Linear(d, C).cuda() for d in dims]
    x = [torch.randn(B, 16, device="cuda") for _ in dims]
    y = torch.randint(0, C, (B,), device="cuda")
    fused = FusedMeanFusionHeads(C)

    def grads(use_fused):
        for mod in enc + heads:
            mod.zero_grad()
        f = [torch.relu(e(v)) for e, v in zip(enc, x)]
        if use_fused:
            *zs, avg, loss = fused(f, heads, y)
        else:
            zs = [h(v) for h, v in zip(heads, f)]
            avg = (zs[0] + zs[1] + zs[2]) / 3
            loss = nn.functional.cross_entropy(avg, y)
        (2.0 * loss).backward()
        return loss.detach(), avg.detach(), [p.grad.clone() for mod in enc + heads for p in mod.parameters()]
    l0, a0, g0 = grads(False)
    l1, a1, g1 = grads(True)
    assert_close(l1, l0, TOL_FP32, "loss"); assert_close(a1, a0, TOL_FP32, "avg")
    for u, v in zip(g1, g0):
        assert_close(u, v, 2e-5, "parameter gradient")


def test_multi_heads_reject_what_they_do_not_take():
    from multimodal_clinical_b200 import _lib
    eng = _eng(40)
    f = [torch.randn(8, 16, device="cuda") for _ in range(2)]
    W = [torch.randn(40, 16, device="cuda") for _ in range(2)]
    b = [torch.randn(40, device="cuda") for _ in range(2)]
    with pytest.raises(_lib.LfError):
        eng.step(f, W, b, torch.zeros(8, dtype=torch.int64, device="cuda"))        # classes > 32: the wide-head paths take those
    with pytest.raises(_lib.LfError):
        _eng(4).step([x.cpu() for x in f], W, b, torch.zeros(8, dtype=torch.int64))  # no CPU fallback


def _mustard_cfg(tmp_path, name, **over):
    import yaml
    import multimodal_clinical_b200 as pkg
    import os
    base = yaml.safe_load(open(os.path.join(os.path.dirname(pkg.__file__), name, name + ".yaml")))
    base.update(over)
    p = tmp_path / (name + ".yaml")
    p.write_text(yaml.safe_dump(base))
    return str(p)


def test_main_entry_point_trains_mustard_three_modalities(tmp_path):
    """`main.py --dir mustard --config <yaml>`: three LSTM encoders + the fused three-head step, fit / validate / test with
    the reference's logged keys (mustard/joint_model.py:137-138, 197-201, 285-289)."""
    from multimodal_clinical_b200 import main, _lib
    cfg = _mustard_cfg(tmp_path, "mustard", num_epochs=2, batch_size=16, max_seq_len=6, synthetic_samples=64)
    lib = _lib.load()
    lib.lf_profile_enable(1)
    try:
        tr = main.main(["--dir", "mustard", "--config", cfg])
    finally:
        prof = _lib.profile_report()
        lib.lf_profile_enable(0)
    assert "multi_heads_step" in prof and prof["multi_heads_step"][0] >= 8, prof
    got = {k: float(v) for k, v in tr.callback_metrics.items()}
    for k in ("train_loss", "train_acc", "val_loss", "val_acc", "x1_val_acc", "x2_val_acc", "x3_val_acc", "avg_test_loss", "avg_test_acc",
              "x1_test_acc", "x3_test_acc"):
        assert k in got and got[k] == got[k], (k, got)
    assert 0.0 <= got["val_acc"] <= 1.0


def test_main_entry_point_trains_avmnist_unequal_widths(tmp_path):
    from multimodal_clinical_b200 import main
    cfg = _mustard_cfg(tmp_path, "avmnist", num_epochs=1, batch_size=16, synthetic_samples=48)
    tr = main.main(["--dir", "avmnist", "--config", cfg])
    got = {k: float(v) for k, v in tr.callback_metrics.items()}
    for k in ("train_loss", "val_acc", "x1_val_acc", "x2_val_acc", "avg_test_acc", "x2_test_acc"):
        assert k in got and got[k] == got[k], (k, got)


def test_avmnist_module_step_matches_torch_replica():
    """One optimizer step of MultimodalAVMnistModel vs the same network with torch heads + CrossEntropyLoss (the reference's
    forward, avmnist/joint_model.py:117-138): same loss, same updated parameters."""
    import argparse, copy
    import torch.nn.functional as F
    from multimodal_clinical_b200.avmnist.joint_model import MultimodalAVMnistModel
    torch.manual_seed(11)
    args = argparse.Namespace(num_classes=10, learning_rate=1e-3)
    mod = MultimodalAVMnistModel(args).cuda()
    rep = copy.deepcopy(mod.model)
    x1 = torch.randn(24, 1, 28, 28, device="cuda"); x2 = torch.randn(24, 1, 112, 112, device="cuda")
    y = torch.randint(0, 10, (24,), device="cuda")
    loss = mod.training_step((x1, x2, y), 0)
    loss.backward()
    f1 = F.relu(rep.x1_model(x1)); f2 = F.relu(rep.x2_model(x2))
    avg = (rep.classifier_x1(f1) + rep.classifier_x2(f2)) / 2
    ref = F.cross_entropy(avg, y)
    ref.backward()
    assert_close(loss, ref, TOL_FP32, "loss")
    for (n, p), (_, q) in zip(mod.model.named_parameters(), rep.named_parameters()):
        assert_close(p.grad, q.grad, 5e-5, n)
