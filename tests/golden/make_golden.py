"""Generate golden input/output vectors by running the UNMODIFIED reference modules on the CPU.

Run in the build container only (``python tests/golden/make_golden.py``): it needs /root/reference,
which does not exist on the GPU box.  The resulting ``tests/golden/*.npz`` files are committed and are
what the tests read.  Recipe: SURVEY.md Appendix B — a test-only ``pytorch_lightning`` shim (the
package is not installed), identity encoders, ``Tensor.cuda`` patched to identity because
existing_algos/QMF.py:63,66 hard-codes ``.cuda()``.
"""
import argparse
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("LF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def install_shims():
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        automatic_optimization = True

        def log(self, *a, **k):
            pass

        def optimizers(self):
            return self._opt

        def lr_schedulers(self):
            return []

        def manual_backward(self, loss):
            loss.backward()

    pl.LightningModule = LightningModule
    pl.seed_everything = lambda s, workers=False: torch.manual_seed(s)
    pl.loggers = types.ModuleType("pytorch_lightning.loggers")
    pl.loggers.WandbLogger = object
    pl.callbacks = types.SimpleNamespace(LearningRateMonitor=object, ModelCheckpoint=object)
    sys.modules.update({"pytorch_lightning": pl, "pytorch_lightning.loggers": pl.loggers})
    if "sympy" not in sys.modules:
        try:
            import sympy  # noqa: F401  (stray import at cremad/joint_model_qmf.py:1)
        except Exception:
            sy = types.ModuleType("sympy")
            sy.Idx = object
            sys.modules["sympy"] = sy
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self      # CPU run of QMF.py:63,66


def lin_init(C, D, g):
    bound = 1.0 / np.sqrt(D)
    W = (torch.rand(C, D, generator=g) * 2 - 1) * bound
    b = (torch.rand(C, generator=g) * 2 - 1) * bound
    return W, b


def set_heads(net, W1, b1, W2, b2):
    C, D = W1.shape
    net.x1_classifier = nn.Linear(D, C)
    net.x2_classifier = nn.Linear(D, C)
    with torch.no_grad():
        net.x1_classifier.weight.copy_(W1); net.x1_classifier.bias.copy_(b1)
        net.x2_classifier.weight.copy_(W2); net.x2_classifier.bias.copy_(b2)


def accs(z1, z2, avg, y, off, zdf=None):
    f = lambda z: float(torch.mean((torch.argmax(z, dim=1) == y).float()))
    out = {"acc_x1_uncal": f(z1), "acc_x2_uncal": f(z2), "acc_x1_cal": f(z1 + off[0]),
           "acc_x2_cal": f(z2 + off[1]), "acc_joint": f(avg)}
    if zdf is not None:
        out["acc_df"] = f(zdf)
    return out


def golden_qmf(name, B, D, C, N, steps, seed, idx_mode="replacement", dtype=torch.float32,
               module="cremad.joint_model_qmf"):
    """``module``.FusionNet (identity encoders; cremad/joint_model_qmf.py or one of its loss ablations
    cremad/joint_model_qmf_ablate_L{joint,unimodal}.py / cremad/joint_model_ogm_ge_lreg.py) + utils/EMA.EMA for
    ``steps`` steps with fresh features each step, fixed weights; records every step's outputs and the History."""
    import importlib
    mq = importlib.import_module(module)
    from utils.EMA import EMA
    mq.resnet18 = lambda modality: nn.Identity()
    g = torch.Generator().manual_seed(seed)
    W1, b1 = lin_init(C, D, g)
    W2, b2 = lin_init(C, D, g)
    net = mq.FusionNet(argparse.Namespace(num_classes=C, num_samples=N), nn.CrossEntropyLoss())
    set_heads(net, W1, b1, W2, b2)
    net = net.to(dtype)
    ema = EMA(torch.zeros(2, C, dtype=dtype))
    rec = {"W1": W1.numpy(), "b1": b1.numpy(), "W2": W2.numpy(), "b2": b2.numpy(),
           "meta": np.array([B, D, C, N, steps], dtype=np.int64)}
    for s in range(steps):
        f1 = torch.randn(B, D, generator=g); f2 = torch.randn(B, D, generator=g)
        y = torch.randint(0, C, (B,), generator=g, dtype=torch.int64)
        if idx_mode == "replacement":
            idx = torch.randint(0, N, (B,), generator=g, dtype=torch.int64)
        else:
            idx = (torch.arange(B, dtype=torch.int64) + s * B) % N
        a = f1.to(dtype).view(B, D, 1, 1).requires_grad_(True)
        v = f2.to(dtype).view(B, D, 1, 1).requires_grad_(True)
        net.zero_grad()
        z1, z2, avg, loss, zdf = net(a, v, y, idx)
        loss.backward()
        ema.update(torch.mean(torch.stack([z1, z2]), dim=1))
        off = ema.offset
        r = {"f1": f1, "f2": f2, "y": y, "idx": idx, "z1": z1, "z2": z2, "avg": avg, "zdf": zdf,
             "loss": loss, "dW1": net.x1_classifier.weight.grad, "db1": net.x1_classifier.bias.grad,
             "dW2": net.x2_classifier.weight.grad, "db2": net.x2_classifier.bias.grad,
             "df1": a.grad.view(B, D), "df2": v.grad.view(B, D), "ema_x": ema.x, "ema_off": off}
        for k, t in r.items():
            rec[f"s{s}_{k}"] = t.detach().to(torch.float64 if dtype == torch.float64 and t.is_floating_point()
                                             else t.dtype).numpy().copy()
        for k, val in accs(z1.detach(), z2.detach(), avg.detach(), y, off, zdf.detach()).items():
            rec[f"s{s}_{k}"] = np.float64(val)
        rec[f"s{s}_corr"] = np.stack([h.correctness for h in net.qmf.history]).copy()
        rec[f"s{s}_confid"] = np.stack([h.confidence for h in net.qmf.history]).copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name, "loss(last) =", float(loss.detach()))


def golden_ogm(name, B, D, C, steps, seed, alpha, enrico=False):
    """cremad/joint_model_ogm_ge.FusionNet (or enrico/joint_model.FusionNet) + EMA + existing_algos.OGM_GE.ogm_ge
    with modulation='OGM' on a stub encoder holding one Dirac 1x1 conv, so the coefficients can be read
    back as grad_after/grad_before."""
    from existing_algos.OGM_GE import ogm_ge
    from utils.EMA import EMA
    g = torch.Generator().manual_seed(seed)
    W1, b1 = lin_init(C, D, g)
    W2, b2 = lin_init(C, D, g)
    if enrico:
        import torchvision
        import enrico.joint_model as ej
        orig = torchvision.models.resnet18
        ej.tmodels.resnet18 = lambda pretrained=True: orig(weights=None)
        net = ej.FusionNet(C, nn.CrossEntropyLoss())
        for xm, (W, b) in zip((net.x1_model, net.x2_model), ((W1, b1), (W2, b2))):
            xm.model = nn.Identity()
            xm.classifier = nn.Linear(D, C)
            with torch.no_grad():
                xm.classifier.weight.copy_(W); xm.classifier.bias.copy_(b)
        heads = [net.x1_model.classifier, net.x2_model.classifier]
    else:
        import cremad.joint_model_ogm_ge as mo

        def stub(modality):
            # ogm_ge indexes str(name).split('.')[1] (existing_algos/OGM_GE.py:47): the parameter
            # name needs a dot, so the conv sits inside a Sequential ("0.weight")
            conv = nn.Conv2d(D, D, 1, bias=False)
            with torch.no_grad():
                nn.init.dirac_(conv.weight)
            return nn.Sequential(conv)
        mo.resnet18 = stub
        net = mo.FusionNet(C, nn.CrossEntropyLoss())
        set_heads(net, W1, b1, W2, b2)
        heads = [net.x1_classifier, net.x2_classifier]
    ema = EMA(torch.zeros(2, C))
    rec = {"W1": W1.numpy(), "b1": b1.numpy(), "W2": W2.numpy(), "b2": b2.numpy(),
           "meta": np.array([B, D, C, 0, steps], dtype=np.int64), "alpha": np.float64(alpha)}
    for s in range(steps):
        # a label-aligned signal in one modality per step makes both coefficient branches occur
        f1 = torch.randn(B, D, generator=g); f2 = torch.randn(B, D, generator=g)
        y = torch.randint(0, C, (B,), generator=g, dtype=torch.int64)
        if s % 2 == 0:
            f1 = f1 + 0.3 * W1[y] * np.sqrt(D)      # modality 1 informative -> score1 > score2
        else:
            f2 = f2 + 0.3 * W2[y] * np.sqrt(D)      # modality 2 informative -> score2 > score1
        a = f1.clone().view(B, D, 1, 1).requires_grad_(True)
        v = f2.clone().view(B, D, 1, 1).requires_grad_(True)
        net.zero_grad()
        z1, z2, avg, loss = net(a, v, y)
        loss.backward()
        ema.update(torch.mean(torch.stack([z1, z2]), dim=1))
        off = ema.offset
        r = {"f1": f1, "f2": f2, "y": y, "z1": z1, "z2": z2, "avg": avg, "loss": loss,
             "dW1": heads[0].weight.grad, "db1": heads[0].bias.grad,
             "dW2": heads[1].weight.grad, "db2": heads[1].bias.grad,
             "ema_x": ema.x, "ema_off": off}
        if not enrico:
            r["df1"] = a.grad.view(B, D); r["df2"] = v.grad.view(B, D)
            g1 = net.x1_model[0].weight.grad.clone(); g2 = net.x2_model[0].weight.grad.clone()
            ogm_ge(net, z1, z2, y, alpha=alpha, modulation="OGM")
            k1 = (net.x1_model[0].weight.grad.flatten() @ g1.flatten()) / (g1.flatten() @ g1.flatten())
            k2 = (net.x2_model[0].weight.grad.flatten() @ g2.flatten()) / (g2.flatten() @ g2.flatten())
            rec[f"s{s}_coeff"] = np.array([float(k1), float(k2)])
        for k, t in r.items():
            rec[f"s{s}_{k}"] = t.detach().numpy().copy()
        for k, val in accs(z1.detach(), z2.detach(), avg.detach(), y, off).items():
            rec[f"s{s}_{k}"] = np.float64(val)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name, "loss(last) =", float(loss.detach()))


def golden_modulate(name, seed):
    """existing_algos.OGM_GE.ogm_ge on a tiny encoder with 4-D and non-4-D parameters: which tensors are
    touched, the scale, and (mode 'noise'/'OGM_GE') the noise moments relative to std(g)+1e-8."""
    from existing_algos.OGM_GE import ogm_ge
    g = torch.Generator().manual_seed(seed)

    class Enc(nn.Module):
        def __init__(self):
            super().__init__()
            self.stem = nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8))
            self.layer1 = nn.Sequential(nn.Conv2d(8, 16, 3, bias=False), nn.BatchNorm2d(16))

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.x1_model = Enc(); self.x2_model = Enc()

    net = Net()
    grads = {}
    for n, p in net.named_parameters():
        p.grad = torch.randn(p.shape, generator=g) * 1e-3
        grads[n] = p.grad.clone()
    B, C = 48, 6
    z1 = torch.randn(B, C, generator=g) * 2; z2 = torch.randn(B, C, generator=g)
    y = torch.randint(0, C, (B,), generator=g)
    ogm_ge(net, z1, z2, y, alpha=0.8, modulation="OGM")
    rec = {"z1": z1.numpy(), "z2": z2.numpy(), "y": y.numpy(), "alpha": np.float64(0.8)}
    for n, p in net.named_parameters():
        rec["before/" + n] = grads[n].numpy()
        rec["after_OGM/" + n] = p.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name)


def golden_ensemble(name, B, D, C, steps, seed, alpha):
    """cremad/ensemble_model_noised.FusionNet (one CE per modality) with stub encoders + existing_algos.OGM_GE.ogm_ge
    (modulation 'OGM'): losses, gradients of (x1_loss + x2_loss) / 2 and the coefficients read back from the grads."""
    from existing_algos.OGM_GE import ogm_ge
    import cremad.ensemble_model_noised as me
    g = torch.Generator().manual_seed(seed)
    W1, b1 = lin_init(C, D, g)
    W2, b2 = lin_init(C, D, g)

    def stub(modality):
        conv = nn.Conv2d(D, D, 1, bias=False)
        with torch.no_grad():
            nn.init.dirac_(conv.weight)
        return nn.Sequential(conv)
    me.resnet18 = stub
    net = me.FusionNet(C, nn.CrossEntropyLoss())
    set_heads(net, W1, b1, W2, b2)
    rec = {"W1": W1.numpy(), "b1": b1.numpy(), "W2": W2.numpy(), "b2": b2.numpy(),
           "meta": np.array([B, D, C, 0, steps], dtype=np.int64), "alpha": np.float64(alpha)}
    for s in range(steps):
        f1 = torch.randn(B, D, generator=g); f2 = torch.randn(B, D, generator=g)
        y = torch.randint(0, C, (B,), generator=g, dtype=torch.int64)
        if s % 2 == 0:
            f1 = f1 + 0.3 * W1[y] * np.sqrt(D)
        else:
            f2 = f2 + 0.3 * W2[y] * np.sqrt(D)
        a = f1.clone().view(B, D, 1, 1).requires_grad_(True)
        v = f2.clone().view(B, D, 1, 1).requires_grad_(True)
        net.zero_grad()
        z1, z2, l1, l2 = net(a, v, y)
        ((l1 + l2) / 2).backward()                                       # ensemble_model_noised.py:103, 119-120
        g1 = net.x1_model[0].weight.grad.clone(); g2 = net.x2_model[0].weight.grad.clone()
        ogm_ge(net, z1, z2, y, alpha=alpha, modulation="OGM")
        k1 = (net.x1_model[0].weight.grad.flatten() @ g1.flatten()) / (g1.flatten() @ g1.flatten())
        k2 = (net.x2_model[0].weight.grad.flatten() @ g2.flatten()) / (g2.flatten() @ g2.flatten())
        avg = (z1 + z2) / 2
        r = {"f1": f1, "f2": f2, "y": y, "z1": z1, "z2": z2, "loss_x1": l1, "loss_x2": l2,
             "dW1": net.x1_classifier.weight.grad, "db1": net.x1_classifier.bias.grad,
             "dW2": net.x2_classifier.weight.grad, "db2": net.x2_classifier.bias.grad,
             "df1": a.grad.view(B, D), "df2": v.grad.view(B, D)}
        for k, t in r.items():
            rec[f"s{s}_{k}"] = t.detach().numpy().copy()
        rec[f"s{s}_coeff"] = np.array([float(k1), float(k2)])
        f = lambda z: float(torch.mean((torch.argmax(z, dim=1) == y).float()))
        rec[f"s{s}_acc_x1"] = np.float64(f(z1.detach())); rec[f"s{s}_acc_x2"] = np.float64(f(z2.detach()))
        rec[f"s{s}_acc_joint"] = np.float64(f(avg.detach()))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name, "losses(last) =", float(l1.detach()), float(l2.detach()))


def golden_epoch_end(name, n_batches, B, C, seed):
    """utils/BaseModel.JointLogitsBaseModel.on_validation_epoch_end (:161-202) of the UNMODIFIED reference class: the
    collected (N, M, C) logits, the labels and every value it logs."""
    import utils.BaseModel as RB

    class Model(RB.JointLogitsBaseModel):
        def _build_model(self):
            return nn.Identity()

    m = Model(argparse.Namespace(num_classes=C, learning_rate=0.1, use_scheduler=False))
    logged = {}
    m.log = lambda key, val, **kw: logged.__setitem__(key, float(val))
    g = torch.Generator().manual_seed(seed)
    rec = {"meta": np.array([n_batches, B, C], dtype=np.int64)}
    shift = torch.randn(2, C, generator=g) * 0.7              # per-modality class bias: what the correction removes
    for i in range(n_batches):
        nb = B if i < n_batches - 1 else B // 2 + 1            # short last batch
        y = torch.randint(0, C, (nb,), generator=g, dtype=torch.int64)
        z = torch.randn(nb, 2, C, generator=g) + shift
        z[torch.arange(nb), :, y] += 1.0                       # some signal
        m.val_metrics["val_logits"].append(z.clone())
        m.val_metrics["val_labels"].append(y.clone())
        m.val_metrics["val_loss"].append(torch.rand((), generator=g))
        m.val_metrics["val_acc"].append(torch.rand((), generator=g))
        rec[f"b{i}_logits"] = z.numpy().copy(); rec[f"b{i}_labels"] = y.numpy().copy()
        rec[f"b{i}_loss"] = m.val_metrics["val_loss"][-1].numpy().copy(); rec[f"b{i}_acc"] = m.val_metrics["val_acc"][-1].numpy().copy()
    m.on_validation_epoch_end()
    for k, v in logged.items():
        rec["log/" + k] = np.float64(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name, logged)


def golden_food101(name, B, C, N, steps, seed):
    """food101/joint_model_qmf.FusionNet of the reference (:28-81): SigLIP stubbed by a module that returns the given
    768-d embeddings, the two 768 -> 512 -> 512 -> C MLP heads in eval() mode (Dropout off), the
    QMF loss; gradients on EVERY MLP parameter and on the embeddings, the History after each step."""
    import food101.joint_model_qmf as fq

    class FakeSiglip(nn.Module):
        @classmethod
        def from_pretrained(cls, *a, **k):
            return cls()

        def forward(self, x1, x2):
            return {"text_embeds": x1, "image_embeds": x2}
    fq.AutoModel = FakeSiglip
    net = fq.FusionNet(argparse.Namespace(num_classes=C, num_samples=N), nn.CrossEntropyLoss())
    net.eval()
    food101_fill(net, seed)
    g = torch.Generator().manual_seed(seed)
    rec = {"meta": np.array([B, 768, C, N, steps, seed], dtype=np.int64)}
    for s in range(steps):
        e1 = torch.randn(B, 768, generator=g).requires_grad_(True)
        e2 = torch.randn(B, 768, generator=g).requires_grad_(True)
        y = torch.randint(0, C, (B,), generator=g, dtype=torch.int64)
        idx = (torch.arange(B, dtype=torch.int64) + s * B) % N          # unshuffled loader (food101/run_training.py:39-45)
        net.zero_grad()
        z1, z2, avg, loss, zdf = net(e1, e2, y, idx)
        loss.backward()
        r = {"e1": e1, "e2": e2, "y": y, "idx": idx, "z1": z1, "z2": z2, "avg": avg, "zdf": zdf, "loss": loss,
             "de1": e1.grad, "de2": e2.grad}
        for k, t in r.items():
            rec[f"s{s}_{k}"] = t.detach().numpy().copy()
        for n, p in net.named_parameters():
            if p.grad is None:
                continue
            if p.grad.numel() <= 60000:                                  # heads (mlp.6) and every bias: in full
                rec[f"s{s}_grad/" + n] = p.grad.numpy().copy()
            else:                                                        # hidden weights: two seeded random projections
                gr = torch.Generator().manual_seed(1000 + s)
                rec[f"s{s}_gradR/" + n] = (p.grad @ torch.randn(p.grad.shape[1], 4, generator=gr)).numpy().copy()
                rec[f"s{s}_gradL/" + n] = (torch.randn(4, p.grad.shape[0], generator=gr) @ p.grad).numpy().copy()
        rec[f"s{s}_corr"] = np.stack([h.correctness for h in net.qmf.history]).copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name, "loss(last) =", float(loss.detach()))


def food101_fill(net, seed):
    """Deterministic parameters for the Food101 MLPs (nn.Linear's default range), regenerated the same way by the test
    instead of being stored: values drawn from a seeded CPU generator, parameters visited in sorted name order."""
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in sorted(net.named_parameters()):
            if not (n.startswith("x1_model.") or n.startswith("x2_model.")):
                continue
            bound = 1.0 / np.sqrt(p.shape[-1] if p.dim() > 1 else {0: 768, 3: 512, 6: 512}[int(n.split(".")[2])])
            p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound)


def golden_multi(name, which, B, seed, only=True):
    """The reference's three-modality and unequal-width FusionNets, unmodified, encoders included:
    mustard/joint_model.py (three LstmClassifiers, heads fc3: 100 -> C) and avmnist/joint_model.py (two LeNets,
    heads 48 -> C and 192 -> C).  Forward pre-hooks on the head modules capture the features that reach them (and retain
    their gradients), so the fixture holds exactly what the fused heads kernel sees and must return."""
    import importlib
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    if which == "mustard":
        mod = importlib.import_module("mustard.joint_model")
        C = 2
        net = mod.FusionNet(num_classes=C, loss_fn=nn.CrossEntropyLoss())
        heads = [net.x1_model.fc3, net.x2_model.fc3, net.x3_model.fc3]
        xs = [torch.randn(B, 7, d, generator=g) for d in (371, 81, 300)]
    else:
        mod = importlib.import_module("avmnist.joint_model")
        C = 10
        net = mod.FusionNet(num_classes=C, loss_fn=nn.CrossEntropyLoss())
        heads = [net.classifier_x1, net.classifier_x2]
        xs = [torch.randn(B, 1, 28, 28, generator=g), torch.randn(B, 1, 112, 112, generator=g)]
    y = torch.randint(0, C, (B,), generator=g, dtype=torch.int64)
    seen = {}

    def grab(m):
        def hook(module, inp):
            seen[m] = inp[0]
            inp[0].retain_grad()
        return hook
    for m, h in enumerate(heads):
        h.register_forward_pre_hook(grab(m))
    net.train()
    out = net(*xs, y)
    zs, avg, loss = out[:-2], out[-2], out[-1]
    loss.backward()
    M = len(heads)
    rec = {"meta": np.array([B, C, M], dtype=np.int64), "y": y.numpy(), "avg": avg.detach().numpy(), "loss": loss.detach().numpy(),
           "acc_joint": np.float64(torch.mean((torch.argmax(avg, dim=1) == y).float()))}
    for m in range(M):
        rec[f"f{m}"] = seen[m].detach().numpy().copy(); rec[f"df{m}"] = seen[m].grad.numpy().copy()
        rec[f"W{m}"] = heads[m].weight.detach().numpy().copy(); rec[f"b{m}"] = heads[m].bias.detach().numpy().copy()
        rec[f"dW{m}"] = heads[m].weight.grad.numpy().copy(); rec[f"db{m}"] = heads[m].bias.grad.numpy().copy()
        rec[f"z{m}"] = zs[m].detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name, "loss =", float(loss.detach()), "feature widths", [int(seen[m].shape[1]) for m in range(M)])


if __name__ == "__main__":
    install_shims()
    torch.manual_seed(5)
    # three modalities (mustard) and unequal feature widths (avmnist): SURVEY.md §8f rank 4
    golden_multi("multi_mustard_b24", "mustard", B=24, seed=71)
    golden_multi("multi_avmnist_b40", "avmnist", B=40, seed=73)
    if "--multi-only" in sys.argv:
        sys.exit(0)
    # K2: Crema-D QMF, B=64 D=512 C=6, sampling with replacement (duplicates), 4 steps of history
    golden_qmf("qmf_cremad_b64", B=64, D=512, C=6, N=997, steps=3, seed=5)
    # ragged / odd sizes, contiguous idx windows (Food101-style loader), C not a multiple of anything
    golden_qmf("qmf_small_c11", B=37, D=64, C=11, N=101, steps=3, seed=7, idx_mode="contiguous")
    # wide head, Food101 class count (smaller batch so the (B,B) reference stays small)
    golden_qmf("qmf_food_c101", B=40, D=256, C=101, N=500, steps=2, seed=11)
    golden_qmf("qmf_d768_c7", B=8, D=768, C=7, N=64, steps=2, seed=29)
    # minimum legal batch for reg_loss
    golden_qmf("qmf_b2", B=2, D=32, C=3, N=5, steps=3, seed=13)
    # loss-term ablations and the QMF + OGM-GE model: same FusionNet, one line differs
    golden_qmf("qmf_ablate_ljoint_b48", B=48, D=512, C=6, N=211, steps=3, seed=31, module="cremad.joint_model_qmf_ablate_Ljoint")
    golden_qmf("qmf_ablate_lunimodal_b48", B=48, D=512, C=6, N=211, steps=3, seed=37, module="cremad.joint_model_qmf_ablate_Lunimodal")
    golden_qmf("qmf_ablate_ljoint_c101", B=36, D=128, C=101, N=300, steps=2, seed=41, module="cremad.joint_model_qmf_ablate_Ljoint")
    golden_qmf("qmf_ogm_ge_lreg_b48", B=48, D=512, C=6, N=211, steps=2, seed=43, module="cremad.joint_model_ogm_ge_lreg")
    # K3-shaped OGM-GE heads (small batch: the reference score loop is O(B^2 C))
    golden_ogm("ogm_cremad_b48", B=48, D=512, C=6, steps=3, seed=5, alpha=0.8)
    golden_ogm("ogm_wide_c309", B=40, D=128, C=309, steps=2, seed=17, alpha=0.8)
    # K1: Enrico jlogits, B=32 D=512 C=20 (frozen encoders: no dfeat)
    golden_ogm("jlogits_enrico_b32", B=32, D=512, C=20, steps=2, seed=19, alpha=0.1, enrico=True)
    golden_modulate("ogm_modulate_small", seed=23)
    # the per-modality ensemble with OGM-GE (model_type ensemble_ogm_ge)
    golden_ensemble("ensemble_ogm_b48", B=48, D=512, C=6, steps=3, seed=47, alpha=0.8)
    golden_ensemble("ensemble_wide_c101", B=40, D=128, C=101, steps=2, seed=53, alpha=0.8)
    # epoch-end unimodal offset correction of the reference's own LightningModule (utils/BaseModel.py:161-202)
    golden_epoch_end("epoch_end_c6", n_batches=5, B=64, C=6, seed=59)
    golden_epoch_end("epoch_end_c101", n_batches=3, B=96, C=101, seed=61)
    # the Food101 module itself (K4's model): MLP heads on SigLIP embeddings, QMF loss
    golden_food101("food101_module_b32", B=32, C=101, N=256, steps=2, seed=67)
