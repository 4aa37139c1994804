"""The LITERAL reference on the CPU: the unmodified modules staged under oracle/_ref/ (oracle/vendor_ref.py) driven through
their own ``FusionNet.forward`` / ``loss.backward()`` / ``EMA.update`` / ``ogm_ge`` on synthetic features.
TEST / BASELINE INFRASTRUCTURE: imported by ``bench.py`` (cpu_baseline, --impl reference) and tests only.

Recipe: SURVEY.md Appendix B -- a minimal ``pytorch_lightning`` shim (the package is not installed), identity encoders,
``Tensor.cuda`` patched to identity because existing_algos/QMF.py:63,66 hard-codes ``.cuda()``."""
import argparse
import os
import sys
import time
import types

import torch
import torch.nn as nn

from . import vendor_ref

_installed = False


def available() -> bool:
    return vendor_ref.verify()


def _install():
    global _installed
    if _installed:
        return
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(nn.Module):
        automatic_optimization = True

        def log(self, *a, **k):
            pass
    pl.LightningModule = LightningModule
    pl.seed_everything = lambda s, workers=False: torch.manual_seed(s)
    pl.loggers = types.ModuleType("pytorch_lightning.loggers")
    pl.loggers.WandbLogger = object
    pl.callbacks = types.SimpleNamespace(LearningRateMonitor=object, ModelCheckpoint=object)
    sys.modules.setdefault("pytorch_lightning", pl)
    sys.modules.setdefault("pytorch_lightning.loggers", pl.loggers)
    if "sympy" not in sys.modules:
        try:
            import sympy  # noqa: F401  (stray import at cremad/joint_model_qmf.py:1)
        except Exception:
            sy = types.ModuleType("sympy")
            sy.Idx = object
            sys.modules["sympy"] = sy
    sys.path.insert(0, vendor_ref.DST)
    _installed = True


def time_step(mode: str, B: int, D: int, C: int, N, alpha, budget_s: float = 15.0, max_steps: int = 5):
    """Time the literal reference step (forward + backward + EMA [+ ogm_ge]) at batch B on all host cores.
    -> (samples/s, steps, seconds).  The reference is quadratic in B (QMF.reg_loss builds (B,B) matrices, ogm_ge's
    score loop is O(B^2 C) Python): callers pass a bounded B and say so."""
    _install()
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self      # CPU run of QMF.py:63,66; restored below (the caller may use a GPU later)
    try:
        return _time_step(mode, B, D, C, N, alpha, budget_s, max_steps)
    finally:
        torch.Tensor.cuda = real_cuda


def _time_step(mode, B, D, C, N, alpha, budget_s, max_steps):
    from utils.EMA import EMA
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(7)
    if mode == "qmf":
        import cremad.joint_model_qmf as mq
        mq.resnet18 = lambda modality: nn.Identity()
        net = mq.FusionNet(argparse.Namespace(num_classes=C, num_samples=N), nn.CrossEntropyLoss())
    else:
        import cremad.joint_model_ogm_ge as mo
        mo.resnet18 = lambda modality: nn.Identity()
        net = mo.FusionNet(C, nn.CrossEntropyLoss())
    net.x1_classifier = nn.Linear(D, C)
    net.x2_classifier = nn.Linear(D, C)
    ema = EMA(torch.zeros(2, C))
    f1 = torch.randn(B, D, 1, 1, generator=g); f2 = torch.randn(B, D, 1, 1, generator=g)
    y = torch.randint(0, C, (B,), generator=g)
    idx = torch.arange(B) % (N or 1)

    def one():
        a = f1.clone().requires_grad_(True); v = f2.clone().requires_grad_(True)
        net.zero_grad()
        out = net(a, v, y, idx) if mode == "qmf" else net(a, v, y)
        out[3].backward()
        ema.update(torch.mean(torch.stack([out[0], out[1]]), dim=1))
        return float(out[3])

    one()
    n, t0 = 0, time.perf_counter()
    while n < max_steps:
        one(); n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return B * n / dt, n, dt
