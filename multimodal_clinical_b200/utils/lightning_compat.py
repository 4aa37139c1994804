"""``pytorch_lightning`` when it is installed, otherwise the small subset of it this path uses.

The reference is driven by PyTorch-Lightning 2.1 (utils/run_trainer.py); the image this framework runs in
does not ship it.  The fallback below implements exactly the hooks the reference's modules rely on --
``LightningModule`` (log / optimizers / lr_schedulers / manual_backward / automatic_optimization),
``Trainer.fit`` / ``Trainer.test`` calling the hooks in Lightning's order, ``ModelCheckpoint`` on a
monitored metric, ``LearningRateMonitor`` and ``seed_everything`` -- so ``main.py --dir`` runs unchanged.
"""
from __future__ import annotations

import os
import random
from typing import Any, Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

try:  # pragma: no cover - exercised only where Lightning exists
    import pytorch_lightning as pl
    from pytorch_lightning import LightningModule, Trainer, seed_everything
    from pytorch_lightning.callbacks import LearningRateMonitor, ModelCheckpoint
    HAVE_LIGHTNING = True
except Exception:
    HAVE_LIGHTNING = False

    def seed_everything(seed: int, workers: bool = False) -> int:
        random.seed(seed); np.random.seed(seed); torch.manual_seed(seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(seed)
        os.environ["PL_GLOBAL_SEED"] = str(seed)
        return seed

    class LightningModule(nn.Module):
        automatic_optimization = True

        def __init__(self):
            super().__init__()
            self.trainer: Optional["Trainer"] = None
            self.logged: Dict[str, Any] = {}

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")

        def log(self, name, value, **kwargs):
            self.logged[name] = value
            if self.trainer is not None:
                self.trainer.callback_metrics[name] = value

        def optimizers(self):
            opts = self.trainer.optimizers
            return opts[0] if len(opts) == 1 else opts

        def lr_schedulers(self):
            s = self.trainer.schedulers
            return s[0] if len(s) == 1 else s

        def manual_backward(self, loss, *a, **k):
            loss.backward(*a, **k)

        def configure_optimizers(self):
            raise NotImplementedError

        # hooks a module may leave out (Lightning's defaults are no-ops too)
        def on_train_epoch_end(self) -> None:
            pass

        def on_validation_epoch_end(self) -> None:
            pass

        def on_test_epoch_end(self) -> None:
            pass

    class LearningRateMonitor:
        def __init__(self, logging_interval: str = "epoch"):
            self.logging_interval = logging_interval

    class ModelCheckpoint:
        def __init__(self, dirpath: str, filename: str, save_top_k: int = 1, monitor: str = "", mode: str = "max"):
            self.dirpath, self.filename, self.monitor, self.mode = dirpath, filename, monitor, mode
            self.best_model_score = None
            self.best_model_path = ""

        def on_validation_end(self, trainer, module):
            v = trainer.callback_metrics.get(self.monitor)
            if v is None:
                return
            v = float(v)
            better = self.best_model_score is None or (v > self.best_model_score if self.mode == "max" else v < self.best_model_score)
            if better:
                os.makedirs(self.dirpath, exist_ok=True)
                self.best_model_score = v
                self.best_model_path = os.path.join(self.dirpath, self.filename + ".ckpt")
                torch.save({"state_dict": module.state_dict(), "epoch": trainer.current_epoch}, self.best_model_path)

    class Trainer:
        """fit: per epoch [train batches -> validation loop -> on_train_epoch_end -> epoch schedulers],
        the order Lightning 2.x uses; ``precision='bf16-mixed'`` wraps the steps in CUDA autocast."""

        def __init__(self, max_epochs: int = 1, precision: str = "32-true", callbacks: Optional[List[Any]] = None,
                     overfit_batches=0, limit_train_batches: Optional[int] = None, **_ignored):
            self.max_epochs = max_epochs
            self.precision = precision
            self.callbacks = callbacks or []
            self.overfit_batches = overfit_batches
            self.limit_train_batches = limit_train_batches
            self.callback_metrics: Dict[str, Any] = {}
            self.optimizers: List[torch.optim.Optimizer] = []
            self.schedulers: List[Any] = []
            self.current_epoch = 0
            self.global_step = 0

        def _autocast(self, device):
            on = self.precision == "bf16-mixed" and device.type == "cuda"
            return torch.autocast("cuda", dtype=torch.bfloat16, enabled=on)

        @staticmethod
        def _to(batch, device):
            return tuple(b.to(device, non_blocking=True) if torch.is_tensor(b) else b for b in batch)

        def _configure(self, module):
            cfg = module.configure_optimizers()
            opts, scheds = (cfg if isinstance(cfg, tuple) and len(cfg) == 2 else (cfg, []))
            self.optimizers = list(opts) if isinstance(opts, (list, tuple)) else [opts]
            self.schedulers = [s["scheduler"] if isinstance(s, dict) else s for s in (scheds or [])]

        def fit(self, module, train_dataloaders=None, val_dataloaders=None):
            module.trainer = self
            device = module.device
            self._configure(module)
            for epoch in range(self.max_epochs):
                self.current_epoch = epoch
                module.train()
                for i, batch in enumerate(train_dataloaders):
                    if self.limit_train_batches is not None and i >= self.limit_train_batches:
                        break
                    batch = self._to(batch, device)
                    with self._autocast(device):
                        if module.automatic_optimization:
                            opt = self.optimizers[0]
                            opt.zero_grad()
                            loss = module.training_step(batch, i)
                            loss.backward()
                            opt.step()
                        else:
                            module.training_step(batch, i)
                    self.global_step += 1
                if val_dataloaders is not None:
                    self._eval_loop(module, val_dataloaders, "validation")
                    for cb in self.callbacks:
                        if hasattr(cb, "on_validation_end"):
                            cb.on_validation_end(self, module)
                module.train()
                module.on_train_epoch_end()
                if module.automatic_optimization:
                    for s in self.schedulers:
                        s.step()

        def _eval_loop(self, module, loader, kind):
            module.eval()
            device = module.device
            step = getattr(module, f"{kind}_step")
            with torch.no_grad():
                for i, batch in enumerate(loader):
                    with self._autocast(device):
                        step(self._to(batch, device), i)
            getattr(module, f"on_{kind}_epoch_end")()

        def test(self, module, dataloaders=None):
            module.trainer = self
            self._eval_loop(module, dataloaders, "test")
            return [dict(self.callback_metrics)]

    class _Callbacks:
        LearningRateMonitor = LearningRateMonitor
        ModelCheckpoint = ModelCheckpoint

    class _PL:
        LightningModule = LightningModule
        Trainer = Trainer
        callbacks = _Callbacks
        seed_everything = staticmethod(seed_everything)

    pl = _PL()
