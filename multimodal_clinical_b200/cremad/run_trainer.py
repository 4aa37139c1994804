"""``run_training()`` for Crema-D (cremad/run_trainer.py of the reference): configs -> data -> loaders ->
model -> run_trainer.  ``data_path: synthetic`` (or a missing get_data module) swaps in synthetic tensors of
the reference's batch layout; a user-supplied ``cremad_get_data`` module with ``get_data(args)`` is used
when importable."""
from torch.utils.data import DataLoader

from ..synthetic_data import splits
from ..utils.run_trainer import run_trainer
from ..utils.setup_configs import setup_configs
from . import get_model


def _datasets(args):
    try:
        from cremad_get_data import get_data            # user-provided loader for the real corpus
        return get_data(args)
    except ImportError:
        n = int(getattr(args, "synthetic_samples", 256))
        # audio spectrogram (1,H,W), visual (3,T,H,W) -- small so the ResNet18 producers stay cheap
        return splits(n, (1, 65, 65), (3, 2, 64, 64), args.num_classes, with_idx=('qmf' in args.model_type or 'lreg' in args.model_type), seed=args.seed)


def run_training(argv=None):
    args = setup_configs(argv)
    train_dataset, val_dataset, test_dataset = _datasets(args)
    setattr(args, "num_samples", len(train_dataset))
    kw = dict(batch_size=args.batch_size, num_workers=0)
    train_loader = DataLoader(train_dataset, shuffle=True, **kw)
    val_loader = DataLoader(val_dataset, **kw)
    test_loader = DataLoader(test_dataset, **kw)
    model = get_model(args)
    return run_trainer(args, model, train_loader, val_loader, test_loader)
