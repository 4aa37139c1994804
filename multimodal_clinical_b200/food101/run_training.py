"""``run_training()`` for Food101 (food101/run_training.py of the reference); unshuffled loaders as there."""
from torch.utils.data import DataLoader

from ..synthetic_data import splits
from ..utils.run_trainer import run_trainer
from ..utils.setup_configs import setup_configs
from . import get_model


def _datasets(args):
    try:
        from food101_get_data import get_data           # user-provided loader for the real corpus
        return get_data(args)
    except ImportError:
        n = int(getattr(args, "synthetic_samples", 1024))
        return splits(n, (768,), (768,), args.num_classes, with_idx=('qmf' in args.model_type or 'lreg' in args.model_type), seed=args.seed)


def run_training(argv=None):
    args = setup_configs(argv)
    train_dataset, val_dataset, test_dataset = _datasets(args)
    setattr(args, "num_samples", len(train_dataset))
    kw = dict(batch_size=args.batch_size, num_workers=0)
    model = get_model(args)
    return run_trainer(args, model, DataLoader(train_dataset, **kw), DataLoader(val_dataset, **kw),
                       DataLoader(test_dataset, **kw))
