"""``run_training()`` for AV-MNIST (avmnist/run_training.py of the reference): batches are (image (B, 1, 28, 28),
spectrogram (B, 1, 112, 112), label)."""
from ..synthetic_data import tuple_splits
from ..utils.run_multi import fit_and_test, load_args, packaged_yaml
from . import get_model


def _datasets(args):
    try:
        from avmnist_get_data import get_data            # user-provided loader for the real corpus
        return get_data(args.data_path)
    except ImportError:
        n = int(getattr(args, "synthetic_samples", 128))
        return tuple_splits(n, [(1, 28, 28), (1, 112, 112)], args.num_classes, seed=args.seed)


def run_training(argv=None):
    args = load_args(argv, packaged_yaml(__file__, "avmnist.yaml"))
    return fit_and_test(args, get_model(args), _datasets(args))
