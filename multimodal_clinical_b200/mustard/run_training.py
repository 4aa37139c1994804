"""``run_training()`` for MUStARD (mustard/run_training.py of the reference): batches are (image (B, S, 371), audio (B, S, 81),
text (B, S, 300), label)."""
from ..synthetic_data import tuple_splits
from ..utils.run_multi import fit_and_test, load_args, packaged_yaml
from . import get_model


def _datasets(args):
    try:
        from mustard_get_data import get_data            # user-provided loader for the real corpus (needs sarcasm.pkl)
        return get_data(args.data_path, max_pad=True, task='classification', data_type='sarcasm', max_seq_len=args.max_seq_len)
    except ImportError:
        n = int(getattr(args, "synthetic_samples", 128))
        S = int(args.max_seq_len)
        return tuple_splits(n, [(S, 371), (S, 81), (S, 300)], args.num_classes, seed=args.seed)


def run_training(argv=None):
    args = load_args(argv, packaged_yaml(__file__, "mustard.yaml"))
    return fit_and_test(args, get_model(args), _datasets(args))
