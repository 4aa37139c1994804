"""Synthetic stand-ins for the reference's datasets (cremad/get_data.py etc. need librosa / timm / the raw
corpora, none of which exist offline).  Shapes follow the reference batches; indices are returned for QMF."""
import torch
from torch.utils.data import Dataset


class SyntheticPairs(Dataset):
    def __init__(self, n, shape1, shape2, num_classes, with_idx, seed=5):
        g = torch.Generator().manual_seed(seed)
        self.label = torch.randint(0, num_classes, (n,), generator=g)
        # a weak class signal in both modalities so accuracies move during a smoke run
        self.x1 = torch.randn(n, *shape1, generator=g) + self.label.view(-1, *([1] * len(shape1))) * 0.1
        self.x2 = torch.randn(n, *shape2, generator=g) - self.label.view(-1, *([1] * len(shape2))) * 0.1
        self.with_idx = with_idx

    def __len__(self):
        return self.label.numel()

    def __getitem__(self, i):
        if self.with_idx:
            return self.x1[i], self.x2[i], self.label[i], torch.tensor(i)
        return self.x1[i], self.x2[i], self.label[i]


def splits(n_train, shape1, shape2, num_classes, with_idx, seed=5):
    mk = lambda n, s: SyntheticPairs(n, shape1, shape2, num_classes, with_idx, seed=s)
    return mk(n_train, seed), mk(max(n_train // 4, 8), seed + 1), mk(max(n_train // 4, 8), seed + 2)


class SyntheticTuples(Dataset):
    """(x_1, ..., x_M, label) for the M-modality loaders (mustard: (S, 371), (S, 81), (S, 300); avmnist: (1, 28, 28), (1, 112, 112))."""

    def __init__(self, n, shapes, num_classes, seed=5):
        g = torch.Generator().manual_seed(seed)
        self.label = torch.randint(0, num_classes, (n,), generator=g)
        self.xs = [torch.randn(n, *s, generator=g) + self.label.view(-1, *([1] * len(s))) * (0.1 if m % 2 == 0 else -0.1)
                   for m, s in enumerate(shapes)]

    def __len__(self):
        return self.label.numel()

    def __getitem__(self, i):
        return tuple(x[i] for x in self.xs) + (self.label[i],)


def tuple_splits(n_train, shapes, num_classes, seed=5):
    mk = lambda n, s: SyntheticTuples(n, shapes, num_classes, seed=s)
    return mk(n_train, seed), mk(max(n_train // 4, 8), seed + 1), mk(max(n_train // 4, 8), seed + 2)
