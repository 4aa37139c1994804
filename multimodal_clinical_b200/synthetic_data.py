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
