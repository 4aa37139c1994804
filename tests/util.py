"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# BASELINE.json north_star tolerances
TOL_FP32 = 1e-5     # relative, fp32 path
TOL_TENSOR = 2e-2   # relative, reduced-precision (bf16 / tf32 tensor-pipe) path


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def relerr(a, b):
    """Norm-wise relative error ||a-b|| / ||b|| (SURVEY.md §8c 'compare norm-wise')."""
    a = np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64)
    b = np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64)
    den = np.linalg.norm(b.ravel())
    num = np.linalg.norm((a - b).ravel())
    if den == 0:
        return num
    return num / den


def assert_close(a, b, tol, what=""):
    e = relerr(a, b)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e}"


def t(x):
    return torch.from_numpy(np.asarray(x))


def cu(x):
    return t(x).cuda()


def loss_terms_of(name: str) -> int:
    """LF_LOSS_* bits of a golden fixture generated from one of the reference's QMF loss ablations."""
    return 1 if "ljoint" in name else 2 if "lunimodal" in name else 0
