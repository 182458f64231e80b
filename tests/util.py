import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel_err(a, b):
    """max-norm relative error: ||a-b||_inf / ||b||_inf (the parity metric of SURVEY.md §8d)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def dice_per_class(a, b, C):
    """metric/utils.py:62-82 style one-hot Dice between two label maps."""
    out = []
    for c in range(C):
        x, y = (a == c), (b == c)
        den = x.sum().item() + y.sum().item()
        out.append(1.0 if den == 0 else 2.0 * (x & y).sum().item() / den)
    return out
