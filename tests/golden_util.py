import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    """npz -> dict of torch tensors (int32 -> int64, as written by tests/golden/make_golden.py)."""
    out = {}
    with np.load(os.path.join(GOLDEN, name + '.npz')) as z:
        for k in z.files:
            t = torch.from_numpy(z[k])
            out[k] = t.to(torch.int64) if t.dtype == torch.int32 else t
    return out


def same_sets(a, b, dim):
    return torch.equal(torch.sort(a, dim=dim)[0], torch.sort(b, dim=dim)[0])
