import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


def topk_sets_equal(a, b, k_dim=2):
    """Compare top-k index tensors as sets along k_dim (torch.topk tie/sort order is unspecified)."""
    return torch.equal(torch.sort(a, dim=k_dim)[0], torch.sort(b, dim=k_dim)[0])


from oracle.compare import children_rows, topk_bad_rows  # noqa: E402,F401
