import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name, dtype=None):
    """tests/golden/<name>.npz -> dict of torch tensors (floats optionally cast to `dtype`)."""
    out = {}
    with np.load(os.path.join(GOLDEN, name + '.npz')) as z:
        for k in z.files:
            t = torch.from_numpy(z[k])
            if dtype is not None and t.is_floating_point():
                t = t.to(dtype)
            out[k] = t
    return out


def sub(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def rel_err(a, b):
    """max-norm relative error max|a-b| / max(|b|, tiny)."""
    a = torch.as_tensor(a).detach().to(torch.float64).cpu()
    b = torch.as_tensor(b).detach().to(torch.float64).cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def l2_rel_err(a, b):
    a = torch.as_tensor(a).detach().to(torch.float64).cpu()
    b = torch.as_tensor(b).detach().to(torch.float64).cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
