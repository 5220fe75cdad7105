import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def rel_err(a, b):
    """max |a - b| / max |b|  (the 'relative' of BASELINE.json's 1e-5 bar)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().item()
    if denom == 0.0:
        return (a - b).abs().max().item()
    return (a - b).abs().max().item() / denom


def row_rel_err(a, b, floor=0.05):
    """Per-row error: max over rows r of max|a_r - b_r| / max(max|b_r|, floor * max|b|).  Tighter than ``rel_err`` for
    small-magnitude rows (atoms with few neighbours, small parameters); the floor keeps all-zero rows meaningful."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    a, b = a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1)
    gmax = b.abs().max().item()
    if gmax == 0.0:
        return (a - b).abs().max().item()
    denom = b.abs().max(dim=1).values.clamp_min(floor * gmax)
    return ((a - b).abs().max(dim=1).values / denom).max().item()
