import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def first_occurrence_numpy(pose, frame):
    """Stand-in for the device call mval_first_occurrence in the CPU tests that stub the kernels out: the dict-insertion
    semantics of strategy.py:1115-1133 (a repeated guid keeps its first position and its last value)."""
    import torch

    keys = list(zip(pose.tolist(), frame.tolist()))
    first, last = {}, {}
    for i, k in enumerate(keys):
        first.setdefault(k, i)
        last[k] = i
    keep = torch.zeros(len(keys), dtype=torch.uint8)
    src = torch.full((len(keys),), -1, dtype=torch.int32)
    for k, i in first.items():
        keep[i] = 1
        src[i] = last[k]
    return keep, src, torch.tensor([len(first)], dtype=torch.int32)
