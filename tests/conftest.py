import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` tests are skipped (not failed) when no device is present, e.g. a plain `pytest tests/` here.
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def rel_err(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


@pytest.fixture(scope="session")
def dumbbell():
    d = np.load(os.path.join(GOLDEN, "dumbbell_data.npz"))
    x = torch.from_numpy(d["sampled_x"])
    y = torch.from_numpy(d["sampled_y"])
    t = torch.from_numpy(d["test_idx"])
    return {"train_x": x[~t].contiguous(), "test_x": x[t].contiguous(),
            "train_y": y[~t].contiguous(), "test_y": y[t].contiguous()}


@pytest.fixture(scope="session")
def golden_k10():
    return dict(np.load(os.path.join(GOLDEN, "dumbbell_k10.npz")))


@pytest.fixture(scope="session")
def golden_k50():
    return dict(np.load(os.path.join(GOLDEN, "dumbbell_k50.npz")))


PARAM_SETS_K10 = [(0.5, 1.3, 2), (0.5, 0.5, 1), (0.05, 0.7, 3)]
NORMALIZATIONS = ["symmetric", "randomwalk"]


def gtag(eps, kappa, nu, normalization, self_loops):
    return f"e{eps}_k{kappa}_nu{nu}_{normalization}_{'sl' if self_loops else 'nosl'}"
