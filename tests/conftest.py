"""pytest configuration: the ``gpu`` marker, repo root on sys.path, golden loader."""

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = os.environ.get("SCATT_REFERENCE", "/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """Return ``(arrays, meta)`` of ``tests/golden/<name>.npz`` as torch tensors."""
    import torch

    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    arrays, meta = {}, {}
    for k in z.files:
        v = z[k]
        if k in ("meta", "cfg"):
            meta[k] = json.loads(str(v))
        elif v.ndim == 0:
            meta[k] = float(v)
        else:
            arrays[k] = torch.from_numpy(v)
    return arrays, meta


@pytest.fixture(scope="session")
def golden():
    return load_golden
