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


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def csr_from(rec, prefix):
    import scipy.sparse as sp
    shape = tuple(int(v) for v in rec[prefix + "_shape"])
    return sp.csr_matrix((rec[prefix + "_data"], rec[prefix + "_indices"], rec[prefix + "_indptr"]), shape=shape)


def rel_err(a, ref):
    """Scale-relative error used for every floating-point parity statement in this repo:
    max|a - ref| / max|ref|  (north_star: 1e-4 relative tolerance in fp32)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = np.max(np.abs(ref)) if ref.size else 1.0
    if scale == 0:
        scale = 1.0
    return float(np.max(np.abs(a - ref)) / scale) if ref.size else 0.0


LAYER_CASES = sorted(f for f in os.listdir(GOLDEN) if f.startswith("layer_") and f.endswith(".npz"))
GRAPH_CASES = sorted(f for f in os.listdir(GOLDEN) if f.startswith("graph_") and f.endswith(".npz"))
