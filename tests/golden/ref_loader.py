"""Import the UNMODIFIED reference (cassianobecker/tgcn) in the build container.

Only used by tests/golden/make_golden.py (fixture generation) -- never at test
or bench run time: /root/reference does not exist on the GPU box.

Three import stubs are needed:
  * matplotlib.pyplot      -- gcn/graph.py:3 imports it, the hot path never calls it
  * torch_geometric.utils  -- tgcn/nn/gcn.py:4 (degree, remove_self_loops: edge-index layers only)
  * torch_scatter          -- tgcn/nn/gcn.py:5 (scatter_add: edge-index layers only)
"""
import importlib
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("TGCN_REF", "/root/reference")


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def _unavailable(*a, **k):
    raise RuntimeError("stubbed third-party symbol: not on the hot path")


def load():
    """Return (gcn_graph, gcn_coarsening, tgcn_nn_gcn, tgcn_nn_gcn_matmul) reference modules."""
    if not os.path.isdir(REF_ROOT):
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    tg = _stub("torch_geometric")
    tg.utils = _stub("torch_geometric.utils", degree=_unavailable, remove_self_loops=_unavailable)
    _stub("torch_scatter", scatter_add=_unavailable)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # SyntaxWarning: "is 1" in gcn/coarsening.py:192,197
        g = importlib.import_module("gcn.graph")
        c = importlib.import_module("gcn.coarsening")
        n = importlib.import_module("tgcn.nn.gcn")
        m = importlib.import_module("tgcn.nn.gcn_matmul")
    return g, c, n, m
