"""bench.py's CPU arms build their graphs from oracle/ (restatements of gcn/graph.py, gcn/coarsening.py) and the pure
generators in tgcn_b200/synth.py, never from the product's graph / coarsening code: the two builds must give the SAME
operands (permutation and every level's rescaled Laplacian, bit for bit), or the arms would time different problems."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", ["hcp360", "hcp360dense", "mnist", "mesh32k"])
def test_oracle_and_product_graph_builds_agree(name):
    b = _bench()
    g1, p1, L1, n1 = b.build_graph(name)
    g2, p2, L2, n2 = b.build_graph_oracle(name)
    assert n1 == n2 and list(p1) == list(p2)
    assert len(L1) == len(L2)
    for a, c in zip(L1, L2):
        assert a.shape == c.shape and a.dtype == c.dtype == np.float32
        assert np.array_equal(a.indptr, c.indptr) and np.array_equal(a.indices, c.indices)
        assert np.array_equal(a.data, c.data)


def test_reference_arm_config_equals_b200_arm_config():
    """`config` holds workload keys only, so both arms print the identical dict (driver check `same_config`)."""
    b = _bench()
    g, p, Ls, n = b.build_graph_oracle("hcp360")
    c1 = b.workload_config("hcp360", 4, 64, Ls, 15)
    g, p, Lp, n = b.build_graph("hcp360")
    c2 = b.workload_config("hcp360", 4, 64, Lp, 15)
    assert c1 == c2 and c1["parallelism"] == "dp4" and c1["global_batch"] == 256
