"""Pin oracle/graph_np.py + oracle/coarsening_np.py: the reference's own KAT
(gcn/coarsening.py:216-217) and golden outputs of the unmodified reference."""
import numpy as np
import pytest

from conftest import GRAPH_CASES, csr_from, load_golden
from oracle import coarsening_np, graph_np


def test_compute_perm_kat_from_reference():
    # verbatim known-answer test of the reference, gcn/coarsening.py:216-217
    got = coarsening_np.compute_perm([np.array([4, 1, 1, 2, 2, 3, 0, 0, 3]), np.array([2, 1, 0, 1, 0])])
    assert got == [[3, 4, 0, 9, 1, 2, 5, 8, 6, 7, 10, 11], [2, 4, 1, 3, 0, 5], [0, 1, 2]]


def same_csr(A, B):
    A = A.tocsr().copy(); A.sort_indices()
    B = B.tocsr().copy(); B.sort_indices()
    return (A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
            and np.array_equal(A.data, B.data) and A.data.dtype == B.data.dtype)


def test_grid_knn_adjacency_bit_exact():
    r = load_golden("knn_grid28_k8.npz")
    z = graph_np.grid_embedding(28)
    assert np.array_equal(z, r["z"]) and z.dtype == r["z"].dtype
    d, idx = graph_np.knn_exact(z, k=8)
    assert np.array_equal(d, r["dist"]) and np.array_equal(idx, r["idx"])
    assert same_csr(graph_np.knn_adjacency(d, idx), csr_from(r, "A"))


@pytest.mark.parametrize("case", GRAPH_CASES)
def test_coarsen_bit_exact(case):
    r = load_golden(case)
    A = csr_from(r, "A")
    levels, seed = int(r["levels"]), int(r["seed"])
    np.random.seed(seed)
    graphs, parents = coarsening_np.metis(A, levels)
    for i, p in enumerate(parents):
        assert np.array_equal(p, r["parents_%d" % i])
    for i, g in enumerate(graphs):
        assert same_csr(g, csr_from(r, "metis_graph_%d" % i))
    perms = coarsening_np.compute_perm(parents)
    for i, p in enumerate(perms):
        assert np.array_equal(np.asarray(p), r["perms_%d" % i])
    np.random.seed(seed)
    cgraphs, perm, _, _ = coarsening_np.coarsen(A, levels)
    assert np.array_equal(np.asarray(perm), r["perm"])
    for i, g in enumerate(cgraphs):
        assert same_csr(g, csr_from(r, "graph_%d" % i))
        L = graph_np.rescale_laplacian(graph_np.laplacian(g, normalized=True), lmax=2)
        assert same_csr(L, csr_from(r, "L_%d" % i))


def test_perm_data_bit_exact():
    r = load_golden("perm_data.npz")
    perm = list(r["perm"])
    y2 = coarsening_np.perm_data(r["x2"], perm)
    assert y2.dtype == np.float64 and np.array_equal(y2, r["y2"])
    assert np.array_equal(coarsening_np.perm_data_time(r["x3"], perm), r["y3"])
