"""Product host-side graph prep (tgcn_b200.graph / tgcn_b200.coarsening) is bit-exact with the
reference goldens, with the reference's own KAT, and with the oracle restatement."""
import numpy as np
import pytest

from conftest import GRAPH_CASES, csr_from, load_golden
from oracle import coarsening_np
from tgcn_b200 import coarsening, graph


def same_csr(A, B):
    A = A.tocsr().copy(); A.sort_indices()
    B = B.tocsr().copy(); B.sort_indices()
    return (A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
            and np.array_equal(A.data, B.data) and A.data.dtype == B.data.dtype)


def test_compute_perm_kat_from_reference():
    got = coarsening.compute_perm([np.array([4, 1, 1, 2, 2, 3, 0, 0, 3]), np.array([2, 1, 0, 1, 0])])
    assert got == [[3, 4, 0, 9, 1, 2, 5, 8, 6, 7, 10, 11], [2, 4, 1, 3, 0, 5], [0, 1, 2]]
    assert coarsening.compute_perm([]) == []


def test_grid_knn_adjacency_bit_exact():
    r = load_golden("knn_grid28_k8.npz")
    z = graph.grid(28)
    assert np.array_equal(z, r["z"]) and z.dtype == r["z"].dtype
    d, idx = graph.distance_sklearn_metrics(z, k=8)
    assert np.array_equal(d, r["dist"]) and np.array_equal(idx, r["idx"])
    assert same_csr(graph.adjacency(d, idx), csr_from(r, "A"))


@pytest.mark.parametrize("case", GRAPH_CASES)
def test_coarsen_bit_exact(case):
    r = load_golden(case)
    A = csr_from(r, "A")
    levels, seed = int(r["levels"]), int(r["seed"])
    np.random.seed(seed)
    graphs, parents = coarsening.metis(A, levels)
    for i, p in enumerate(parents):
        assert np.array_equal(p, r["parents_%d" % i])
    for i, g in enumerate(graphs):
        assert same_csr(g, csr_from(r, "metis_graph_%d" % i))
    perms = coarsening.compute_perm(parents)
    for i, p in enumerate(perms):
        assert np.array_equal(np.asarray(p), r["perms_%d" % i])
    np.random.seed(seed)
    cgraphs, perm = coarsening.coarsen(A, levels)
    assert np.array_equal(np.asarray(perm), r["perm"])
    for i, g in enumerate(cgraphs):
        assert same_csr(g, csr_from(r, "graph_%d" % i))
        L = graph.rescale_L(graph.laplacian(g, normalized=True), lmax=2)
        assert same_csr(L, csr_from(r, "L_%d" % i))
        assert same_csr(graph.rescaled_laplacian_csr(g), csr_from(r, "L_%d" % i))


def test_rescale_L_inplace_flag_and_dense_input():
    r = load_golden("graph_dense40_seed3.npz")
    g = csr_from(r, "graph_0")
    L = graph.laplacian(g)
    keep = L.copy()
    out = graph.rescale_L(L, lmax=2)
    assert (L != keep).nnz == 0                                  # not mutated by default
    out2 = graph.rescale_L(L, lmax=2, inplace=True)
    assert same_csr(out, out2)                                   # (scipy rebinds on `-=`: same values either way)
    dense = graph.rescale_L(graph.laplacian(g).todense(), lmax=2)  # what the example scripts do
    assert np.array_equal(np.asarray(dense), out.toarray())


def test_perm_data_bit_exact():
    r = load_golden("perm_data.npz")
    perm = list(r["perm"])
    y2 = coarsening.perm_data(r["x2"], perm)
    assert y2.dtype == np.float64 and np.array_equal(y2, r["y2"])
    assert np.array_equal(coarsening.perm_data_time(r["x3"], perm), r["y3"])
    assert coarsening.perm_data(r["x2"], None) is r["x2"]


def test_native_pairing_matches_oracle_on_random_graphs():
    """Includes float64 weights and graphs where the reference's row-extent quirk matters."""
    import scipy.sparse as sp
    for seed in range(6):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(20, 200))
        dt = np.float32 if seed % 2 == 0 else np.float64
        M = sp.random(n, n, density=0.08, random_state=seed, dtype=np.float64)
        M = (M + M.T).tocsr().astype(dt)
        M.setdiag(0); M.eliminate_zeros()
        ring = sp.coo_matrix((np.full(n, 0.25, dt), (np.arange(n), (np.arange(n) + 1) % n)), shape=(n, n))
        M = (M + ring + ring.T).tocsr().astype(dt)
        np.random.seed(seed)
        g1, p1 = coarsening.metis(M, 3)
        np.random.seed(seed)
        g2, p2 = coarsening_np.metis(M, 3)
        for a, b in zip(p1, p2):
            assert np.array_equal(a, b)
        for a, b in zip(g1, g2):
            assert same_csr(a, b)
        assert coarsening.compute_perm(p1) == coarsening_np.compute_perm(p2)
