"""Multilevel coarsening + binary-tree vertex permutation -- drop-in for the reference's
`gcn/coarsening.py` (`coarsen`, `metis`, `metis_one_level`, `compute_perm`, `perm_adjacency`,
`perm_data`) plus `perm_data_time` from the example scripts.  All index outputs are bit-exact
with the reference (tests/test_coarsening_product.py against golden fixtures).

What is different from the reference is only speed:
  * `metis_one_level` is a native loop (libtgcn_b200 `tgcn_pair_one_level_*`) instead of a
    pure-Python one;
  * `compute_perm` is O(N) per level (stable bucket of children) instead of `np.where` per node
    (O(N^2): 1.5 s at 32 k vertices, hours at 1 M);
  * `perm_data*` are single vectorised gathers.
"""
import ctypes

import numpy as np
import scipy.sparse

from . import _lib


def metis_one_level(rr, cc, vv, rid, weights):
    """Greedy pairing of one level (reference coarsening.py:119-165); rr must be sorted."""
    lib = _lib.load()
    rr = np.ascontiguousarray(rr, dtype=np.int64)
    cc = np.ascontiguousarray(cc, dtype=np.int64)
    rid = np.ascontiguousarray(rid, dtype=np.int64)
    nnz = rr.shape[0]
    n = int(rr[nnz - 1]) + 1
    # numpy >= 2 evaluates vv * (1.0/w + 1.0/w) in the common dtype of vv and weights
    dt = np.result_type(np.asarray(vv).dtype, np.asarray(weights).dtype)
    if dt == np.float32:
        fn, ct = lib.tgcn_pair_one_level_f32, np.float32
    else:
        fn, ct = lib.tgcn_pair_one_level_f64, np.float64
    vv = np.ascontiguousarray(vv, dtype=ct)
    weights = np.ascontiguousarray(weights, dtype=ct)
    if weights.shape[0] < n or rid.shape[0] < n:
        raise ValueError("weights / visiting order shorter than the vertex count")
    cluster = np.zeros(n, np.int32)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = fn(ptr(rr), ptr(cc), ptr(vv), nnz, ptr(rid), ptr(weights), n, ptr(cluster))
    _lib.check(rc, "tgcn_pair_one_level")
    return cluster


def metis(W, levels, rid=None):
    """`levels` rounds of pairing and contraction (reference coarsening.py:34-115).
    Returns (graphs, parents); draws the level-0 visiting order from numpy's global RNG."""
    N = W.shape[0]
    if rid is None:
        rid = np.random.permutation(range(N))
    degree = W.sum(axis=0) - W.diagonal()
    graphs, parents = [W], []
    for _ in range(levels):
        weights = np.array(degree).squeeze()
        rows, cols, vals = scipy.sparse.find(W)
        order = np.argsort(rows)
        rr, cc, vv = rows[order], cols[order], vals[order]
        cluster_id = metis_one_level(rr, cc, vv, rid, weights)
        parents.append(cluster_id)
        Nnew = cluster_id.max() + 1
        W = scipy.sparse.csr_matrix((vv, (cluster_id[rr], cluster_id[cc])), shape=(Nnew, Nnew))
        W.eliminate_zeros()
        graphs.append(W)
        degree = W.sum(axis=0)
        rid = np.argsort(np.array(W.sum(axis=0)).squeeze())
    return graphs, parents


def compute_perm(parents):
    """Per-level vertex orderings such that siblings are adjacent; missing children become fake
    vertices numbered after the real ones (reference coarsening.py:167-214), in O(N) per level."""
    if len(parents) == 0:
        return []
    top = int(np.max(parents[-1])) + 1
    orders = [np.arange(top, dtype=np.int64)]
    for parent in parents[::-1]:
        parent = np.asarray(parent, dtype=np.int64)
        n_child = parent.shape[0]
        cur = orders[-1]
        n_par = int(cur.max()) + 1 if cur.size else 0
        # children of each parent id, ascending (what np.where(parent == i)[0] yields)
        by_parent = np.argsort(parent, kind="stable")
        counts = np.bincount(parent, minlength=n_par)[:n_par] if n_child else np.zeros(n_par, np.int64)
        if counts.size and counts.max() > 2:
            raise AssertionError("a cluster has more than two children")
        starts = np.zeros(n_par + 1, dtype=np.int64)
        starts[1:] = np.cumsum(counts)
        c = counts[cur]                                     # children per node, in visiting order
        fakes_needed = 2 - c
        fake_base = n_child + np.concatenate([[0], np.cumsum(fakes_needed)[:-1]])
        first = np.where(c >= 1, by_parent[np.minimum(starts[cur], max(n_child - 1, 0))] if n_child else 0, fake_base)
        second_real = by_parent[np.minimum(starts[cur] + 1, max(n_child - 1, 0))] if n_child else 0
        second = np.where(c == 2, second_real, np.where(c == 1, fake_base, fake_base + 1))
        layer = np.empty(2 * cur.size, dtype=np.int64)
        layer[0::2] = first
        layer[1::2] = second
        orders.append(layer)
    for i, layer in enumerate(orders):
        M = top * 2 ** i
        if not np.array_equal(np.sort(layer), np.arange(M)):
            raise AssertionError("permutation of level %d does not cover 0..%d" % (i, M - 1))
    return [layer.tolist() for layer in orders[::-1]]


def perm_adjacency(A, indices):
    """Append isolated fake vertices and relabel (reference coarsening.py:242-269). COO out."""
    if indices is None:
        return A
    M = A.shape[0]
    Mnew = len(indices)
    if Mnew < M:
        raise AssertionError("permutation shorter than the graph")
    A = A.tocoo()
    if Mnew > M:
        A = scipy.sparse.vstack([A, scipy.sparse.coo_matrix((Mnew - M, M), dtype=np.float32)])
        A = scipy.sparse.hstack([A, scipy.sparse.coo_matrix((Mnew, Mnew - M), dtype=np.float32)])
    new_label = np.argsort(indices)
    A.row = np.array(new_label)[A.row]
    A.col = np.array(new_label)[A.col]
    return A


def coarsen(A, levels, self_connections=False, verbose=False):
    """Reference coarsening.py:5-31: returns (graphs[0..levels] as CSR, perm of the finest level)."""
    graphs, parents = metis(A, levels)
    perms = compute_perm(parents)
    for i, G in enumerate(graphs):
        M = G.shape[0]
        if not self_connections:
            G = G.tocoo()
            G.setdiag(0)
        if i < levels:
            G = perm_adjacency(G, perms[i])
        G = G.tocsr()
        G.eliminate_zeros()
        graphs[i] = G
        if verbose:
            print('Layer {0}: M_{0} = |V| = {1} nodes ({2} added),|E| = {3} edges'.format(
                i, G.shape[0], G.shape[0] - M, G.nnz // 2))
    return graphs, perms[0] if levels > 0 else None


def _gather_plan(indices, M):
    idx = np.asarray(indices, dtype=np.int64)
    real = idx < M
    return np.where(real, idx, 0), real


def perm_data(x, indices):
    """[Ns, M] -> [Ns, len(indices)] float64, zeros at fake vertices (reference coarsening.py:219-240)."""
    if indices is None:
        return x
    N, M = x.shape
    if len(indices) < M:
        raise AssertionError("permutation shorter than the data")
    src, real = _gather_plan(indices, M)
    out = np.asarray(x, dtype=np.float64)[:, src]
    out[:, ~real] = 0.0
    return out


def perm_data_time(x, indices):
    """[Ns, M, T] -> [Ns, len(indices), T] (reference pytorch_mnist_tgcn.py:18-39, load/data_hcp.py:272-293)."""
    if indices is None:
        return x
    N, M, T = x.shape
    if len(indices) < M:
        raise AssertionError("permutation shorter than the data")
    src, real = _gather_plan(indices, M)
    out = np.asarray(x, dtype=np.float64)[:, src, :]
    out[:, ~real, :] = 0.0
    return out
