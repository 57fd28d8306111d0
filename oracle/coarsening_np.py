"""Restatement of the multilevel coarsening + binary-tree permutation -- TEST INFRASTRUCTURE ONLY.

Reference: cassianobecker/tgcn ``gcn/coarsening.py``.  All outputs are integer index
lists (bit-exact parity required); the floating-point comparisons that steer them are kept
in the reference's evaluation order and dtypes (float32 edge weight times float64
reciprocal-degree sum, strict ``>``; coarsening.py:152-153).
"""
import numpy as np
import scipy.sparse as sp


def pair_one_level(rr, cc, vv, visit_order, weights):
    """Greedy heavy-edge matching with the Graclus normalisation (coarsening.py:119-165).

    rr/cc/vv: COO triplets sorted by row.  Vertices are visited in ``visit_order``; an
    unmatched vertex grabs its unmatched neighbour maximising w_ij (1/d_i + 1/d_j), first
    maximum wins, and both get the next cluster id; vertices with no free neighbour become
    singletons."""
    nnz = rr.shape[0]
    n = rr[nnz - 1] + 1
    taken = np.zeros(n, bool)
    first = np.zeros(n, np.int32)
    length = np.zeros(n, np.int32)
    cluster = np.zeros(n, np.int32)

    # Row extents exactly as the reference derives them (coarsening.py:134-139): a
    # *slot* counter that only advances when the row id grows, so a vertex with an
    # empty row shifts every later slot -- the reference then indexes the slots by
    # vertex id.  Restated literally because it decides the pairing on such inputs.
    last = rr[0]
    slot = 0
    for e in range(nnz):
        length[slot] += 1
        if rr[e] > last:
            last = rr[e]
            first[slot + 1] = e
            slot += 1

    n_clusters = 0
    for pos in range(n):
        v = visit_order[pos]
        if taken[v]:
            continue
        taken[v] = True
        best_val = 0.0
        best = -1
        base = first[v]
        for off in range(length[v]):
            u = cc[base + off]
            if taken[u]:
                score = 0.0
            else:
                score = vv[base + off] * (1.0 / weights[v] + 1.0 / weights[u])
            if score > best_val:
                best_val = score
                best = u
        cluster[v] = n_clusters
        if best > -1:
            cluster[best] = n_clusters
            taken[best] = True
        n_clusters += 1
    return cluster


def metis(W, levels, rid=None):
    """``levels`` rounds of pairing + graph contraction (coarsening.py:34-115).
    Returns (graphs[0..levels], parents[0..levels-1])."""
    N = W.shape[0]
    if rid is None:
        rid = np.random.permutation(range(N))                    # coarsening.py:55-56
    degree = W.sum(axis=0) - W.diagonal()                        # coarsening.py:58
    graphs = [W]
    parents = []
    for _ in range(levels):
        weights = np.array(degree).squeeze()
        r, c, v = sp.find(W)                                     # coarsening.py:77
        order = np.argsort(r)                                    # coarsening.py:78 (default quicksort)
        rr, cc, vv = r[order], c[order], v[order]
        cluster = pair_one_level(rr, cc, vv, rid, weights)
        parents.append(cluster)
        n_new = cluster.max() + 1
        # coarse weights: duplicate (row,col) pairs are summed by the CSR constructor
        W = sp.csr_matrix((vv, (cluster[rr], cluster[cc])), shape=(n_new, n_new))  # :98
        W.eliminate_zeros()
        graphs.append(W)
        degree = W.sum(axis=0)                                   # coarsening.py:105 (self loops kept)
        rid = np.argsort(np.array(W.sum(axis=0)).squeeze())      # coarsening.py:112-113
    return graphs, parents


def compute_perm(parents):
    """Per-level orderings that make siblings adjacent, inventing fake ids for missing
    children (coarsening.py:167-214)."""
    orders = []
    if len(parents) > 0:
        orders.append(list(range(max(parents[-1]) + 1)))
    for parent in parents[::-1]:
        next_fake = len(parent)
        layer = []
        for node in orders[-1]:
            kids = list(np.where(parent == node)[0])
            assert 0 <= len(kids) <= 2
            if len(kids) == 1:                                   # singleton: one fake sibling
                kids.append(next_fake)
                next_fake += 1
            elif len(kids) == 0:                                 # fake parent: two fake children
                kids.extend([next_fake, next_fake + 1])
                next_fake += 2
            layer.extend(kids)
        orders.append(layer)
    for i, layer in enumerate(orders):
        assert sorted(layer) == list(range(len(orders[0]) * 2 ** i))
    return orders[::-1]


def perm_adjacency(A, indices):
    """Pad with isolated fake vertices and relabel rows/cols (coarsening.py:242-269)."""
    if indices is None:
        return A
    M = A.shape[0]
    Mnew = len(indices)
    assert Mnew >= M
    A = A.tocoo()
    if Mnew > M:
        A = sp.vstack([A, sp.coo_matrix((Mnew - M, M), dtype=np.float32)])
        A = sp.hstack([A, sp.coo_matrix((Mnew, Mnew - M), dtype=np.float32)])
    rank = np.argsort(indices)
    A.row = np.array(rank)[A.row]
    A.col = np.array(rank)[A.col]
    return A


def coarsen(A, levels, self_connections=False):
    """coarsening.py:5-31.  Returns (graphs, perm of the finest level)."""
    graphs, parents = metis(A, levels)
    perms = compute_perm(parents)
    for i, G in enumerate(graphs):
        if not self_connections:
            G = G.tocoo()
            G.setdiag(0)
        if i < levels:
            G = perm_adjacency(G, perms[i])
        G = G.tocsr()
        G.eliminate_zeros()
        graphs[i] = G
    return graphs, (perms[0] if levels > 0 else None), parents, perms


def perm_data(x, indices):
    """Column gather with zero fill for fake vertices (coarsening.py:219-240); float64 out
    like the reference's ``np.empty`` default."""
    if indices is None:
        return x
    N, M = x.shape
    out = np.zeros((N, len(indices)))
    for i, j in enumerate(indices):
        if j < M:
            out[:, i] = x[:, j]
    return out


def perm_data_time(x, indices):
    """3-D variant used by the TGCN examples
    (examples/pytorch_based/pytorch_mnist_tgcn.py:18-39, load/data_hcp.py:272-293)."""
    if indices is None:
        return x
    N, M, T = x.shape
    out = np.zeros((N, len(indices), T))
    for i, j in enumerate(indices):
        if j < M:
            out[:, i, :] = x[:, j, :]
    return out
