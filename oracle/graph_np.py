"""Restatement of the graph-construction inputs of the hot path -- TEST INFRASTRUCTURE ONLY.

Reference: cassianobecker/tgcn ``gcn/graph.py``.  Bit-exact parity with the reference
is required downstream (the coarsening compares float32 edge weights with strict ``>``),
so every floating-point operation keeps the reference's dtype and library call; the
third-party arithmetic (sklearn ``pairwise_distances``, numpy ``argsort``, scipy sparse
sums) is the same de-facto pin as the reference's (it pins no
versions, the installed ones define the results).
"""
import numpy as np
import scipy.sparse as sp
import sklearn.metrics


def grid_embedding(m, dtype=np.float32):
    """m*m points on the unit square, x fastest (graph.py:10-19)."""
    ticks = np.linspace(0, 1, m, dtype=dtype)
    gx, gy = np.meshgrid(ticks, ticks)
    pts = np.empty((m * m, 2), dtype)
    pts[:, 0] = gx.reshape(-1)
    pts[:, 1] = gy.reshape(-1)
    return pts


def knn_exact(z, k=4, metric="euclidean"):
    """Exact k nearest neighbours from the full pairwise matrix (graph.py:33-41).
    Column 0 (self) is dropped; ties are resolved by numpy's default argsort."""
    full = sklearn.metrics.pairwise.pairwise_distances(z, metric=metric, n_jobs=1)
    order = np.argsort(full)[:, 1:k + 1]
    full.sort()
    return full[:, 1:k + 1], order


def knn_adjacency(dist, idx):
    """Gaussian-weighted symmetric kNN adjacency (graph.py:57-83)."""
    M, k = dist.shape
    assert dist.min() >= 0
    sigma2 = np.mean(dist[:, -1]) ** 2                           # graph.py:64
    w = np.exp(-dist ** 2 / sigma2)                              # graph.py:65
    rows = np.arange(0, M).repeat(k)
    W = sp.coo_matrix((w.reshape(M * k), (rows, idx.reshape(M * k))), shape=(M, M))
    W.setdiag(0)                                                 # graph.py:75
    mask = W.T > W                                               # graph.py:77-78: symmetrise by max
    W = W - W.multiply(mask) + W.T.multiply(mask)
    assert W.nnz % 2 == 0
    assert sp.isspmatrix_csr(W)
    return W


def laplacian(W, normalized=True):
    """Combinatorial or symmetric-normalised Laplacian (graph.py:117-136)."""
    deg = W.sum(axis=0)                                          # np.matrix [1,N], dtype of W
    if not normalized:
        return sp.diags(deg.A.squeeze(), 0) - W
    deg += np.spacing(np.array(0, W.dtype))                      # graph.py:128 (isolated vertices)
    deg = 1 / np.sqrt(deg)
    D = sp.diags(deg.A.squeeze(), 0)
    I = sp.identity(deg.size, dtype=W.dtype)
    L = I - D * W * D
    assert sp.isspmatrix_csr(L)
    return L


def rescale_laplacian(L, lmax=2):
    """Map the spectrum to [-1,1]: L <- L/(lmax/2) - I, IN PLACE on the caller's object
    like the reference (graph.py:232-238)."""
    M = L.shape[0]
    I = sp.identity(M, format="csr", dtype=L.dtype)
    L /= lmax / 2
    L -= I
    return L
