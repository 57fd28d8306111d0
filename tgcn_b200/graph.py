"""Host-side producers of the layer's graph operand -- drop-in for the reference's `gcn/graph.py`
functions that sit on the hot path: `grid`,
`distance_sklearn_metrics`, `adjacency`, `laplacian`, `rescale_L`, `lmax`.

Same names, arguments and results (bit-exact: the coarsening downstream compares float32 edge
weights with strict '>').  Run once per graph on the host, exactly like the reference; the device
work starts at `tgcn_b200.nn`.  Differences from the reference: `rescale_L` does not mutate its
argument unless `inplace=True` is passed (the reference always does), and the unused spectral
helpers (`fourier`, `lanczos`, `plot_spectrum`, LSH kNN) are out of scope.
"""
import numpy as np
import scipy.sparse


def grid(m, dtype=np.float32):
    """[m*m, 2] embedding of the regular grid on the unit square (reference graph.py:10-19)."""
    axis = np.linspace(0, 1, m, dtype=dtype)
    xs, ys = np.meshgrid(axis, axis)
    return np.stack([xs.reshape(m * m), ys.reshape(m * m)], axis=1).astype(dtype, copy=False)


def distance_sklearn_metrics(z, k=4, metric='euclidean'):
    """Exact kNN (distances, indices) from the full pairwise matrix (reference graph.py:33-41)."""
    import sklearn.metrics
    d = sklearn.metrics.pairwise.pairwise_distances(z, metric=metric, n_jobs=1)
    nearest = np.argsort(d)[:, 1:k + 1]
    d.sort()
    return d[:, 1:k + 1], nearest


def adjacency(dist, idx):
    """Symmetric Gaussian-kernel kNN adjacency, CSR (reference graph.py:57-83)."""
    M, k = dist.shape
    if idx.shape != (M, k):
        raise ValueError("dist and idx must have the same [M, k] shape")
    if dist.min() < 0:
        raise ValueError("negative distance")
    sigma2 = np.mean(dist[:, -1]) ** 2
    weights = np.exp(- dist ** 2 / sigma2)
    src = np.arange(0, M).repeat(k)
    W = scipy.sparse.coo_matrix((weights.reshape(M * k), (src, idx.reshape(M * k))), shape=(M, M))
    W.setdiag(0)
    larger_t = W.T > W                       # keep the larger of w_ij, w_ji
    W = W - W.multiply(larger_t) + W.T.multiply(larger_t)
    assert W.nnz % 2 == 0
    return W.tocsr()


def laplacian(W, normalized=True):
    """L = D - W or I - D^-1/2 W D^-1/2, CSR, dtype of W (reference graph.py:117-136)."""
    d = W.sum(axis=0)
    if not normalized:
        return (scipy.sparse.diags(d.A.squeeze(), 0) - W).tocsr()
    d += np.spacing(np.array(0, W.dtype))    # isolated (fake) vertices: avoid 1/0
    d = 1 / np.sqrt(d)
    D = scipy.sparse.diags(d.A.squeeze(), 0)
    I = scipy.sparse.identity(d.size, dtype=W.dtype)
    return (I - D * W * D).tocsr()


def lmax(L, normalized=True):
    """Upper bound of the spectrum (reference graph.py:139-145)."""
    if normalized:
        return 2
    import scipy.sparse.linalg
    return scipy.sparse.linalg.eigsh(L, k=1, which='LM', return_eigenvectors=False)[0]


def rescale_L(L, lmax=2, inplace=False):
    """L / (lmax/2) - I: spectrum into [-1, 1] (reference graph.py:232-238).  Works on scipy sparse
    and on dense `np.matrix`/ndarray inputs like the reference."""
    if not inplace:
        L = L.copy()
    M = L.shape[0]
    I = scipy.sparse.identity(M, format='csr', dtype=L.dtype)
    L /= lmax / 2
    L -= I
    return L


def rescaled_laplacian_csr(A, lmax=2):
    """Convenience: adjacency -> rescaled normalised Laplacian as float32 scipy CSR, the form
    `tgcn_b200.nn` layers ingest without the reference's `.todense()` detour."""
    L = rescale_L(laplacian(A, normalized=True), lmax=lmax, inplace=True).tocsr()
    L.eliminate_zeros()
    return L.astype(np.float32)
