"""Synthetic inputs of the BASELINE.json configs: adjacency matrices and signals.  Pure numpy / scipy / torch-CPU, and
independent of the rest of the package (no kernels, no `tgcn_b200.graph` / `.coarsening`): bench.py's reference arm
builds the SAME graphs from these functions with the oracle's (or the reference's own) graph and coarsening code.

  * `hcp_adjacency`  config 2: HCP-shaped parcellation connectome
  * `mesh_adjacency` config 3: spherical triangulation with the fsLR-32k vertex count (load/data_hcp.py:86; edges from
                     the faces with unit weights as load/create_hcp.py:330-361,459-460 does)
  * `rgg_adjacency`  config 4: random geometric graph, strip + Morton vertex order
"""
import math

import numpy as np
import scipy.sparse as sp
import torch


def hcp_adjacency(n_real=360, knn=16, seed=0, dense=False):
    """Lognormal symmetric weights; the sparse variant keeps the top-`knn` entries per row and symmetrises by max."""
    rng = np.random.default_rng(seed)
    M = rng.lognormal(0.0, 1.0, size=(n_real, n_real)).astype(np.float32)
    M = np.maximum(M, M.T)
    np.fill_diagonal(M, 0.0)
    if not dense:
        thresh = np.sort(M, axis=1)[:, -knn][:, None]
        M = np.where(M >= thresh, M, 0.0).astype(np.float32)
        M = np.maximum(M, M.T)
    return sp.csr_matrix(M)


def fibonacci_sphere(n):
    i = np.arange(n, dtype=np.float64) + 0.5
    phi = np.arccos(1.0 - 2.0 * i / n)
    theta = math.pi * (1.0 + 5.0 ** 0.5) * i
    return np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], axis=1)


def mesh_adjacency(n_real=32492):
    """Closed genus-0 triangulated surface: Fibonacci-lattice points on the sphere, spherical Delaunay triangulation
    (= convex hull), unit weights on the face edges."""
    from scipy.spatial import ConvexHull
    pts = fibonacci_sphere(n_real)
    faces = ConvexHull(pts).simplices
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0)
    e = np.concatenate([e, e[:, ::-1]], axis=0)
    A = sp.coo_matrix((np.ones(e.shape[0], np.float32), (e[:, 0], e[:, 1])), shape=(n_real, n_real)).tocsr()
    A.data[:] = 1.0                                            # duplicate edges collapse to weight 1
    return A


def _morton2(ix, iy):
    """Interleave the bits of two uint32 arrays (Z-order code)."""
    def spread(v):
        v = v.astype(np.uint64) & np.uint64(0xFFFFFFFF)
        v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
        return v
    return spread(ix) | (spread(iy) << np.uint64(1))


def rgg_adjacency(n=1_000_000, mean_degree=12.0, seed=0, order="strip-morton", strips=8):
    """Points uniform in the unit square, edges within r = sqrt(mean_degree / (pi n)), Gaussian weights; no coarsening.
    Returns (A, points in vertex order).

    Vertex order ("sorted by x (strip partition) or Morton order"): `strips` vertical strips of equal
    population in x order -- so a contiguous row partition over <= `strips` ranks is a strip partition with thin
    halos -- and Z-order (Morton) inside each strip, so that consecutive rows are spatial neighbours and share most
    of their gathered rows.  order="x" keeps the plain x sort."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 2))
    pts = pts[np.argsort(pts[:, 0], kind="stable")]
    if order == "strip-morton":
        strip = (np.arange(n) * strips) // n                      # equal-population strips of the x-sorted points
        code = _morton2((pts[:, 0] * 65535).astype(np.uint32), (pts[:, 1] * 65535).astype(np.uint32))
        pts = pts[np.lexsort((code, strip))]
    elif order != "x":
        raise ValueError(order)
    r = math.sqrt(mean_degree / (math.pi * n))
    pairs = cKDTree(pts).query_pairs(r, output_type='ndarray')
    d = np.linalg.norm(pts[pairs[:, 0]] - pts[pairs[:, 1]], axis=1)
    w = np.exp(-(d / (0.5 * r)) ** 2).astype(np.float32)
    rows = np.concatenate([pairs[:, 0], pairs[:, 1]])
    cols = np.concatenate([pairs[:, 1], pairs[:, 0]])
    A = sp.coo_matrix((np.concatenate([w, w]), (rows, cols)), shape=(n, n)).tocsr()
    return A, pts


def synthetic_signals(Q, N0, H, n_real, perm, seed, F_in=None):
    """x[Q, N0, H(,F)] ~ N(0,1) on real vertices, exact zeros on the fake vertices the coarsening
    added (what perm_data_time produces)."""
    g = torch.Generator().manual_seed(seed)
    shape = (Q, N0, H) if F_in is None else (Q, N0, H, F_in)
    x = torch.randn(shape, generator=g)
    if perm is not None:
        fake = torch.tensor(np.asarray(perm) >= n_real)
        x[:, fake] = 0.0
    return x
