"""Synthetic graphs and the reference's model compositions for the BASELINE.json configs.

Graph builders run on the host with `tgcn_b200.graph` / `tgcn_b200.coarsening` (the drop-in
producers); everything is seeded.  The models mirror the reference's experiment scripts:

  * `NetTGCN_HCP`   examples/pytorch_based/pytorch_hcp_tgcn.py:93-155
  * `NetTGCN_MNIST` examples/pytorch_based/pytorch_mnist_tgcn.py:67-92 (+ tgcn_mnist.py hyper-parameters)

The conv layers and pooling are tgcn_b200's CUDA path; the dense classifier head
(Linear/BatchNorm/log_softmax) is plain torch -- SURVEY.md section 8f ranks it "next", outside
the hot path.
"""
import math

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import coarsening, graph
from .nn.gcn import GCNCheb, TGCNCheb_H, gcn_pool, gcn_pool_4, relu_pool


# ------------------------------------------------------------------------------------------------
# graphs
# ------------------------------------------------------------------------------------------------
def _coarsened_laplacians(A, levels, seed):
    np.random.seed(seed)                       # coarsening.metis draws its visiting order here
    graphs, perm = coarsening.coarsen(A, levels=levels, self_connections=False)
    Ls = [graph.rescaled_laplacian_csr(g) for g in graphs]
    return graphs, perm, Ls


def mnist_grid(k=8, levels=4, seed=0):
    """Config 1: 28x28 k-NN grid graph (seed 0 -> N = [992, 496, 248, 124, 62])."""
    z = graph.grid(28)
    dist, idx = graph.distance_sklearn_metrics(z, k=k, metric='euclidean')
    A = graph.adjacency(dist, idx)
    return _coarsened_laplacians(A, levels, seed) + (784,)


def hcp_parcellation(n_real=360, knn=16, levels=4, seed=0, dense=False):
    """Config 2: HCP-shaped parcellation connectome, lognormal symmetric weights; sparse variant
    keeps the top-`knn` entries per row and symmetrises by max."""
    rng = np.random.default_rng(seed)
    M = rng.lognormal(0.0, 1.0, size=(n_real, n_real)).astype(np.float32)
    M = np.maximum(M, M.T)
    np.fill_diagonal(M, 0.0)
    if not dense:
        thresh = np.sort(M, axis=1)[:, -knn][:, None]
        M = np.where(M >= thresh, M, 0.0).astype(np.float32)
        M = np.maximum(M, M.T)
    A = sp.csr_matrix(M)
    return _coarsened_laplacians(A, levels, seed) + (n_real,)


def fibonacci_sphere(n):
    i = np.arange(n, dtype=np.float64) + 0.5
    phi = np.arccos(1.0 - 2.0 * i / n)
    theta = math.pi * (1.0 + 5.0 ** 0.5) * i
    return np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], axis=1)


def cortical_mesh(n_real=32492, levels=4, seed=0):
    """Config 3: closed genus-0 triangulated surface with the fsLR-32k vertex count
    (load/data_hcp.py:86), unit edge weights from the faces (load/create_hcp.py:330-361,459-460)."""
    from scipy.spatial import ConvexHull
    pts = fibonacci_sphere(n_real)
    faces = ConvexHull(pts).simplices
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], axis=0)
    e = np.concatenate([e, e[:, ::-1]], axis=0)
    A = sp.coo_matrix((np.ones(e.shape[0], np.float32), (e[:, 0], e[:, 1])), shape=(n_real, n_real)).tocsr()
    A.data[:] = 1.0                                            # duplicate edges collapse to weight 1
    return _coarsened_laplacians(A, levels, seed) + (n_real,)


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


def random_geometric(n=1_000_000, mean_degree=12.0, seed=0, order="strip-morton", strips=8):
    """Config 4: points uniform in the unit square, edges within r = sqrt(mean_degree / (pi n)), Gaussian weights;
    no coarsening.  Returns L~ (CSR) and the points in vertex order.

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
    return graph.rescaled_laplacian_csr(A), pts


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


# ------------------------------------------------------------------------------------------------
# models
# ------------------------------------------------------------------------------------------------
class NetTGCN_HCP(nn.Module):
    """real(fft_t(x)) -> TGCNCheb_H(L0,1,32,K,H) -> ReLU -> Dropout(0.1) -> pool4 -> GCNCheb(L2,32,64,K) -> ReLU -> pool4
    -> fc(200) -> BN -> ReLU -> Dropout(0.5) -> fc(n_classes) -> log_softmax        (pytorch_hcp_tgcn.py:93-155).

    `time_dft`: the model's prologue `torch.rfft(x, 1, onesided=False)[..., 0]` (:133, the real part of the DFT along the
    time window) is folded into tgcn1's weights (TGCNCheb_H(time_dft=True)) instead of transforming every batch.
    The dropouts are fused into the pool / head kernels (counter-based masks); `set_dropout_step(t)` hands every fused
    dropout the device counter the optimizer advances once per step (else each owner bumps its own counter).
    `fc1_update`: an `nn.head.Fc1FusedSGD` -- fc1.weight's optimizer step then happens inside the backward."""

    def __init__(self, L, horizon=15, K=10, g1=32, g2=64, hidden=200, n_classes=6, fused_relu_pool=True, fused_head=True,
                 time_dft=True, drop1=0.1, drop2=0.5, **layer_kw):
        super().__init__()
        self.tgcn1 = TGCNCheb_H(L[0], 1, g1, K, horizon, time_dft=time_dft, **layer_kw)
        self.drop1 = nn.Dropout(drop1)
        self.gcn2 = GCNCheb(L[2], g1, g2, K, **layer_kw)
        n2 = L[2].shape[0]
        self.fc1 = nn.Linear(int(n2 * g2 / 4), hidden)
        self.dense1_bn = nn.BatchNorm1d(hidden)
        self.drop2 = nn.Dropout(drop2)
        self.fc2 = nn.Linear(hidden, n_classes)
        self.fused = fused_relu_pool
        self.fused_head = fused_head
        self.time_dft = time_dft
        self.fc1_update = None
        self._head_step = None
        self._head_ctr = None

    def set_dropout_step(self, step):
        """step: device int32/uint32 tensor advanced once per training step (e.g. PeerAllreduceSGD.state)."""
        self.tgcn1.dropout_step = step
        self._head_step = step

    def _head_drop(self, device):
        p = self.drop2.p if self.training else 0.0
        if not p:
            return None
        step = self._head_step
        if step is None:
            if self._head_ctr is None or self._head_ctr.device != device:
                self._head_ctr = torch.zeros(1, dtype=torch.int32, device=device)
            self._head_ctr.add_(1)
            step = self._head_ctr
        return (float(p), (torch.initial_seed() * 0x9E3779B1 + 0x5bd1e995) & 0xFFFFFFFF, step)

    def forward(self, x):
        p1 = self.drop1.p if self.training else 0.0
        if self.fused:
            x = self.gcn2.forward_relu_pool(self.tgcn1.forward_relu_pool(x, 4, dropout=p1), 4)
        else:
            x = gcn_pool_4(self.drop1(F.relu(self.tgcn1(x))))
            x = gcn_pool_4(F.relu(self.gcn2(x)))
        x = x.reshape(x.shape[0], -1)
        if self.fused_head and x.is_cuda:
            from .nn.head import fused_head, head_covers
            if head_covers(x.shape[0], self.fc1.in_features, self.fc1.out_features, self.training, torch.is_grad_enabled()):
                upd = self.fc1_update if (self.training and torch.is_grad_enabled()) else None
                return fused_head(x, self.fc1, self.dense1_bn, self.fc2, drop=self._head_drop(x.device), fc1_update=upd)
        if self.fc1_update is not None and self.training:
            raise RuntimeError("fc1_update is set but the fused head does not cover this batch / mode")
        x = self.drop2(F.relu(self.dense1_bn(self.fc1(x))))
        return F.log_softmax(self.fc2(x), dim=1)


class NetTGCN_MNIST(nn.Module):
    """TGCNCheb_H(L0,1,15,K=10,H=12) -> ReLU -> fc(10) -> log_softmax (pytorch_mnist_tgcn.py:67-92)."""

    def __init__(self, L, horizon=12, K=10, g1=15, n_classes=10, **layer_kw):
        super().__init__()
        self.tgcn1 = TGCNCheb_H(L[0], 1, g1, K, horizon, **layer_kw)
        self.fc1 = nn.Linear(L[0].shape[0] * g1, n_classes)

    def forward(self, x):
        x = F.relu(self.tgcn1(x))
        return F.log_softmax(self.fc1(x.reshape(x.shape[0], -1)), dim=1)


def as_torch_operands(Ls, device=None, dense=False):
    """Laplacians in a form with `.shape` and `L[0].shape[0]` (what the reference constructors
    index): torch sparse COO by default, dense like the example scripts on request."""
    out = []
    for L in Ls:
        coo = L.tocoo()
        t = torch.sparse_coo_tensor(np.vstack([coo.row, coo.col]), coo.data.astype(np.float32), coo.shape).coalesce()
        if dense:
            t = t.to_dense()
        out.append(t if device is None else t.to(device))
    return out
