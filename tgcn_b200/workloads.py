"""Synthetic graphs and the reference's model compositions for the BASELINE.json configs.

Graph builders run on the host with `tgcn_b200.graph` / `tgcn_b200.coarsening` (the drop-in
producers); everything is seeded.  The models mirror the reference's experiment scripts:

  * `NetTGCN_HCP`   examples/pytorch_based/pytorch_hcp_tgcn.py:93-155
  * `NetTGCN_MNIST` examples/pytorch_based/pytorch_mnist_tgcn.py:67-92 (+ tgcn_mnist.py hyper-parameters)

The conv layers, pooling, dropouts and the dense classifier head all run in tgcn_b200's CUDA kernels
(nn/gcn.py, nn/head.py); plain torch modules own the head's parameters.
"""
import math

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import coarsening, graph, synth
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
    """Config 2: HCP-shaped parcellation connectome (synth.hcp_adjacency), coarsened."""
    return _coarsened_laplacians(synth.hcp_adjacency(n_real, knn, seed, dense), levels, seed) + (n_real,)


def cortical_mesh(n_real=32492, levels=4, seed=0):
    """Config 3: closed genus-0 triangulated surface with the fsLR-32k vertex count (synth.mesh_adjacency), coarsened."""
    return _coarsened_laplacians(synth.mesh_adjacency(n_real), levels, seed) + (n_real,)


def random_geometric(n=1_000_000, mean_degree=12.0, seed=0, order="strip-morton", strips=8):
    """Config 4: random geometric graph (synth.rgg_adjacency); returns L~ (CSR) and the points in vertex order."""
    A, pts = synth.rgg_adjacency(n, mean_degree, seed, order, strips)
    return graph.rescaled_laplacian_csr(A), pts


synthetic_signals = synth.synthetic_signals


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
