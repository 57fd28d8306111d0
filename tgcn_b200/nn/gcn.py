"""Drop-in replacements for the reference's `tgcn/nn/gcn.py` layers.

Same class names, constructor signatures, parameter names / shapes / initialisation (including
the order of draws from the global torch RNG), `forward(x)` signature, `__repr__`, `state_dict`
keys (`weight`, `bias`) -- see SURVEY.md section 8b.  The compute runs in the sm_100a CUDA
kernels behind include/tgcn_b200.h; inputs must be CUDA fp32 tensors (no CPU fallback).

Recursion semantics: by default the layers reproduce what the reference actually computes
(`recursion="reference"`: Xt_k = 2 L^k X - Xt_{k-2}, see oracle/layers_np.py); the textbook
recursion T_k = 2 L T_{k-1} - T_{k-2} is available as `recursion="chebyshev"` (keyword-only).
"""
import math

import torch
from torch.nn import Parameter

from .. import _lib
from ..csr import CSRCache
from . import functional as F_

_RECURSION = {"reference": _lib.RECURSION_REFERENCE, "chebyshev": _lib.RECURSION_CHEBYSHEV}
_ENGINE = {"auto": _lib.ENGINE_AUTO, "ffma": _lib.ENGINE_FFMA, "tcgen05": _lib.ENGINE_TCGEN05,
           "resident": _lib.ENGINE_RESIDENT}


def dft_real_matrix(H):
    """C[h, k] = cos(2 pi h k / H): real part of the length-H DFT (`npa.real(npa.fft.fft(x, axis=2))`,
    pytorch_mnist_tgcn.py:87), computed in float64 and rounded once to fp32."""
    h = torch.arange(H, dtype=torch.float64)
    return torch.cos(2.0 * math.pi * torch.outer(h, h) / H).to(torch.float32)


def uniform(size, tensor):
    """U(-1/sqrt(size), 1/sqrt(size)) in place; no-op for None (reference gcn.py:240-243)."""
    bound = 1.0 / math.sqrt(size)
    if tensor is not None:
        tensor.data.uniform_(-bound, bound)


class _ChebBase(torch.nn.Module):
    _bias_mode = _lib.BIAS_PER_VERTEX

    def _setup(self, L, in_channels, out_channels, filter_order, weight_shape, bias_shape, bias, recursion, engine):
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.weight = Parameter(torch.Tensor(*weight_shape))
        self.L = L                      # plain attribute, exactly like the reference (not in state_dict)
        self.filter_order = filter_order
        if bias:
            self.bias = Parameter(torch.Tensor(*bias_shape))
        else:
            self.register_parameter('bias', None)
        if recursion not in _RECURSION:
            raise ValueError("recursion must be 'reference' or 'chebyshev'")
        if engine not in _ENGINE:
            raise ValueError("engine must be one of %s" % sorted(_ENGINE))
        self.recursion = recursion
        self.engine = engine
        self._csr = CSRCache()
        self.reset_parameters()

    def reset_parameters(self):
        size = self.in_channels * self.weight.size(0)
        uniform(size, self.weight)      # weight first, then bias: same RNG draw order as the reference
        uniform(size, self.bias)

    def __repr__(self):
        return '{}({}, {}, filter_order={})'.format(self.__class__.__name__, self.in_channels, self.out_channels,
                                                    self.weight.size(0))

    # -- helpers ---------------------------------------------------------------------------------
    def _plan(self, device):
        return self._csr.get(self.L, device)

    def _canon(self, x):
        raise NotImplementedError

    def _use_resident(self, plan, D, G, K):
        """engine="auto": the sample-resident fused kernels whenever one sample's slab fits in shared
        memory, else the streaming path; engine="resident" insists (error if it does not fit)."""
        if self.engine not in ("auto", "resident"):
            return False
        ok = F_.resident_supported(plan, D, G, K)
        if self.engine == "resident" and not ok:
            raise RuntimeError("engine='resident': N=%d D=%d G=%d K=%d nnz=%d does not fit the resident kernels"
                               % (plan.n, D, G, K, plan.nnz))
        return ok

    def _effective_weight(self):
        return self.weight

    def _run(self, x3, pool_p=0, relu=False):
        F_._require_cuda_f32(x3, "x")
        w = self._effective_weight()
        F_._require_cuda_f32(w, "weight")
        K = w.shape[0]
        w3 = w.reshape(K, -1, w.shape[-1])
        plan = self._plan(x3.device)
        rec = _RECURSION[self.recursion]
        if self._use_resident(plan, w3.shape[1], w3.shape[2], K):
            res = F_.ResidentChebFunction.apply(x3, w3, self.bias, plan, self._bias_mode, rec, pool_p, relu)
            return res[0] if pool_p else res
        out = F_.ChebLayerFunction.apply(x3, w3, self.bias, plan, self._bias_mode, rec, _ENGINE[self.engine])
        if pool_p:
            out = F_.PoolFunction.apply(out, pool_p, relu)[0]
        return out

    def forward(self, x):
        x3 = self._canon(x)
        return self._run(x3)

    def forward_relu_pool(self, x, p):
        """gcn_pool / gcn_pool_4 (p = 2 / 4) of F.relu(self(x)) as one fused operation
        (pytorch_hcp_tgcn.py:133-141 applies exactly this chain after each conv layer): identical
        values, pool indices and gradients, without writing the un-pooled activation."""
        if p not in (2, 4):
            raise ValueError("pool size must be 2 (gcn_pool) or 4 (gcn_pool_4)")
        return self._run(self._canon(x), pool_p=p, relu=True)

    def _basis(self, x):
        x3 = self._canon(x)
        F_._require_cuda_f32(x3, "x")
        return F_.cheb_basis(x3, self._plan(x3.device), self.filter_order, _RECURSION[self.recursion])


class TGCNCheb_H(_ChebBase):
    """Time-vertex layer: x [Q,N,H] or [Q,N,H,F] -> [Q,N,G]; weight [K,H,F,G], bias [1,N,G]
    (reference gcn.py:82-154)."""

    def __init__(self, L, in_channels, out_channels, filter_order, horizon, bias=True, *, recursion="reference",
                 engine="auto", time_dft=False):
        super(TGCNCheb_H, self).__init__()
        self._setup(L, in_channels, out_channels, filter_order,
                    (filter_order, horizon, in_channels, out_channels), (1, L[0].shape[0], out_channels), bias,
                    recursion, engine)
        # time_dft=True folds the models' prologue `x = real(fft(x, axis=2))` (pytorch_mnist_tgcn.py:87,
        # pytorch_hcp_tgcn.py:133) into the weights: the real DFT is linear along h and commutes with L~, so
        # layer(real(fft(x))) == layer_with(W_eff)(x), W_eff[k,h] = sum_h' cos(2 pi h h'/H) W[k,h'].
        self.time_dft = bool(time_dft)
        self._dft = None

    def _effective_weight(self):
        if not self.time_dft:
            return self.weight
        H = self.weight.shape[1]
        if self._dft is None or self._dft.device != self.weight.device:
            self._dft = dft_real_matrix(H).to(self.weight.device)
        return torch.einsum("hk,jkfg->jhfg", self._dft, self.weight)

    def _canon(self, x):
        if x.dim() == 3:
            x = x.unsqueeze(3)
        if x.dim() != 4:
            raise RuntimeError("TGCNCheb_H expects [Q,N,H] or [Q,N,H,F], got %s" % (tuple(x.shape),))
        K, H, F, G = self.weight.shape
        if x.shape[2] != H or x.shape[3] != F:
            raise RuntimeError("einsum(): operands do not broadcast: input [.,.,%d,%d] vs weight [%d,%d,%d,%d]"
                               % (x.shape[2], x.shape[3], K, H, F, G))
        return x.reshape(x.shape[0], x.shape[1], H * F)

    def _time_chebyshev(self, X):
        """Xt[K,Q,N,H,F] exactly as the reference method of the same name returns it."""
        if X.dim() == 3:
            X = X.unsqueeze(3)
        Q, N, H, F = X.shape
        return self._basis(X).reshape(self.filter_order, Q, N, H, F)


class TGCNCheb(_ChebBase):
    """x [Q,N,F] -> [Q,N,G]; weight [K,F,G], bias [1,N,G] (reference gcn.py:8-79)."""

    def __init__(self, L, in_channels, out_channels, filter_order, bias=True, *, recursion="reference", engine="auto"):
        super(TGCNCheb, self).__init__()
        self._setup(L, in_channels, out_channels, filter_order, (filter_order, in_channels, out_channels),
                    (1, L[0].shape[0], out_channels), bias, recursion, engine)

    def _canon(self, x):
        if x.dim() != 3:
            raise RuntimeError("TGCNCheb expects [Q,N,F], got %s" % (tuple(x.shape),))
        if x.shape[2] != self.weight.shape[1]:
            raise RuntimeError("einsum(): operands do not broadcast: input F=%d vs weight F=%d"
                               % (x.shape[2], self.weight.shape[1]))
        return x

    def _time_chebyshev(self, X):
        return self._basis(X)


class GCNCheb(_ChebBase):
    """Spatial layer: x [Q,N] or [Q,N,F] -> [Q,N,G]; weight [K,F,G], bias [1,1,G] (reference gcn.py:158-237)."""
    _bias_mode = _lib.BIAS_PER_FILTER

    def __init__(self, L, in_channels, out_channels, filter_order, bias=True, *, recursion="reference", engine="auto"):
        super(GCNCheb, self).__init__()
        self._setup(L, in_channels, out_channels, filter_order, (filter_order, in_channels, out_channels),
                    (1, 1, out_channels), bias, recursion, engine)

    def _canon(self, x):
        if x.dim() == 2:
            x = x.unsqueeze(2)
        if x.dim() != 3:
            raise RuntimeError("GCNCheb expects [Q,N] or [Q,N,F], got %s" % (tuple(x.shape),))
        if x.shape[2] != self.weight.shape[1]:
            raise RuntimeError("einsum(): operands do not broadcast: input F=%d vs weight F=%d"
                               % (x.shape[2], self.weight.shape[1]))
        return x

    def _chebyshev(self, X):
        return self._basis(X)


def gcn_pool(x):
    """Max over sibling pairs (reference gcn.py:246-249)."""
    return F_.PoolFunction.apply(x, 2, False)[0]


def gcn_pool_4(x):
    """Max over sibling quadruples (reference gcn.py:252-255)."""
    return F_.PoolFunction.apply(x, 4, False)[0]


def gcn_pool_with_indices(x, p):
    """(values, int64 argmax in 0..p-1) -- the pair torch.max(dim=2) yields inside the reference."""
    y, idx = F_.PoolFunction.apply(x, p, False)
    return y, idx.to(torch.int64)


def relu_pool(x, p):
    """F.relu followed by gcn_pool / gcn_pool_4 in one kernel (pytorch_hcp_tgcn.py:135-137 without
    the dropout); identical values, indices and gradients to the unfused pair."""
    return F_.PoolFunction.apply(x, p, True)[0]


# ------------------------------------------------------------------------------------------------
# Edge-index operator family (reference gcn.py:348-538): ChebConv(in, out, K) / ChebTimeConv(in, out, K, H)
# with forward(x, edge_index, edge_weight=None).  The reference rebuilds -D^-1/2 A D^-1/2 from the edge list
# on every call and applies it with gather + scatter_add (`spmm_batch_2/3`, gcn.py:281-345); here the edge
# list is converted ONCE per (edge_index, edge_weight) into the CSR operand of the kernels above and cached.
# ------------------------------------------------------------------------------------------------
def laplacian_from_edges(edge_index, edge_weight, num_nodes, dtype=torch.float32):
    """(row, col, lap) of the rescaled Laplacian exactly as gcn.py:383-398 / :497-512 builds it:
    self-loops removed (torch_geometric.utils.remove_self_loops), deg = number of remaining edges leaving
    each vertex (torch_geometric.utils.degree counts index occurrences; it is NOT weight-aware),
    lap = -deg[row]^-1/2 * w * deg[col]^-1/2 with 1/sqrt(0) := 0."""
    row, col = edge_index[0], edge_index[1]
    keep = row != col
    row, col = row[keep], col[keep]
    if edge_weight is None:
        w = torch.ones(row.shape[0], dtype=dtype, device=row.device)
    else:
        w = edge_weight[keep].reshape(-1).to(dtype)
    deg = torch.zeros(num_nodes, dtype=dtype, device=row.device).scatter_add_(0, row, torch.ones_like(w))
    dis = deg.pow(-0.5)
    dis[dis == float("inf")] = 0
    return row, col, -dis[row] * w * dis[col]


class _EdgeChebBase(_ChebBase):
    _bias_mode = _lib.BIAS_PER_FILTER

    def _edge_setup(self, in_channels, out_channels, weight_shape, bias, engine):
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.weight = Parameter(torch.Tensor(*weight_shape))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        if engine not in _ENGINE:
            raise ValueError("engine must be one of %s" % sorted(_ENGINE))
        self.recursion = "chebyshev"          # gcn.py:404-414: Tx_2 = 2 * L Tx_1 - Tx_0 with the roles rotated
        self.engine = engine
        self.filter_order = weight_shape[0]
        self.L = None
        self._edge_cache = {}
        self.reset_parameters()

    def __repr__(self):
        return '{}({}, {}, K={})'.format(self.__class__.__name__, self.in_channels, self.out_channels, self.weight.size(0))

    def _edge_plan(self, edge_index, edge_weight, num_nodes, device):
        key = (edge_index.data_ptr(), edge_index._version, tuple(edge_index.shape),
               None if edge_weight is None else (edge_weight.data_ptr(), edge_weight._version), num_nodes, str(device))
        plan = self._edge_cache.get(key)
        if plan is None:
            import scipy.sparse as sp
            from ..csr import build_csr
            row, col, lap = laplacian_from_edges(edge_index, edge_weight, num_nodes)
            m = sp.coo_matrix((lap.detach().cpu().numpy(), (row.cpu().numpy(), col.cpu().numpy())),
                              shape=(num_nodes, num_nodes)).tocsr()      # duplicate edges add up, like scatter_add
            plan = build_csr(m, device)
            if len(self._edge_cache) >= 8:
                self._edge_cache.pop(next(iter(self._edge_cache)))
            self._edge_cache[key] = plan
        return plan

    def _run_edges(self, x3, edge_index, edge_weight):
        F_._require_cuda_f32(x3, "x")
        w = self.weight
        K = w.shape[0]
        w3 = w.reshape(K, -1, w.shape[-1])
        plan = self._edge_plan(edge_index, edge_weight, x3.shape[1], x3.device)
        bias = None if self.bias is None else self.bias.reshape(1, 1, -1)
        rec = _RECURSION["chebyshev"]
        if self._use_resident(plan, w3.shape[1], w3.shape[2], K):
            out = F_.ResidentChebFunction.apply(x3, w3, bias, plan, self._bias_mode, rec, 0, False)
        else:
            out = F_.ChebLayerFunction.apply(x3, w3, bias, plan, self._bias_mode, rec, _ENGINE[self.engine])
        return out


class ChebConv(_EdgeChebBase):
    """x [Q,N] or [Q,N,F], edge_index [2,E], edge_weight [E] or None -> [Q,N,G]; weight [K,F,G], bias [G]
    (reference gcn.py:348-429)."""

    def __init__(self, in_channels, out_channels, K, bias=True, *, engine="auto"):
        torch.nn.Module.__init__(self)
        self._edge_setup(in_channels, out_channels, (K, in_channels, out_channels), bias, engine)

    def forward(self, x, edge_index, edge_weight=None):
        if x.dim() < 3:
            x = x.unsqueeze(-1)
        if x.dim() != 3 or x.shape[2] != self.weight.shape[1]:
            raise RuntimeError("ChebConv expects [Q,N] or [Q,N,%d], got %s" % (self.weight.shape[1], tuple(x.shape)))
        return self._run_edges(x, edge_index, edge_weight)


class ChebTimeConv(_EdgeChebBase):
    """x [Q,N,H] or [Q,N,H,F] -> [Q,N,G]; weight [K,H,F,G], bias [G] (reference gcn.py:432-538)."""

    def __init__(self, in_channels, out_channels, K, H, bias=True, *, engine="auto"):
        torch.nn.Module.__init__(self)
        self._edge_setup(in_channels, out_channels, (K, H, in_channels, out_channels), bias, engine)

    def forward(self, x, edge_index, edge_weight=None):
        if x.dim() < 4:
            x = x.unsqueeze(-1)
        K, H, F, G = self.weight.shape
        if x.dim() != 4 or x.shape[2] != H or x.shape[3] != F:
            raise RuntimeError("ChebTimeConv expects [Q,N,%d] or [Q,N,%d,%d], got %s" % (H, H, F, tuple(x.shape)))
        return self._run_edges(x.reshape(x.shape[0], x.shape[1], H * F), edge_index, edge_weight)


def perm_data_time(x, indices):
    """Device-side `perm_data_time` (pytorch_mnist_tgcn.py:18-39, load/data_hcp.py:272-293): x [Ns,M,T] ->
    [Ns,len(indices),T], vertex i takes x[:, indices[i]] when indices[i] < M and zeros otherwise (the fake
    vertices coarsening adds).  One gather on the tensor's device; dtype preserved (the reference's NumPy
    version returns float64)."""
    if indices is None:
        return x
    Ns, M, T = x.shape
    idx = torch.as_tensor(indices, dtype=torch.long, device=x.device)
    if idx.numel() < M:
        raise AssertionError("permutation shorter than the data")
    real = idx < M
    out = x.index_select(1, torch.where(real, idx, torch.zeros_like(idx)))
    return out * real.to(x.dtype).view(1, -1, 1)
