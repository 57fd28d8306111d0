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


def _num_vertices(L):
    """N of the operand.  The reference writes `L[0].shape[0]` (gcn.py:22,96), which is N for the dense / torch-sparse
    tensors its examples pass; for a scipy CSR matrix (also accepted here) `L[0]` is a 1 x N row, so N is taken from
    the matrix shape itself."""
    shape = getattr(L, "shape", None)
    if shape is not None and len(shape) == 2:
        if shape[0] != shape[1]:
            raise ValueError("L must be square, got %s" % (tuple(shape),))
        return int(shape[0])
    return int(L[0].shape[0])


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
        self._csr = CSRCache()
        self.L = L                      # attribute like the reference's (not in state_dict); see the `L` property
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
        self.reset_parameters()

    # The reference re-reads `self.L` on every forward (gcn.py:141,223); here the operand is converted once per device
    # and cached, so assigning a new `layer.L` drops the cached CSR plans.
    @property
    def L(self):
        return self.__dict__.get("_L")

    @L.setter
    def L(self, value):
        self.__dict__["_L"] = value
        cache = self.__dict__.get("_csr")
        if cache is not None:
            cache.clear()

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

    # -- fused dropout stream (the nn.Dropout between ReLU and pool of the reference models) ------------------------
    dropout_step = None      # optional external device uint32 counter advanced once per training step by the caller
                             # (e.g. PeerAllreduceSGD.state): saves the layer's own counter bump launch

    def _drop_spec(self, p, device):
        """(p, seed, step tensor) for the kernels.  The mask of a step is a hash of (seed, step counter, element): the
        seed derives from torch.initial_seed() and the layer's shape (no draw from the global RNG); the counter is the
        caller's `dropout_step` tensor, or a per-layer device counter bumped on every dropout forward."""
        if not p:
            return None
        seed = (torch.initial_seed() * 0x9E3779B1 + self.weight.numel() * 0x85EBCA77 + self.out_channels) & 0xFFFFFFFF
        step = self.dropout_step
        if step is None:
            ctr = self.__dict__.get("_drop_ctr")
            if ctr is None or ctr.device != device:
                ctr = torch.zeros(1, dtype=torch.int32, device=device)
                self.__dict__["_drop_ctr"] = ctr
            ctr.add_(1)
            step = ctr
        return (float(p), seed, step)

    def _run(self, x3, pool_p=0, relu=False, dropout=0.0):
        F_._require_cuda_f32(x3, "x")
        w = self._effective_weight()
        F_._require_cuda_f32(w, "weight")
        K = w.shape[0]
        w3 = w.reshape(K, -1, w.shape[-1])
        plan = self._plan(x3.device)
        rec = _RECURSION[self.recursion]
        if self.bias is not None:
            # the kernels index the bias by (vertex, filter) or (filter): a wrong size would be an out-of-bounds access
            want = plan.n * w3.shape[2] if self._bias_mode == _lib.BIAS_PER_VERTEX else w3.shape[2]
            if self.bias.numel() != want:
                raise RuntimeError("bias has %d elements, expected %d (%s, N=%d, G=%d)"
                                   % (self.bias.numel(), want,
                                      "per-vertex" if self._bias_mode == _lib.BIAS_PER_VERTEX else "per-filter",
                                      plan.n, w3.shape[2]))
        drop = self._drop_spec(dropout, x3.device) if (pool_p and relu) else None
        if self._use_resident(plan, w3.shape[1], w3.shape[2], K):
            res = F_.ResidentChebFunction.apply(x3, w3, self.bias, plan, self._bias_mode, rec, pool_p, relu, drop)
            return res[0] if pool_p else res
        out = F_.ChebLayerFunction.apply(x3, w3, self.bias, plan, self._bias_mode, rec, _ENGINE[self.engine])
        if pool_p:
            out = F_.PoolFunction.apply(out, pool_p, relu, drop)[0]
        return out

    def forward(self, x):
        x3 = self._canon(x)
        return self._run(x3)

    def forward_relu_pool(self, x, p, dropout=0.0):
        """gcn_pool / gcn_pool_4 (p = 2 / 4) of dropout(F.relu(self(x))) as one fused operation
        (pytorch_hcp_tgcn.py:134-141 applies exactly this chain after each conv layer; `dropout` is the rate of the
        nn.Dropout between the ReLU and the pool, 0 in evaluation mode): identical values, pool indices and
        gradients to the unfused chain with the same mask, without writing the un-pooled activation."""
        if p not in (2, 4):
            raise ValueError("pool size must be 2 (gcn_pool) or 4 (gcn_pool_4)")
        return self._run(self._canon(x), pool_p=p, relu=True, dropout=dropout)

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
                    (filter_order, horizon, in_channels, out_channels), (1, _num_vertices(L), out_channels), bias,
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
                    (1, _num_vertices(L), out_channels), bias, recursion, engine)

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


def relu_pool(x, p, drop=None):
    """F.relu (-> dropout) -> gcn_pool / gcn_pool_4 in one kernel (pytorch_hcp_tgcn.py:135-137); identical values,
    indices and gradients to the unfused chain.  drop: None or (rate, seed, device step counter or None)."""
    return F_.PoolFunction.apply(x, p, True, drop)[0]


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
        self._edge_cache_list = []
        self.reset_parameters()

    def __repr__(self):
        return '{}({}, {}, K={})'.format(self.__class__.__name__, self.in_channels, self.out_channels, self.weight.size(0))

    def _edge_plan(self, edge_index, edge_weight, num_nodes, device):
        """CSR operand of this edge list, converted once and cached.  A hit requires the SAME tensor objects
        (identity, not address: a freed per-batch `edge_index` can be re-allocated at the same address) at the same
        in-place version; the cache holds references to them, so their storage cannot be recycled while cached."""
        if edge_weight is not None and edge_weight.requires_grad:
            raise NotImplementedError("edge_weight.requires_grad: the Laplacian is built once per edge list and is not "
                                      "differentiated (the reference back-propagates through it, gcn.py:383-398); "
                                      "detach the weights or build L~ with laplacian_from_edges yourself")
        for ent in self._edge_cache_list:
            ei, ev, ew, wv, n, dev, plan = ent
            if ei is edge_index and ev == edge_index._version and ew is edge_weight and \
                    (ew is None or wv == edge_weight._version) and n == num_nodes and dev == device:
                return plan
        from ..csr import build_csr_from_coo
        row, col, lap = laplacian_from_edges(edge_index, edge_weight, num_nodes)
        plan = build_csr_from_coo(row, col, lap.detach(), num_nodes, device)   # duplicate edges add up, like scatter_add
        if len(self._edge_cache_list) >= 8:
            self._edge_cache_list.pop(0)
        self._edge_cache_list.append((edge_index, edge_index._version, edge_weight,
                                      None if edge_weight is None else edge_weight._version, num_nodes, device, plan))
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
