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

    def _run(self, x3, pool_p=0, relu=False):
        F_._require_cuda_f32(x3, "x")
        w = self.weight
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
                 engine="auto"):
        super(TGCNCheb_H, self).__init__()
        self._setup(L, in_channels, out_channels, filter_order,
                    (filter_order, horizon, in_channels, out_channels), (1, L[0].shape[0], out_channels), bias,
                    recursion, engine)

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
