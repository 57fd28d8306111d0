"""torch-CPU port of the reference layers -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

The reference's hot path *is* a handful of ATen calls: a dense
``einsum("nm,qm..->qn..")`` per recursion step (-> ``aten::bmm``), an axpy, a copy into
the stacked basis, one big einsum against the weights, a broadcast bias add and
``torch.max(dim)`` for pooling; the backward is PyTorch autograd over those calls.
This module performs the same calls in the same order, written functionally, so that
timing it on the host cores measures what the reference costs on a CPU
(``bench.py`` ``cpu_baseline`` with ``kind="port"``).  With a torch-sparse ``L`` it
follows the ``torch.mm`` slab formulation of ``tgcn/nn/gcn_matmul.py:152-156`` (the only
reference path that is feasible for N >= 10^4, where dense L does not fit).

Nothing here is imported by the product (tgcn_b200/).
"""
import torch


def _vertex_mix(L, X):
    """One application of L along the vertex axis of X[Q,N,...]."""
    if L.layout == torch.strided:
        # gcn.py:147,152 / :72,77 / :230,235 -- dense einsum, lowers to bmm
        if X.dim() == 4:
            return torch.einsum("nm,qmhf->qnhf", L, X)
        return torch.einsum("nm,qmf->qnf", L, X)
    # gcn_matmul.py:152-156 -- [N, F*H*Q] slab through torch.mm (sparse L supported)
    order_fwd = (1, 3, 2, 0) if X.dim() == 4 else (1, 2, 0)
    order_bwd = (3, 0, 2, 1) if X.dim() == 4 else (2, 0, 1)
    Xp = X.permute(*order_fwd)
    slab = Xp.reshape(Xp.shape[0], -1)
    return torch.mm(L, slab).reshape(Xp.shape).permute(*order_bwd)


def basis_stack(L, X, K, recursion="reference"):
    """gcn.py:126-154 (see oracle/layers_np.py for the reference-vs-textbook note)."""
    stack = torch.empty((K,) + tuple(X.shape), dtype=X.dtype, device=X.device)
    stack[0] = X
    run = X
    if K > 1:
        run = _vertex_mix(L, run)
        stack[1] = run
    for k in range(2, K):
        if recursion == "reference":
            run = _vertex_mix(L, run)
            stack[k] = 2 * run - stack[k - 2]
        else:
            run = _vertex_mix(L, stack[k - 1])
            stack[k] = 2 * run - stack[k - 2]
    return stack


def cheb_layer(L, x, W, b=None, recursion="reference"):
    """TGCNCheb_H (W 4-D: [K,H,F,G]) or TGCNCheb / GCNCheb (W 3-D: [K,F,G]) forward."""
    if W.dim() == 4:
        if x.dim() == 3:
            x = x.unsqueeze(3)                                   # gcn.py:134-135
        out = torch.einsum("kqnhf,khfg->qng", basis_stack(L, x, W.shape[0], recursion), W)
    else:
        if x.dim() == 2:
            x = x.unsqueeze(2)                                   # gcn.py:216-217
        out = torch.einsum("kqnf,kfg->qng", basis_stack(L, x, W.shape[0], recursion), W)
    if b is not None:
        out = out + b
    return out


def pool(x, p):
    """gcn.py:246-255."""
    Q, N, G = x.shape
    return torch.max(x.reshape(Q, N // p, p, G), dim=2)[0]


def pool_with_indices(x, p):
    Q, N, G = x.shape
    return torch.max(x.reshape(Q, N // p, p, G), dim=2)
