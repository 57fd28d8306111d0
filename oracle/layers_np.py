"""numpy (float64) restatement of the tgcn/nn layers -- TEST INFRASTRUCTURE ONLY.

Reference: cassianobecker/tgcn ``tgcn/nn/gcn.py`` (identical maths in
``tgcn/nn/gcn_matmul.py``).  Nothing here is imported by the product.

Important semantic note (verified against the unmodified reference, see
tests/golden/make_golden.py): the reference's "Chebyshev" loop never re-assigns
its running variable to the stacked polynomial -- the line ``X = Xt[k-1]`` is
commented out (gcn.py:75,150,232) -- so what the layers really stack is

    P_0 = X,  P_j = L P_{j-1}                      (plain powers of L)
    Xt_0 = P_0, Xt_1 = P_1, Xt_k = 2 P_k - Xt_{k-2}     (k >= 2)

which differs from the true recursion T_k = 2 L T_{k-1} - T_{k-2} for k >= 3.
``recursion="reference"`` restates that behaviour (the parity target);
``recursion="chebyshev"`` is the textbook recursion (as in
``examples/tgcn_mnist.py:200-216`` and the 4-D branch of ``gcn/graph.py:266-283``).
"""
import numpy as np
import scipy.sparse as sp

F64 = np.float64


def _as_op(L):
    """Return (matvec over [N, cols], transpose-matvec) for dense or scipy-sparse L."""
    if sp.issparse(L):
        Lc = L.tocsr().astype(F64)
        Lt = Lc.T.tocsr()
        return (lambda M: Lc @ M), (lambda M: Lt @ M)
    Ld = np.asarray(L, dtype=F64)
    return (lambda M: Ld @ M), (lambda M: Ld.T @ M)


def _apply_vertex_op(op, X):
    """Apply an [N,N] operator along axis 1 of X[Q,N,...]  (einsum "nm,qm...->qn...",
    gcn.py:147; equivalently the permute/mm/permute of gcn_matmul.py:152-156)."""
    Xv = np.moveaxis(X, 1, 0)                 # [N, Q, ...]
    shp = Xv.shape
    Y = op(Xv.reshape(shp[0], -1)).reshape(shp)
    return np.moveaxis(Y, 0, 1)


def cheb_basis(L, X, K, recursion="reference"):
    """Stacked basis Xt[K, *X.shape]  (gcn.py:126-154 ``_time_chebyshev``,
    :208-237 ``_chebyshev``, :52-79 for TGCNCheb)."""
    X = np.asarray(X, dtype=F64)
    mv, _ = _as_op(L)
    Xt = np.empty((K,) + X.shape, dtype=F64)
    Xt[0] = X                                                    # gcn.py:143
    if recursion == "reference":
        run = X
        if K > 1:
            run = _apply_vertex_op(mv, run)                      # gcn.py:146-148
            Xt[1] = run
        for k in range(2, K):
            run = _apply_vertex_op(mv, run)                      # gcn.py:152 (run is L^k X)
            Xt[k] = 2.0 * run - Xt[k - 2]                        # gcn.py:153
    elif recursion == "chebyshev":
        if K > 1:
            Xt[1] = _apply_vertex_op(mv, X)
        for k in range(2, K):
            Xt[k] = 2.0 * _apply_vertex_op(mv, Xt[k - 1]) - Xt[k - 2]
    else:
        raise ValueError(recursion)
    return Xt


def _canon_x(x, kind):
    """Input shape rules: TGCNCheb_H accepts [Q,N,H] or [Q,N,H,F] (gcn.py:134-135);
    GCNCheb accepts [Q,N] or [Q,N,F] (gcn.py:216-217); TGCNCheb takes [Q,N,F] (gcn.py:52-63)."""
    x = np.asarray(x, dtype=F64)
    if kind == "tgcn_h":
        if x.ndim == 3:
            x = x[..., None]
        assert x.ndim == 4
    elif kind == "gcn":
        if x.ndim == 2:
            x = x[..., None]
        assert x.ndim == 3
    elif kind == "tgcn":
        assert x.ndim == 3
    else:
        raise ValueError(kind)
    return x


def layer_forward(L, x, W, b=None, kind="tgcn_h", recursion="reference"):
    """out[q,n,g] = sum_{k,(h),f} Xt[k,q,n,(h),f] W[k,(h),f,g] (+ bias)
    (gcn.py:108-118 / :34-44 / :189-199).  bias broadcasts: [1,N,G] or [1,1,G]."""
    x = _canon_x(x, kind)
    W = np.asarray(W, dtype=F64)
    K = W.shape[0]
    Xt = cheb_basis(L, x, K, recursion)
    if kind == "tgcn_h":
        out = np.einsum("kqnhf,khfg->qng", Xt, W, optimize=True)   # gcn.py:113 (optimize: BLAS contraction, same sum)
    else:
        out = np.einsum("kqnf,kfg->qng", Xt, W, optimize=True)     # gcn.py:39,194
    if b is not None:
        out = out + np.asarray(b, dtype=F64)                     # gcn.py:115-116
    return out


def layer_backward(L, x, W, dout, bias_shape=None, kind="tgcn_h", recursion="reference",
                   need_dx=True):
    """Analytic reverse pass of ``layer_forward`` (what ``loss.backward()`` computes in the
    reference through PyTorch autograd).  Returns (dW, db, dx)."""
    x0 = np.asarray(x)
    xin = _canon_x(x, kind)
    W = np.asarray(W, dtype=F64)
    dout = np.asarray(dout, dtype=F64)
    K = W.shape[0]
    Xt = cheb_basis(L, xin, K, recursion)
    if kind == "tgcn_h":
        dW = np.einsum("kqnhf,qng->khfg", Xt, dout, optimize=True)
        dXt = np.einsum("qng,khfg->kqnhf", dout, W, optimize=True) if need_dx else None
    else:
        dW = np.einsum("kqnf,qng->kfg", Xt, dout, optimize=True)
        dXt = np.einsum("qng,kfg->kqnf", dout, W, optimize=True) if need_dx else None
    db = None
    if bias_shape is not None:
        if tuple(bias_shape)[1] == 1:
            db = dout.sum(axis=(0, 1)).reshape(bias_shape)       # bias [1,1,G]  (gcn.py:172)
        else:
            db = dout.sum(axis=0).reshape(bias_shape)            # bias [1,N,G]  (gcn.py:96)
    dx = None
    if need_dx:
        _, mvT = _as_op(L)
        dXt = [dXt[k].copy() for k in range(K)]
        if recursion == "reference":
            dP = [np.zeros_like(xin) for _ in range(K)]
            for k in range(K - 1, 1, -1):                        # Xt_k = 2 P_k - Xt_{k-2}
                dP[k] += 2.0 * dXt[k]
                dXt[k - 2] -= dXt[k]
            if K > 1:
                dP[1] += dXt[1]
            dP[0] += dXt[0]
            for k in range(K - 1, 0, -1):                        # P_k = L P_{k-1}
                dP[k - 1] += _apply_vertex_op(mvT, dP[k])
            dx = dP[0]
        else:
            for k in range(K - 1, 1, -1):                        # T_k = 2 L T_{k-1} - T_{k-2}
                dXt[k - 1] += 2.0 * _apply_vertex_op(mvT, dXt[k])
                dXt[k - 2] -= dXt[k]
            if K > 1:
                dXt[0] += _apply_vertex_op(mvT, dXt[1])
            dx = dXt[0]
        dx = dx.reshape(x0.shape)
    return dW, db, dx


def mix_matrix(K, recursion="reference"):
    """M[k,j] with Xt_k = sum_j M[k,j] L^j X for the reference recursion (identity for the
    textbook recursion, whose basis the product computes directly)."""
    M = np.zeros((K, K), dtype=F64)
    if recursion == "chebyshev":
        return np.eye(K)
    for k in range(K):
        if k < 2:
            M[k, k] = 1.0
        else:
            M[k] = -M[k - 2]
            M[k, k] += 2.0
    return M


def pool_forward(x, p):
    """``gcn_pool`` (p=2) / ``gcn_pool_4`` (p=4): reshape [Q,N/p,p,G], max over the p siblings
    (gcn.py:246-255).  Returns (values, argmax in 0..p-1) with torch.max's CPU rule: first
    maximal element wins, and a NaN beats everything (first NaN wins)."""
    x = np.asarray(x)
    Q, N, G = x.shape
    if N % p:
        raise ValueError("vertex count %d not divisible by pool size %d" % (N, p))
    xr = x.reshape(Q, N // p, p, G)
    best = xr[:, :, 0, :].copy()
    idx = np.zeros((Q, N // p, G), dtype=np.int64)
    for s in range(1, p):
        cand = xr[:, :, s, :]
        take = (cand > best) | (np.isnan(cand) & ~np.isnan(best))
        best = np.where(take, cand, best)
        idx = np.where(take, s, idx)
    return best, idx


def pool_backward(dy, idx, p):
    """Route each pooled gradient to its argmax sibling (autograd of torch.max(dim))."""
    dy = np.asarray(dy)
    Q, Np, G = dy.shape
    dx = np.zeros((Q, Np, p, G), dtype=dy.dtype)
    q, n, g = np.meshgrid(np.arange(Q), np.arange(Np), np.arange(G), indexing="ij")
    dx[q, n, idx, g] = dy
    return dx.reshape(Q, Np * p, G)
