"""autograd Functions over the C-ABI (include/tgcn_b200.h).  Host side of the drop-in boundary.

Every function requires CUDA fp32 tensors and raises otherwise -- there is no CPU path.
"""
import torch

from .. import _lib


import os as _os
# Register-tiled SpMM (csr.LaplacianCSR.ensure_rowtile_plans): TGCN_SPMM_ROWTILE=4|8 builds row-tile plans for every
# streaming layer whose row order has locality (measured: 1M-vertex geometric graph 636 -> 535 us per recursion step,
# cortical mesh 31 -> 28 us).  Opt-in because its summation order differs from the per-entry kernels' (results agree
# to fp32 rounding, ~1e-7 relative, not bit for bit); bench.py switches it on for the streaming workloads.
_ROWTILE_SPMM = int(_os.environ.get("TGCN_SPMM_ROWTILE", "0") or 0)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def _require_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: tgcn_b200 has no CPU path (CUDA sm_100a only)" % (name, t.device))
    if t.dtype != torch.float32:
        raise RuntimeError("%s has dtype %s: tgcn_b200 kernels are fp32" % (name, t.dtype))


def _drop_arg(drop):
    """drop = None or (p, seed, step_tensor_or_None) -> ctypes argument for a `const tgcn_dropout_t*` parameter."""
    if drop is None or not drop[0]:
        return None
    p, seed, step = drop
    if not 0.0 <= p < 1.0:
        raise ValueError("dropout probability has to be in [0, 1), got %r" % (p,))
    return _lib.dropout_arg(p, seed, None if step is None else step.data_ptr())


class _DeviceGuard:
    """Make the tensor's device current for the raw launches (ctypes calls bypass torch's guard)."""

    def __init__(self, device):
        self.dev = device.index if device.index is not None else torch.cuda.current_device()
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.dev:
            self.prev = cur
            torch.cuda.set_device(self.dev)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


class ChebLayerFunction(torch.autograd.Function):
    """out[q,n,g] = sum_k Xt_k[q,n,:] . W[k,:,g] + bias   (tgcn/nn/gcn.py:108-118 and friends)."""

    @staticmethod
    def forward(ctx, x, weight, bias, plan, bias_mode, recursion, engine):
        lib = _lib.load()
        Q, N, D = x.shape
        K, Dw, G = weight.shape
        if Dw != D:
            raise RuntimeError("weight expects %d features per vertex, input has %d" % (Dw, D))
        if N != plan.n:
            raise RuntimeError("input has %d vertices, Laplacian has %d" % (N, plan.n))
        dev = x.device
        x = x.contiguous()
        w = weight.contiguous()
        b = None if bias is None else bias.contiguous()
        if _ROWTILE_SPMM:
            plan.ensure_rowtile_plans(rows_per_tile=_ROWTILE_SPMM)
        out = torch.empty((Q, N, G), dtype=torch.float32, device=dev)
        # slab width of the streaming kernels: D, or D padded to whole 128-byte blocks for the TMA-fed contraction
        Dp = int(lib.tgcn_layer_slab_width(Q, N, D, G, K, engine))
        stack = torch.empty((K, N, Q * Dp), dtype=torch.float32, device=dev)
        fw_bytes = int(lib.tgcn_layer_fwd_workspace(Q, N, Dp, G, K))
        wmix = torch.empty((max(fw_bytes, 4) + 3) // 4, dtype=torch.float32, device=dev)   # [Wmix | engine scratch]
        with _DeviceGuard(dev):
            rc = lib.tgcn_layer_fwd(_ptr(plan.rowptr), _ptr(plan.col), _ptr(plan.val), N, _ptr(x), _ptr(w), _ptr(b),
                                    bias_mode if b is not None else _lib.BIAS_NONE, _ptr(out), _ptr(stack), _ptr(wmix),
                                    Q, D, Dp, G, K, recursion, engine, _stream(dev))
        _lib.check(rc, "tgcn_layer_fwd")
        ctx.plan = plan
        ctx.dims = (Q, N, D, G, K)
        ctx.Dp = Dp
        ctx.cfg = (bias_mode if b is not None else _lib.BIAS_NONE, recursion, engine)
        ctx.bias_shape = None if bias is None else tuple(bias.shape)
        ctx.w_shape = tuple(weight.shape)
        ctx.x_shape = tuple(x.shape)
        ctx.save_for_backward(stack, wmix)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        stack, wmix = ctx.saved_tensors
        Q, N, D, G, K = ctx.dims
        bias_mode, recursion, engine = ctx.cfg
        plan = ctx.plan
        dev = dout.device
        dout = dout.contiguous()
        need_dx = ctx.needs_input_grad[0]
        dW = torch.empty(ctx.w_shape, dtype=torch.float32, device=dev)
        db = torch.empty(ctx.bias_shape, dtype=torch.float32, device=dev) if bias_mode != _lib.BIAS_NONE else None
        dx = torch.empty(ctx.x_shape, dtype=torch.float32, device=dev) if need_dx else None
        Dp = ctx.Dp
        gstack = torch.empty((K, N, Q * Dp), dtype=torch.float32, device=dev) if need_dx else None
        ws_bytes = lib.tgcn_layer_bwd_workspace(Q, N, Dp, G, K)
        ws = torch.empty((max(int(ws_bytes), 4) + 3) // 4, dtype=torch.float32, device=dev)
        with _DeviceGuard(dev):
            rc = lib.tgcn_layer_bwd(_ptr(plan.rowptr_t), _ptr(plan.col_t), _ptr(plan.val_t), N, _ptr(dout), _ptr(stack),
                                    _ptr(wmix), _ptr(dW), _ptr(db), bias_mode, _ptr(dx), _ptr(gstack), _ptr(ws),
                                    Q, D, Dp, G, K, recursion, engine, _stream(dev))
        _lib.check(rc, "tgcn_layer_bwd")
        return dx, dW, db, None, None, None, None


def resident_supported(plan, D, G, K):
    """True when the sample-resident fused kernels cover this layer (per-sample slab fits in shared memory)."""
    return bool(_lib.load().tgcn_resident_supported(plan.n, D, G, K, plan.nnz))


class ResidentChebFunction(torch.autograd.Function):
    """Whole layer in one launch per direction (csrc/resident.cu): basis + contraction + bias, optionally
    followed by ReLU + permuted max-pool (`pool_p` in {2, 4}) without materialising the un-pooled
    activation.  Same math as ChebLayerFunction (+ PoolFunction); gcn.py:108-154, :246-255."""

    @staticmethod
    def forward(ctx, x, weight, bias, plan, bias_mode, recursion, pool_p, relu, drop=None):
        lib = _lib.load()
        if drop is not None and drop[0] and not (pool_p and relu):
            raise RuntimeError("dropout is fused between the ReLU and the pool only")
        Q, N, D = x.shape
        K, Dw, G = weight.shape
        if Dw != D:
            raise RuntimeError("weight expects %d features per vertex, input has %d" % (Dw, D))
        if N != plan.n:
            raise RuntimeError("input has %d vertices, Laplacian has %d" % (N, plan.n))
        if pool_p and N % pool_p:
            raise RuntimeError("shape '[%d, %d, %d, %d]' is invalid for input of size %d"
                               % (Q, N // pool_p, pool_p, G, Q * N * G))
        dev = x.device
        x = x.contiguous()
        w = weight.contiguous()
        b = None if bias is None else bias.contiguous()
        bm = bias_mode if b is not None else _lib.BIAS_NONE
        need_grad = any(ctx.needs_input_grad[:3])   # grad mode is off inside Function.forward
        stack = torch.empty(int(lib.tgcn_resident_stack_bytes(Q, N, D, K)) // 4, dtype=torch.float32, device=dev) \
            if need_grad else None
        if pool_p:
            out = None
            y = torch.empty((Q, N // pool_p, G), dtype=torch.float32, device=dev)
            idx = torch.empty((Q, N // pool_p, G), dtype=torch.uint8, device=dev)
        else:
            out = torch.empty((Q, N, G), dtype=torch.float32, device=dev)
            y = idx = None
        wimg = torch.empty(int(lib.tgcn_resident_weights_bytes(D, G, K)) // 4, dtype=torch.float32, device=dev)
        rowinfo, entries, E = plan.packed(lib.tgcn_resident_pack_classes(Q, N, D, 0))
        with _DeviceGuard(dev):
            rc = lib.tgcn_resident_layer_fwd(_ptr(rowinfo), _ptr(entries), N, E, _ptr(x), _ptr(w),
                                             _ptr(b), bm, _ptr(out), _ptr(y), _ptr(idx), pool_p, int(relu), _drop_arg(drop),
                                             _ptr(stack), _ptr(wimg), Q, D, G, K, recursion, _stream(dev))
        _lib.check(rc, "tgcn_resident_layer_fwd")
        ctx.plan = plan
        ctx.dims = (Q, N, D, G, K)
        ctx.cfg = (bm, recursion, pool_p, bool(relu))
        ctx.drop_p = float(drop[0]) if (drop is not None and drop[0]) else 0.0
        ctx.bias_shape = None if bias is None else tuple(bias.shape)
        ctx.w_shape = tuple(weight.shape)
        ctx.x_shape = tuple(x.shape)
        if pool_p:
            ctx.save_for_backward(stack, wimg, y, idx)
            ctx.mark_non_differentiable(idx)
            ctx.set_materialize_grads(False)      # no zero tensor for the (non-differentiable) index output
            return y, idx
        ctx.save_for_backward(stack, wimg)
        return out

    @staticmethod
    def backward(ctx, grad, _didx=None):
        lib = _lib.load()
        if grad is None:
            return (None,) * 9
        Q, N, D, G, K = ctx.dims
        bm, recursion, pool_p, relu = ctx.cfg
        plan = ctx.plan
        saved = ctx.saved_tensors
        stack, wimg = saved[0], saved[1]
        y, idx = (saved[2], saved[3]) if pool_p else (None, None)
        if stack is None:
            raise RuntimeError("resident layer was run without gradient tracking; nothing saved for backward")
        dev = grad.device
        grad = grad.contiguous()
        need_dx = ctx.needs_input_grad[0]
        dW = torch.empty(ctx.w_shape, dtype=torch.float32, device=dev)
        db = torch.empty(ctx.bias_shape, dtype=torch.float32, device=dev) if bm != _lib.BIAS_NONE else None
        dx = torch.empty(ctx.x_shape, dtype=torch.float32, device=dev) if need_dx else None
        ws = torch.empty(max(int(lib.tgcn_resident_bwd_workspace(Q, N, D, G, K)) // 4, 1), dtype=torch.float32, device=dev)
        rowinfo, entries, E = plan.packed(lib.tgcn_resident_pack_classes(Q, N, D, 1), transpose=True)
        with _DeviceGuard(dev):
            rc = lib.tgcn_resident_layer_bwd(_ptr(rowinfo), _ptr(entries), N, E,
                                             None if pool_p else _ptr(grad), _ptr(grad) if pool_p else None, _ptr(idx), _ptr(y),
                                             pool_p, int(relu), _lib.dropout_arg(ctx.drop_p), _ptr(stack), _ptr(wimg), _ptr(dW),
                                             _ptr(db), bm, _ptr(dx), _ptr(ws), Q, D, G, K, recursion, _stream(dev))
        _lib.check(rc, "tgcn_resident_layer_bwd")
        return dx, dW, db, None, None, None, None, None, None


class PoolFunction(torch.autograd.Function):
    """Permuted max-pool with first-argmax gradient routing (tgcn/nn/gcn.py:246-255), optional fused ReLU."""

    @staticmethod
    def forward(ctx, x, p, relu, drop=None):
        lib = _lib.load()
        _require_cuda_f32(x, "x")
        if drop is not None and drop[0] and not relu:
            raise RuntimeError("dropout is fused between the ReLU and the pool only")
        if x.dim() != 3:
            raise RuntimeError("pool expects [Q, N, G], got %s" % (tuple(x.shape),))
        Q, N, G = x.shape
        if N % p:
            raise RuntimeError("shape '[%d, %d, %d, %d]' is invalid for input of size %d"
                               % (Q, N // p, p, G, x.numel()))  # same failure mode as the reference's reshape
        x = x.contiguous()
        dev = x.device
        y = torch.empty((Q, N // p, G), dtype=torch.float32, device=dev)
        idx = torch.empty((Q, N // p, G), dtype=torch.uint8, device=dev)
        with _DeviceGuard(dev):
            rc = lib.tgcn_pool_max_fwd(_ptr(x), _ptr(y), _ptr(idx), Q, N, G, p, int(relu), _drop_arg(drop), _stream(dev))
        _lib.check(rc, "tgcn_pool_max_fwd")
        ctx.p, ctx.relu, ctx.dims = p, relu, (Q, N, G)
        ctx.drop_p = float(drop[0]) if (drop is not None and drop[0]) else 0.0
        if relu:
            # the ReLU (and dropout) gate of the backward is read off the pooled OUTPUT: y > 0 <=> its source passed
            # the ReLU (and was kept) -- a quarter of the bytes of the un-pooled input
            ctx.save_for_backward(idx, y)
        else:
            ctx.save_for_backward(idx)
        ctx.mark_non_differentiable(idx)
        ctx.set_materialize_grads(False)
        return y, idx

    @staticmethod
    def backward(ctx, dy, _didx=None):
        lib = _lib.load()
        if dy is None:
            return None, None, None, None
        saved = ctx.saved_tensors
        idx = saved[0]
        y = saved[1] if ctx.relu else None
        Q, N, G = ctx.dims
        dy = dy.contiguous()
        dev = dy.device
        dx = torch.empty((Q, N, G), dtype=torch.float32, device=dev)
        with _DeviceGuard(dev):
            rc = lib.tgcn_pool_max_bwd(_ptr(dy), _ptr(idx), None, _ptr(y), _ptr(dx), Q, N, G, ctx.p, int(ctx.relu),
                                       _lib.dropout_arg(ctx.drop_p), _stream(dev))
        _lib.check(rc, "tgcn_pool_max_bwd")
        return dx, None, None, None


def cheb_basis(x, plan, K, recursion=_lib.RECURSION_REFERENCE, reference_layout=True):
    """Stacked basis of x[Q,N,D].  reference_layout: Xt[K,Q,N,D] as `_time_chebyshev` returns it
    (gcn.py:126-154); otherwise the internal stack[K,N,Q*D]."""
    lib = _lib.load()
    _require_cuda_f32(x, "x")
    Q, N, D = x.shape
    dev = x.device
    x = x.contiguous()
    if _ROWTILE_SPMM:
        plan.ensure_rowtile_plans(rows_per_tile=_ROWTILE_SPMM)
    stack = torch.empty((K, N, Q * D), dtype=torch.float32, device=dev)
    with _DeviceGuard(dev):
        rc = lib.tgcn_cheb_basis(_ptr(plan.rowptr), _ptr(plan.col), _ptr(plan.val), N, _ptr(x), _ptr(stack), Q, D, K,
                                 recursion, _stream(dev))
        _lib.check(rc, "tgcn_cheb_basis")
        if not reference_layout:
            return stack
        Xt = torch.empty((K, Q, N, D), dtype=torch.float32, device=dev)
        rc = lib.tgcn_basis_to_reference(_ptr(stack), _ptr(Xt), Q, N, D, K, recursion, _stream(dev))
        _lib.check(rc, "tgcn_basis_to_reference")
    return Xt
