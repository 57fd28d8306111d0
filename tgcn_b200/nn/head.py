"""Fused classifier head (csrc/head.cu): fc1 -> BatchNorm1d -> ReLU -> fc2 -> log_softmax in two launches forward
and two backward, for the tail of the reference models (pytorch_hcp_tgcn.py:143-155).  The modules that own the
parameters stay ordinary torch.nn.Linear / BatchNorm1d (same state_dict); only the arithmetic is fused."""
import torch

from .. import _lib
from .functional import _DeviceGuard, _ptr, _require_cuda_f32, _stream


class _HeadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W1, b1, gamma, beta, W2, b2, running_mean, running_var, momentum, eps, training):
        lib = _lib.load()
        for t, nm in ((x, "x"), (W1, "fc1.weight"), (W2, "fc2.weight")):
            _require_cuda_f32(t, nm)
        x = x.contiguous()
        Q, I = x.shape
        Hd, C = W1.shape[0], W2.shape[0]
        dev = x.device
        act = torch.empty((Q, Hd), dtype=torch.float32, device=dev)
        xhat = torch.empty((Q, Hd), dtype=torch.float32, device=dev)
        invstd = torch.empty((Hd,), dtype=torch.float32, device=dev)
        logp = torch.empty((Q, C), dtype=torch.float32, device=dev)
        W1c, W2c = W1.contiguous(), W2.contiguous()
        with _DeviceGuard(dev):
            rc = lib.tgcn_head_fwd(_ptr(x), _ptr(W1c), _ptr(b1), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var),
                                   float(momentum), float(eps), int(training), _ptr(W2c), _ptr(b2), _ptr(act), _ptr(xhat),
                                   _ptr(invstd), _ptr(logp), Q, I, Hd, C, _stream(dev))
        _lib.check(rc, "tgcn_head_fwd")
        ctx.save_for_backward(x, W1c, gamma, W2c, act, xhat, invstd, logp)
        ctx.has = (b1 is not None, gamma is not None, beta is not None, b2 is not None)
        ctx.training = bool(training)
        return logp

    @staticmethod
    def backward(ctx, dlogp):
        if not ctx.training:
            raise RuntimeError("fused head: backward is implemented for training-mode batch statistics only")
        lib = _lib.load()
        x, W1, gamma, W2, act, xhat, invstd, logp = ctx.saved_tensors
        Q, I = x.shape
        Hd, C = W1.shape[0], W2.shape[0]
        dev = x.device
        dlogp = dlogp.contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        dx = torch.empty((Q, I), **f32) if ctx.needs_input_grad[0] else None
        dW1, dW2 = torch.empty((Hd, I), **f32), torch.empty((C, Hd), **f32)
        db1 = torch.empty((Hd,), **f32) if ctx.has[0] else None
        dgamma = torch.empty((Hd,), **f32) if ctx.has[1] else None
        dbeta = torch.empty((Hd,), **f32) if ctx.has[2] else None
        db2 = torch.empty((C,), **f32) if ctx.has[3] else None
        dh = torch.empty((Q, Hd), **f32)
        with _DeviceGuard(dev):
            rc = lib.tgcn_head_bwd(_ptr(dlogp), _ptr(logp), _ptr(act), _ptr(xhat), _ptr(invstd), _ptr(x), _ptr(W1), _ptr(gamma),
                                   _ptr(W2), _ptr(dx), _ptr(dW1), _ptr(db1), _ptr(dgamma), _ptr(dbeta), _ptr(dW2), _ptr(db2),
                                   _ptr(dh), Q, I, Hd, C, _stream(dev))
        _lib.check(rc, "tgcn_head_bwd")
        return dx, dW1, db1, dgamma, dbeta, dW2, db2, None, None, None, None, None


def fused_head(x, fc1, bn, fc2):
    """log_softmax(fc2(relu(bn(fc1(x)))), dim=1) with fc1, fc2: torch.nn.Linear and bn: torch.nn.BatchNorm1d."""
    training = bn.training or bn.running_mean is None
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = 0.1 if bn.momentum is None else bn.momentum
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    return _HeadFunction.apply(x, fc1.weight, fc1.bias, bn.weight, bn.bias, fc2.weight, fc2.bias, rm, rv, momentum, bn.eps,
                               training)
