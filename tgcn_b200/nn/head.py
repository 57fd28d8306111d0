"""Fused classifier head (csrc/head.cu, csrc/bighead.cu): fc1 -> BatchNorm1d -> ReLU -> dropout -> fc2 -> log_softmax
for the tail of the reference models (pytorch_hcp_tgcn.py:143-155).  The modules that own the parameters stay ordinary
torch.nn.Linear / BatchNorm1d (same state_dict); only the arithmetic is fused.

Small heads (parcellation-sized graphs) run in two launches forward and two backward.  A large fc1 (cortical mesh:
167 424 x 200 = 134 MB) is weight streaming: its forward reads the weight once, and its backward can apply the
optimizer step of `fc1.weight` in the same pass (`Fc1FusedSGD`), so the 134 MB gradient is never written -- nor
exchanged between data-parallel ranks, which trade their activations over NVLink peer memory instead.
"""
import ctypes

import torch
import torch.distributed as dist

from .. import _lib
from .functional import _DeviceGuard, _drop_arg, _ptr, _require_cuda_f32, _stream


class Fc1FusedSGD:
    """`torch.optim.SGD([fc1.weight], lr, momentum)` executed inside the head's backward (csrc/bighead.cu), with the
    gradient averaged over the data-parallel ranks.  Keep `fc1.weight` OUT of the model's regular optimizer and pass
    this object to `fused_head(..., fc1_update=...)`; the weight is then updated during `loss.backward()` and its
    `.grad` stays None.  world > 1: every rank owns an IPC-mapped region holding its x [Q, I] and dh [Q, Hd] of the
    current step (two buffers alternating with the step parity + step flags, csrc/peer.cu protocol); replicas stay
    bit-identical.  CUDA-graph capturable."""

    def __init__(self, weight, lr, momentum=0.0, batch=None, group=None):
        self.lib = _lib.load()
        if not (weight.is_cuda and weight.dtype == torch.float32 and weight.is_contiguous()):
            raise ValueError("Fc1FusedSGD needs a contiguous fp32 CUDA weight")
        self.weight = weight
        self.lr, self.momentum = float(lr), float(momentum)
        self.mom = torch.zeros_like(weight)
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.device = weight.device
        self.state = None
        self.gather = None
        self._regions_c = None
        self._regions = []
        self._batch = None
        if self.world > 1:
            if batch is None:
                raise ValueError("Fc1FusedSGD: world > 1 needs the per-rank batch size to size the exchange region")
            self._alloc(int(batch))

    def _alloc(self, Q):
        lib = self.lib
        Hd, I = self.weight.shape
        nbytes = int(lib.tgcn_peer_region_bytes(Q * I + Q * Hd, 2))
        own = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.tgcn_peer_alloc(nbytes, ctypes.byref(own)), "tgcn_peer_alloc")
            handle = (ctypes.c_ubyte * 64)()
            _lib.check(lib.tgcn_peer_export(own, handle), "tgcn_peer_export")
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=self.group)
            for r in range(self.world):
                if r == self.rank:
                    self._regions.append(own.value)
                else:
                    ptr = ctypes.c_void_p()
                    buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
                    _lib.check(lib.tgcn_peer_import(buf, ctypes.byref(ptr)), "tgcn_peer_import")
                    self._regions.append(ptr.value)
        dist.barrier(group=self.group)
        self._regions_c = (ctypes.c_void_p * self.world)(*self._regions)
        self.state = torch.zeros(4, dtype=torch.int32, device=self.device)
        nflat = ((Q * I + 3) // 4 * 4) + ((Q * Hd + 3) // 4 * 4)
        self.gather = torch.empty(self.world * nflat, dtype=torch.float32, device=self.device)
        self._batch = Q

    def supported(self, Q):
        Hd, I = self.weight.shape
        return bool(self.lib.tgcn_head_fused_update_supported(int(Q), int(I), int(Hd))) and \
            (self.world == 1 or Q == self._batch)

    def descriptor(self):
        """ctypes tgcn_fc1_update_t for one backward call (kept alive by the caller for the duration of the call)."""
        return _lib.Fc1Update(self.lr, self.momentum, self.mom.data_ptr(), self.world, self.rank,
                              None if self._regions_c is None else ctypes.cast(self._regions_c, ctypes.c_void_p),
                              None if self.state is None else self.state.data_ptr(),
                              None if self.gather is None else self.gather.data_ptr())


class _HeadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W1, b1, gamma, beta, W2, b2, running_mean, running_var, momentum, eps, training, drop, upd):
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
        ws_bytes = int(lib.tgcn_head_workspace(Q, I, Hd))
        ws = torch.empty((ws_bytes + 3) // 4, dtype=torch.float32, device=dev) if ws_bytes else None
        drop = drop if training else None
        with _DeviceGuard(dev):
            rc = lib.tgcn_head_fwd(_ptr(x), _ptr(W1c), _ptr(b1), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var),
                                   float(momentum), float(eps), int(training), _ptr(W2c), _ptr(b2), _drop_arg(drop), _ptr(act),
                                   _ptr(xhat), _ptr(invstd), _ptr(logp), _ptr(ws), Q, I, Hd, C, _stream(dev))
        _lib.check(rc, "tgcn_head_fwd")
        ctx.save_for_backward(x, W1c, gamma, W2c, act, xhat, invstd, logp)
        ctx.has = (b1 is not None, gamma is not None, beta is not None, b2 is not None)
        ctx.training = bool(training)
        ctx.drop_p = float(drop[0]) if (drop is not None and drop[0]) else 0.0
        ctx.upd = upd
        if upd is not None:
            if W1c.data_ptr() != upd.weight.data_ptr():
                raise RuntimeError("fused fc1 update: the descriptor belongs to a different weight tensor")
            if not upd.supported(Q):
                raise RuntimeError("fused fc1 update: batch %d / fc1 %d x %d is not covered" % (Q, Hd, I))
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
        upd = ctx.upd
        dx = torch.empty((Q, I), **f32) if ctx.needs_input_grad[0] else None
        dW1 = torch.empty((Hd, I), **f32) if upd is None else None
        dW2 = torch.empty((C, Hd), **f32)
        db1 = torch.empty((Hd,), **f32) if ctx.has[0] else None
        dgamma = torch.empty((Hd,), **f32) if ctx.has[1] else None
        dbeta = torch.empty((Hd,), **f32) if ctx.has[2] else None
        db2 = torch.empty((C,), **f32) if ctx.has[3] else None
        dh = torch.empty((Q, Hd), **f32)
        desc = upd.descriptor() if upd is not None else None
        with _DeviceGuard(dev):
            rc = lib.tgcn_head_bwd(_ptr(dlogp), _ptr(logp), _ptr(act), _ptr(xhat), _ptr(invstd), _ptr(x), _ptr(W1), _ptr(gamma),
                                   _ptr(W2), _ptr(dx), _ptr(dW1), _ptr(db1), _ptr(dgamma), _ptr(dbeta), _ptr(dW2), _ptr(db2),
                                   _ptr(dh), _lib.dropout_arg(ctx.drop_p), None if desc is None else ctypes.byref(desc),
                                   Q, I, Hd, C, _stream(dev))
        _lib.check(rc, "tgcn_head_bwd")
        return dx, dW1, db1, dgamma, dbeta, dW2, db2, None, None, None, None, None, None, None


def fused_head(x, fc1, bn, fc2, drop=None, fc1_update=None):
    """log_softmax(fc2(dropout(relu(bn(fc1(x))))), dim=1) with fc1, fc2: torch.nn.Linear and bn: torch.nn.BatchNorm1d.
    drop: None or (p, seed, step tensor) for the fused dropout (training only).  fc1_update: an `Fc1FusedSGD` whose
    optimizer step is applied inside the backward (large fc1 only)."""
    training = bn.training or bn.running_mean is None
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = 0.1 if bn.momentum is None else bn.momentum
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    return _HeadFunction.apply(x, fc1.weight, fc1.bias, bn.weight, bn.bias, fc2.weight, fc2.bias, rm, rv, momentum, bn.eps,
                               training, drop, fc1_update)


def head_covers(Q, I, Hd, training, grad_enabled):
    """True when the fused head handles this shape and mode: small heads always (batch statistics need Q > 1 when
    training; evaluation mode is fused for inference only), large ones when csrc/bighead.cu applies (Q <= 8)."""
    usable = (training and Q > 1) or (not training and not grad_enabled)
    if not usable:
        return False
    if I * Hd <= (1 << 21):
        return True
    return bool(_lib.load().tgcn_head_workspace(int(Q), int(I), int(Hd)))
