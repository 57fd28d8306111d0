"""Data-parallel plumbing for the TGCN path: one process per GPU, the CSR
Laplacian and the weights replicated, the batch sharded, and ONE allreduce of a flat fp32
gradient buffer per step (NCCL over NVLink on GPUs; gloo in the CPU tests).  Replaces the
reference's single-process `torch.nn.DataParallel` wrapper (pytorch_hcp_tgcn.py:271-272).
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, world, local_rank).
    A single process without the env vars runs as world 1 without a process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(total, rank, world):
    """Contiguous [lo, hi) of `total` units owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class FlatGradients:
    """All parameter gradients as views into one flat fp32 buffer, so that a step needs exactly one
    collective.  `p.grad` is bound to its view; autograd accumulates in place."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatGradients needs fp32 parameters on one device")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def allreduce_mean(self, group=None):
        """Sum over ranks and divide by the world size: per-rank mean losses then reproduce the
        reference's global-batch mean loss ('DataParallel semantics')."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))
        return self.flat


class GradientBucket:
    """Gradient exchange for a training step that lets autograd ALLOCATE the gradients
    (`zero_grad(set_to_none=True)`: no zero fill, no accumulate kernels): after the backward the
    per-parameter gradients are packed into one flat fp32 buffer by multi-tensor copies, averaged over the
    ranks, and `p.grad` is re-bound to the views of the flat buffer so the optimizer reads the averaged values.

    Overlap: parameters listed in `late` (the layers whose gradients are produced LAST by the backward, i.e.
    the first layers of the model) form a second bucket.  As soon as every other gradient exists (autograd
    post-accumulate hooks) the first bucket's allreduce is issued asynchronously, so it runs under the
    remaining backward; `sync()` reduces the late bucket and joins.  Everything is capturable in a CUDA graph
    (NCCL included).  With world size 1 it does nothing."""

    def __init__(self, params, group=None, late=()):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("no trainable parameters")
        late_ids = {id(p) for p in late}
        self.early = [p for p in params if id(p) not in late_ids]
        self.late = [p for p in params if id(p) in late_ids]
        self.params = self.early + self.late
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.n_early = sum(p.numel() for p in self.early)
        self._pending = None
        self._seen = set()
        self._hooks = []
        if self.world > 1 and self.late and self.early:
            for p in self.early:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _reduce(self, lo, hi, async_op):
        buf = self.flat[lo:hi]
        if dist.get_backend(self.group) == "nccl":
            return dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

    def _pack(self, params, views):
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
        torch._foreach_copy_(views, grads)

    def _on_grad(self, param):
        if self._pending is not None:
            return
        self._seen.add(id(param))                  # parameters seen, not hook firings
        if len(self._seen) == len(self.early):     # every early gradient exists: reduce them under the rest of the backward
            self._pack(self.early, self.views[:len(self.early)])
            self._pending = self._reduce(0, self.n_early, async_op=True)

    def sync(self):
        if self.world == 1:
            return
        ne = len(self.early)
        if self._pending is None:                  # no hooks (single bucket) or a gradient never arrived: do it now
            self._pack(self.early, self.views[:ne])
            self._reduce(0, self.n_early, async_op=False)
        if self.late:
            self._pack(self.late, self.views[ne:])
            self._reduce(self.n_early, self.flat.numel(), async_op=False)
        if self._pending is not None:
            self._pending.wait()
        self._pending, self._seen = None, set()
        if dist.get_backend(self.group) != "nccl":
            self.flat.div_(self.world)
        for p, v in zip(self.params, self.views):
            p.grad = v


class _PeerGroup:
    """One IPC-shared gradient region (two flat buffers + ready flag, csrc/peer.cu) for a fixed list of parameters,
    with its own momentum buffers and device-side step counter."""

    def __init__(self, lib, params, world, rank, group):
        import ctypes
        from . import _lib
        self.lib, self._ct = lib, ctypes
        self.params = params
        self.world, self.rank = world, rank
        dev = params[0].device
        self.device = dev
        self.n = sum(p.numel() for p in params)
        self.moms = [torch.zeros_like(p) for p in params]
        self.state = torch.zeros(4, dtype=torch.int32, device=dev)
        nbytes = int(lib.tgcn_peer_region_bytes(self.n, len(params)))
        own = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.tgcn_peer_alloc(nbytes, ctypes.byref(own)), "tgcn_peer_alloc")
            handle = (ctypes.c_ubyte * 64)()
            _lib.check(lib.tgcn_peer_export(own, handle), "tgcn_peer_export")
            handles = [bytes(handle)]
            if world > 1:
                handles = [None] * world
                dist.all_gather_object(handles, bytes(handle), group=group)
            self.regions = []
            for r in range(world):
                if r == rank:
                    self.regions.append(own.value)
                else:
                    ptr = ctypes.c_void_p()
                    buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
                    _lib.check(lib.tgcn_peer_import(buf, ctypes.byref(ptr)), "tgcn_peer_import")
                    self.regions.append(ptr.value)
        if world > 1:
            dist.barrier(group=group)
        self._regions_c = (ctypes.c_void_p * world)(*self.regions)
        self._numels_c = (ctypes.c_int64 * len(params))(*[p.numel() for p in params])
        self._keep = []

    def launch(self, lr, momentum, stream):
        from . import _lib
        ct = self._ct
        k = len(self.params)
        # contiguous copies (if any were needed) must stay alive until the launches that read them are enqueued ON THE
        # STREAM THAT USES THEM -- they are held until the next launch of this group, and recorded on that stream so the
        # caching allocator does not hand their memory to the producing stream early
        held = [None if p.grad is None else p.grad.contiguous() for p in self.params]
        for g in held:
            if g is not None:
                g.record_stream(stream)
        self._keep = held
        grads = (ct.c_void_p * k)(*[None if g is None else g.data_ptr() for g in held])
        prms = (ct.c_void_p * k)(*[p.data_ptr() for p in self.params])
        moms = (ct.c_void_p * k)(*[m.data_ptr() for m in self.moms])
        with torch.cuda.device(self.device):
            rc = self.lib.tgcn_peer_allreduce_sgd(self._regions_c, self.world, self.rank, grads, prms, moms, self._numels_c, k,
                                                  lr, momentum, self.state.data_ptr(), stream.cuda_stream)
        _lib.check(rc, "tgcn_peer_allreduce_sgd")


class PeerAllreduceSGD:
    """Data-parallel optimizer step as ONE fused operation over NVLink peer memory (csrc/peer.cu): pack the local
    gradients into this rank's IPC-shared region, then read every rank's gradients with P2P loads, average them in
    rank order and apply SGD with momentum -- two launches, no NCCL, no averaged-gradient tensor.  Semantics of
    `torch.optim.SGD(params, lr, momentum)` (no weight decay / dampening / nesterov), with the gradient averaged
    over the ranks; replicas stay bit-identical.  Single node, world <= 8.  CUDA-graph capturable.

    Overlap (`late`, world > 1): the parameters listed in `late` are the ones whose gradients the backward produces
    LAST (the first layers of the model).  All other parameters form an early group with its own region: as soon as
    the last of their gradients exists (autograd post-accumulate hooks) their exchange + update is launched on a
    side stream and runs under the rest of the backward; `step()` then exchanges only the late group and joins the
    side stream.  Requirement: nothing enqueued after the last early gradient reads an early parameter (true for a
    feed-forward model whose `late` parameters belong to its first layers), and ONE backward pass per `step()` -- with
    gradient accumulation over several backward passes construct the optimizer without `late`."""

    def __init__(self, params, lr, momentum=0.0, group=None, late=()):
        from . import _lib
        self.lib = _lib.load()
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.momentum = float(lr), float(momentum)
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        dev = self.params[0].device
        self.device = dev
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev or not p.is_contiguous():
                raise ValueError("PeerAllreduceSGD needs contiguous fp32 parameters on one CUDA device")
        late_ids = {id(p) for p in late}
        early = [p for p in self.params if id(p) not in late_ids]
        latep = [p for p in self.params if id(p) in late_ids]
        self._hooks, self._seen, self._launched = [], set(), False
        if self.world > 1 and early and latep:
            self._early = _PeerGroup(self.lib, early, self.world, self.rank, group)
            self._main = _PeerGroup(self.lib, latep, self.world, self.rank, group)
            self._side = torch.cuda.Stream(device=dev)
            for p in early:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        else:
            self._early = None
            self._main = _PeerGroup(self.lib, self.params, self.world, self.rank, group)
        self.n = sum(p.numel() for p in self.params)

    @property
    def moms(self):
        return (self._early.moms if self._early else []) + self._main.moms

    @property
    def state(self):
        """Device-side counters of the group exchanged at `step()` (element 0 = steps taken)."""
        return self._main.state

    def _on_grad(self, param):
        # which parameters have produced a gradient, not how often a hook fired: gradient accumulation over several
        # backward passes must not trigger the exchange early
        if self._launched:
            return
        self._seen.add(id(param))
        if len(self._seen) == len(self._early.params):      # every early gradient exists: exchange + update them now
            cur = torch.cuda.current_stream(self.device)
            self._side.wait_stream(cur)
            self._early.launch(self.lr, self.momentum, self._side)
            self._launched = True

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def step(self):
        cur = torch.cuda.current_stream(self.device)
        if self._early is not None and not self._launched:     # gradients were set by hand (no backward): do it here
            self._early.launch(self.lr, self.momentum, cur)
        self._main.launch(self.lr, self.momentum, cur)
        if self._launched:
            cur.wait_stream(self._side)
        self._seen, self._launched = set(), False


def broadcast_parameters(module, src=0, group=None):
    """Replicate rank `src`'s parameters and buffers (what DataParallel's replicate does per step,
    done once here)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


# ------------------------------------------------------------------------------------------------
# Row-partitioned SpMM with halo exchange (graphs too large for
# one GPU's bandwidth budget).  Rank p owns a contiguous block of rows of L~ and of every slab;
# before each recursion step it needs the previous slab's rows for the columns it references
# outside its block (the halo).  The plan is computed once on the host.
# ------------------------------------------------------------------------------------------------
class RowPartition:
    """Host-side plan for one rank.

    local CSR: rows [lo, hi) of L~, columns renumbered to [0, n_own) for owned rows followed by
    [n_own, n_own + n_halo) for halo rows (sorted by global id, grouped by owner).
    send_idx[q]: local row ids (0..n_own) this rank sends to rank q each step.
    recv_cnt[q]: number of halo rows received from rank q (they land contiguously, in owner order).
    """

    def __init__(self, L_csr, rank, world, bounds=None):
        import numpy as np
        L = L_csr.tocsr()
        n = L.shape[0]
        self.rank, self.world, self.n_global = rank, world, n
        if bounds is None:
            bounds = [shard_range(n, r, world) for r in range(world)]
        self.bounds = bounds
        lo, hi = bounds[rank]
        self.lo, self.hi, self.n_own = lo, hi, hi - lo
        starts = np.array([b[0] for b in bounds] + [n])
        sub = L[lo:hi].tocsr()
        sub.sort_indices()
        cols = sub.indices.astype(np.int64)
        outside = (cols < lo) | (cols >= hi)
        halo_ids = np.unique(cols[outside])                       # sorted => grouped by owner
        owner = np.searchsorted(starts, halo_ids, side="right") - 1
        self.halo_ids = halo_ids
        self.n_halo = int(halo_ids.size)
        self.recv_cnt = [int((owner == q).sum()) for q in range(world)]
        remap = np.empty(cols.shape, dtype=np.int64)
        remap[~outside] = cols[~outside] - lo
        remap[outside] = self.n_own + np.searchsorted(halo_ids, cols[outside])
        self.rowptr = sub.indptr.astype(np.int32)
        self.col = remap.astype(np.int32)
        self.val = sub.data.astype(np.float32)
        self.recv_ids = [halo_ids[owner == q] for q in range(world)]   # global ids wanted from q
        self.send_idx = None                                           # filled by exchange_plans

    @staticmethod
    def exchange_plans(plans):
        """Single-process helper (tests / one host building every rank's plan): derive send lists."""
        import numpy as np
        for p in plans:
            p.send_idx = [None] * p.world
        for p in plans:
            for q in range(p.world):
                plans[q].send_idx[p.rank] = (p.recv_ids[q] - plans[q].lo).astype(np.int64)
        return plans

    def build_send_lists(self, group=None):
        """Multi-process: tell every owner which of its rows this rank needs (one all-gather of the
        wanted-id lists; run once per graph)."""
        import numpy as np
        wanted = [None] * self.world
        dist.all_gather_object(wanted, [ids.tolist() for ids in self.recv_ids], group=group)
        self.send_idx = [np.asarray(wanted[q][self.rank], dtype=np.int64) - self.lo for q in range(self.world)]
        return self


def halo_exchange(slab_own, plan, halo_out=None, group=None):
    """Gather the rows other ranks need from `slab_own` [n_own, C] and receive this rank's halo rows
    into `halo_out` [n_halo, C] (point-to-point sends grouped per step: NCCL send/recv over NVLink on
    GPUs, gloo on CPU)."""
    C = slab_own.shape[1]
    if halo_out is None:
        halo_out = slab_own.new_empty((plan.n_halo, C))
    ops, keep = [], []
    off = 0
    for q in range(plan.world):
        cnt = plan.recv_cnt[q]
        if q != plan.rank and cnt:
            ops.append(dist.P2POp(dist.irecv, halo_out[off:off + cnt], q, group=group))
        off += cnt
    for q in range(plan.world):
        idx = plan.send_idx[q]
        if q != plan.rank and idx is not None and len(idx):
            buf = slab_own.index_select(0, torch.as_tensor(idx, device=slab_own.device))
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, q, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return halo_out


class RowPartitionedLayer:
    """One TGCNCheb_H-style layer on a graph whose ROWS are partitioned over the ranks (SURVEY.md section 8e,
    config 4): forward and weight/bias backward, driven straight through the C-ABI (no autograd).

    Each rank keeps its rows of every basis slab in an EXTENDED slab `[n_own + n_halo, C]`: the owned rows
    followed by the halo rows other ranks send before each recursion step (`halo_exchange`).  The contraction
    and the weight gradient run on the extended slabs (halo rows of dOut are zero, so they add nothing to dW);
    dW is summed over the ranks with one allreduce.  Per-row results are bit-identical to the unpartitioned
    layer because the local CSR keeps each row's entry order.  Input gradients are not produced (first-layer
    use, as in the reference models where the input does not require grad)."""

    def __init__(self, L_csr, K, D, G, rank=0, world=1, device=None, recursion=0, engine=0, group=None,
                 rows_per_tile=0, rowtile_pad=1, halo="auto"):
        """rows_per_tile = 4 or 8: register the row-tile plan of this rank's rows (register-tiled SpMM kernel,
        include/tgcn_b200.h `tgcn_rowtile_plan_create`) when the row order has locality; 0: per-entry kernels.
        rowtile_pad: see csr.make_rowtile_plan.
        halo: "peer" = the basis slabs live in IPC-mapped regions and the halo rows are read from their owners with
        P2P loads inside a gather kernel (tgcn_halo_signal / tgcn_halo_pull: no host round trip, no NCCL);
        "nccl" = index_select + grouped NCCL send/recv per step; "auto" = peer on CUDA with world > 1."""
        from . import _lib
        self.lib = _lib.load()
        self.K, self.D, self.G = K, D, G
        self.rank, self.world, self.group = rank, world, group
        self.halo_mode = halo
        self.recursion, self.engine = recursion, engine
        self.device = torch.device(device if device is not None else "cuda")
        self.plan = RowPartition(L_csr, rank, world)
        if world > 1:
            self.plan.build_send_lists(group=group)
        else:
            self.plan.send_idx = [None]
        pl = self.plan
        self.n_own, self.n_ext = pl.n_own, pl.n_own + pl.n_halo
        dev = self.device
        self.rowptr = torch.tensor(pl.rowptr, device=dev)
        self.col = torch.tensor(pl.col, device=dev)
        self.val = torch.tensor(pl.val, device=dev)
        self.send_idx_dev = [None if (i is None or len(i) == 0) else torch.as_tensor(i, device=dev) for i in pl.send_idx]
        self._bufs = {}
        self._peer = None
        if world > 1 and self.halo_mode in ("auto", "peer") and self.device.type == "cuda":
            import numpy as np
            starts = np.array([b[0] for b in pl.bounds] + [pl.n_global])
            owner = (np.searchsorted(starts, pl.halo_ids, side="right") - 1).astype(np.int32)
            row = (pl.halo_ids - starts[owner]).astype(np.int32)
            self.halo_owner = torch.as_tensor(owner, device=dev)
            self.halo_row = torch.as_tensor(row, device=dev)
            self.halo_mode = "peer"
        elif self.halo_mode == "auto":
            self.halo_mode = "nccl"
        self.rowtile = None
        if rows_per_tile:
            from .csr import make_rowtile_plan
            self.rowtile = make_rowtile_plan(pl.rowptr, pl.col, pl.val, self.n_own, rows_per_tile, self.col, pad=rowtile_pad)

    def __del__(self):
        try:
            if self.rowtile is not None:
                self.lib.tgcn_rowtile_plan_destroy(self.rowtile[0])
        except Exception:
            pass

    def _buffers(self, Q):
        b = self._bufs.get(Q)
        if b is None:
            lib, K, D, G, n = self.lib, self.K, self.D, self.G, self.n_ext
            dev = self.device
            C = Q * D
            f32 = dict(dtype=torch.float32, device=dev)
            stack = self._peer_stack(Q) if self.halo_mode == "peer" else torch.zeros((K, n, C), **f32)
            b = dict(stack=stack, out=torch.empty((Q, n, G), **f32),
                     dout=torch.zeros((Q, n, G), **f32), bias=torch.zeros((n, G), **f32), xext=torch.zeros((Q, n, D), **f32),
                     wmix=torch.empty((K, D, G), **f32), dwmix=torch.empty((K, D, G), **f32),
                     scr=torch.empty(max(int(lib.tgcn_contract_fwd_scratch(Q, n, D, G, K)), 16) // 4 + 64, **f32),
                     ws=torch.empty(max(int(lib.tgcn_layer_bwd_workspace(Q, n, D, G, K)), 16) // 4 + 64, **f32),
                     db=torch.empty((n, G), **f32))
            self._bufs[Q] = b
        return b

    def _peer_stack(self, Q):
        """The basis stack [K, n_ext, C] of this rank inside an IPC-mapped region (+ a 256-byte flag line), mapped by every
        peer.  Collective (all ranks allocate together)."""
        import ctypes
        import numpy as np
        from . import _lib
        lib = self.lib
        C = Q * self.D
        nfl = self.K * self.n_ext * C
        flag_off = (nfl * 4 + 255) // 256 * 256
        own = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.tgcn_peer_alloc(flag_off + 256, ctypes.byref(own)), "tgcn_peer_alloc")
            handle = (ctypes.c_ubyte * 64)()
            _lib.check(lib.tgcn_peer_export(own, handle), "tgcn_peer_export")
            info = [None] * self.world
            dist.all_gather_object(info, (bytes(handle), self.n_ext, flag_off), group=self.group)
            regions = []
            for r in range(self.world):
                if r == self.rank:
                    regions.append(own.value)
                else:
                    ptr = ctypes.c_void_p()
                    buf = (ctypes.c_ubyte * 64).from_buffer_copy(info[r][0])
                    _lib.check(lib.tgcn_peer_import(buf, ctypes.byref(ptr)), "tgcn_peer_import")
                    regions.append(ptr.value)
        dist.barrier(group=self.group)
        # a torch view of the own region (the library keeps ownership of the allocation for the process lifetime)
        iface = {"shape": (self.K, self.n_ext, C), "typestr": "<f4", "data": (own.value, False), "version": 3}
        holder = type("PeerRegion", (), {"__cuda_array_interface__": iface})()
        stack = torch.as_tensor(holder, device=self.device)
        self._peer = dict(regions=(ctypes.c_void_p * self.world)(*regions),
                          flag_off=(ctypes.c_int64 * self.world)(*[i[2] for i in info]),
                          slab_elems=[i[1] * C for i in info], holder=holder, Q=Q,
                          state=torch.zeros(4, dtype=torch.int32, device=self.device))
        return stack

    def _halo(self, slab_ext, j=None):
        if self.world == 1:
            return
        pl = self.plan
        if self.halo_mode != "peer" and pl.n_halo == 0:
            return
        if self.halo_mode == "peer":
            import ctypes
            from . import _lib
            pr = self._peer
            st = torch.cuda.current_stream(self.device).cuda_stream
            offs = (ctypes.c_int64 * self.world)(*[j * e for e in pr["slab_elems"]])
            _lib.check(self.lib.tgcn_halo_signal(pr["regions"], pr["flag_off"], self.world, self.rank, pr["state"].data_ptr(), st),
                       "tgcn_halo_signal")
            if pl.n_halo:          # a rank without halo rows still signals: its peers read ITS rows
                _lib.check(self.lib.tgcn_halo_pull(pr["regions"], offs, pr["flag_off"], self.world, self.rank,
                                                   self.halo_owner.data_ptr(), self.halo_row.data_ptr(), pl.n_halo, slab_ext.shape[1],
                                                   slab_ext[self.n_own:].data_ptr(), pr["state"].data_ptr(), st), "tgcn_halo_pull")
            return
        ops, keep, off = [], [], 0
        for q in range(pl.world):
            cnt = pl.recv_cnt[q]
            if q != pl.rank and cnt:
                ops.append(dist.P2POp(dist.irecv, slab_ext[self.n_own + off:self.n_own + off + cnt], q, group=self.group))
            off += cnt
        for q in range(pl.world):
            idx = self.send_idx_dev[q]
            if q != pl.rank and idx is not None:
                buf = slab_ext[:self.n_own].index_select(0, idx)
                keep.append(buf)
                ops.append(dist.P2POp(dist.isend, buf, q, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def forward(self, x_own, weight, bias_own):
        """x_own [Q, n_own, D], weight [K, D, G], bias_own [n_own, G] or None -> out [Q, n_own, G] (a view)."""
        from . import _lib
        lib = self.lib
        Q = x_own.shape[0]
        b = self._buffers(Q)
        K, D, G, n = self.K, self.D, self.G, self.n_ext
        C = Q * D
        st = torch.cuda.current_stream(self.device).cuda_stream
        b["xext"][:, :self.n_own] = x_own
        _lib.check(lib.tgcn_to_slab(b["xext"].data_ptr(), b["stack"][0].data_ptr(), Q, n, D, st), "tgcn_to_slab")
        for j in range(1, K):
            self._halo(b["stack"][j - 1], j - 1)
            prev = None
            alpha, beta = 1.0, 0.0
            if self.recursion == 1 and j >= 2:
                prev, alpha, beta = b["stack"][j - 2].data_ptr(), 2.0, -1.0
            _lib.check(lib.tgcn_spmm_step(self.rowptr.data_ptr(), self.col.data_ptr(), self.val.data_ptr(), self.n_own,
                                          b["stack"][j - 1].data_ptr(), prev, b["stack"][j].data_ptr(), C, alpha, beta, st),
                       "tgcn_spmm_step")
        _lib.check(lib.tgcn_mix_weights(weight.data_ptr(), b["wmix"].data_ptr(), K, D * G, self.recursion, 0, st), "tgcn_mix_weights")
        mode = 0
        if bias_own is not None:
            b["bias"][:self.n_own] = bias_own
            mode = 1
        _lib.check(lib.tgcn_contract_fwd(b["stack"].data_ptr(), b["wmix"].data_ptr(), b["bias"].data_ptr() if mode else None, mode,
                                         b["out"].data_ptr(), b["scr"].data_ptr(), Q, n, D, G, K, self.engine, st), "tgcn_contract_fwd")
        return b["out"][:, :self.n_own]

    def backward(self, dout_own, want_bias=True):
        """dout_own [Q, n_own, G] -> (dW [K, D, G] summed over the ranks, db_own [n_own, G] or None)."""
        from . import _lib
        lib = self.lib
        Q = dout_own.shape[0]
        b = self._buffers(Q)
        K, D, G, n = self.K, self.D, self.G, self.n_ext
        st = torch.cuda.current_stream(self.device).cuda_stream
        b["dout"][:, :self.n_own] = dout_own
        _lib.check(lib.tgcn_contract_bwd_w(b["stack"].data_ptr(), b["dout"].data_ptr(), b["dwmix"].data_ptr(), b["ws"].data_ptr(),
                                           Q, n, D, G, K, self.engine, st), "tgcn_contract_bwd_w")
        dW = torch.empty((K, D, G), dtype=torch.float32, device=self.device)
        _lib.check(lib.tgcn_mix_weights(b["dwmix"].data_ptr(), dW.data_ptr(), K, D * G, self.recursion, 1, st), "tgcn_mix_weights")
        db = None
        if want_bias:
            _lib.check(lib.tgcn_bias_grad(b["dout"].data_ptr(), b["db"].data_ptr(), b["ws"].data_ptr(), Q, n, G, 1, st), "tgcn_bias_grad")
            db = b["db"][:self.n_own]
        if self.world > 1:
            dist.all_reduce(dW, op=dist.ReduceOp.SUM, group=self.group)
        return dW, db
