"""Rescaled-Laplacian operand: whatever the reference layers accept -> int32 CSR on the device.

The reference keeps `L` as a plain attribute and re-uploads the dense N x N tensor on every
forward (`L = self.L.to(X.device)`, tgcn/nn/gcn.py:141,223).  Here `L` is converted ONCE per
device into two CSR triplets (L for the forward recursion, L^T for the adjoint recursion).
"""
import threading

import numpy as np
import torch


class LaplacianCSR:
    """int32 CSR of L and of L^T resident on one CUDA device."""

    __slots__ = ("n", "nnz", "rowptr", "col", "val", "rowptr_t", "col_t", "val_t", "symmetric", "device",
                 "_host", "_packed", "_lock", "_rowtiles")

    def __init__(self, n, rowptr, col, val, rowptr_t, col_t, val_t, symmetric, device):
        self.n, self.nnz = n, int(col.numel())
        self.rowptr, self.col, self.val = rowptr, col, val
        self.rowptr_t, self.col_t, self.val_t = rowptr_t, col_t, val_t
        self.symmetric = symmetric
        self.device = device
        self._host = None          # (crow, col, val, crow_t, col_t, val_t) as int32/float32 numpy, for packing
        self._packed = {}
        self._lock = threading.Lock()
        self._rowtiles = None      # row-tile plans (register-tiled SpMM): list of (handle, tensors kept alive, stats)

    def _host_arrays(self):
        if self._host is None:
            self._host = tuple(np.ascontiguousarray(t.cpu().numpy()) for t in
                               (self.rowptr, self.col, self.val, self.rowptr_t, self.col_t, self.val_t))
        return self._host

    def ensure_rowtile_plans(self, rows_per_tile=8, min_gain=1.5, pad=1):
        """Row-tile plans for the register-tiled SpMM kernel (include/tgcn_b200.h, tgcn_rowtile_plan_create): built
        once per device for L (and L^T when it differs) and registered with the library when the row order has enough
        locality (see `make_rowtile_plan`).  Idempotent."""
        if self._rowtiles is not None:
            return self._rowtiles
        with self._lock:
            if self._rowtiles is not None:
                return self._rowtiles
            host = self._host_arrays()
            made = []
            variants = [(host[0], host[1], host[2], self.col)]
            if not self.symmetric:
                variants.append((host[3], host[4], host[5], self.col_t))
            for rp, c, v, col_dev in variants:
                one = make_rowtile_plan(rp, c, v, self.n, rows_per_tile, col_dev, min_gain, pad)
                if one is not None:
                    made.append(one)
            self._rowtiles = made
        return self._rowtiles

    def __del__(self):
        try:
            if self._rowtiles:
                from . import _lib
                lib = _lib.load()
                for h, _, _ in self._rowtiles:
                    lib.tgcn_rowtile_plan_destroy(h)
        except Exception:
            pass

    def packed(self, classes, transpose=False):
        """Packed, bank-ordered CSR for the sample-resident kernels (include/tgcn_b200.h,
        tgcn_pack_csr_host): (rowinfo int32 [N,2], entries int32 [E,2], E) on this device, cached."""
        if self.symmetric:
            transpose = False
        key = (int(classes), bool(transpose))
        hit = self._packed.get(key)
        if hit is not None:
            return hit
        with self._lock:
            hit = self._packed.get(key)
            if hit is not None:
                return hit
            from . import _lib
            lib = _lib.load()
            host = self._host_arrays()
            rp, c, v = host[3:] if transpose else host[:3]
            rowinfo = np.zeros((max((self.n + 1) & ~1, 2), 2), dtype=np.int32)
            E = int(lib.tgcn_pack_csr_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, self.n, key[0],
                                           rowinfo.ctypes.data, None))
            if E < 0:
                raise RuntimeError("tgcn_pack_csr_host rejected the CSR operand")
            entries = np.zeros((max(E, 2), 2), dtype=np.int32)
            lib.tgcn_pack_csr_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, self.n, key[0], rowinfo.ctypes.data,
                                   entries.ctypes.data)
            hit = (torch.from_numpy(rowinfo).to(self.device), torch.from_numpy(entries).to(self.device), E)
            self._packed[key] = hit
        return hit


def make_rowtile_plan(rowptr, col, val, n, rows_per_tile, col_dev, min_gain=1.5, pad=1):
    """Build (host, tgcn_rowtile_plan_host) and register (tgcn_rowtile_plan_create) the row-tile plan of the CSR
    operand `rowptr/col/val` (int32/int32/float32 numpy, `n` rows) whose column array lives on the device as
    `col_dev`.  Returns (handle, device arrays to keep alive, stats), or None when the row order has too little
    locality: CSR entries / distinct (tile, source row) pairs < min_gain -- below that the kernel would do as many
    gathers as the per-entry kernels and R times their multiply-adds.  pad > 1 pads every tile to a multiple of `pad`
    entries (zero coefficients) so the kernel never runs its one-at-a-time tail loop: measured -2.4 % at pad=8 on the
    1M-vertex graph (profiles/r01/spmm_variants.txt); the recorded bench lines use pad=1.  The gain is computed on the
    unpadded count."""
    from . import _lib
    lib = _lib.load()
    R = int(rows_per_tile)
    rp = np.ascontiguousarray(rowptr, dtype=np.int32)
    c = np.ascontiguousarray(col, dtype=np.int32)
    v = np.ascontiguousarray(val, dtype=np.float32)
    if n < 1 or c.size == 0:
        return None
    tile_ptr = np.zeros((n + R - 1) // R + 1, dtype=np.int32)
    distinct = int(lib.tgcn_rowtile_plan_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, n, R, 1, tile_ptr.ctypes.data, None, None))
    if distinct <= 0:
        raise RuntimeError("tgcn_rowtile_plan_host rejected the CSR operand (rows_per_tile=%d)" % R)
    if c.size / distinct < min_gain:
        return None
    pad = int(pad)
    total = distinct if pad == 1 else int(lib.tgcn_rowtile_plan_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, n, R, pad,
                                                                     tile_ptr.ctypes.data, None, None))
    if total <= 0:
        raise RuntimeError("tgcn_rowtile_plan_host rejected pad=%d" % pad)
    src = np.zeros(total, dtype=np.int32)
    w = np.zeros((total, R), dtype=np.float32)
    lib.tgcn_rowtile_plan_host(rp.ctypes.data, c.ctypes.data, v.ctypes.data, n, R, pad, tile_ptr.ctypes.data, src.ctypes.data,
                               w.ctypes.data)
    dev_arrays = tuple(torch.from_numpy(a).to(col_dev.device) for a in (tile_ptr, src, w))
    h = int(lib.tgcn_rowtile_plan_create(col_dev.data_ptr(), n, int(src.max()) + 1, R, dev_arrays[0].data_ptr(),
                                         dev_arrays[1].data_ptr(), dev_arrays[2].data_ptr()))
    if h < 0:
        raise RuntimeError("tgcn_rowtile_plan_create failed: %s" % _lib.last_error())
    return h, dev_arrays, {"gain": c.size / distinct, "sources": distinct, "rows_per_tile": R, "pad": pad, "entries": total}


def _to_scipy_like(L):
    """Return (crow int64, col int64, val float32, n) as CPU torch tensors, rows sorted by column."""
    if isinstance(L, torch.Tensor):
        if L.layout == torch.strided:
            if L.dim() != 2 or L.shape[0] != L.shape[1]:
                raise ValueError("L must be a square matrix, got shape %s" % (tuple(L.shape),))
            csr = L.detach().to(torch.float32).to_sparse_csr()
        elif L.layout == torch.sparse_coo:
            csr = L.detach().to(torch.float32).coalesce().to_sparse_csr()
        elif L.layout == torch.sparse_csr:
            csr = L.detach().to(torch.float32)
        else:
            raise TypeError("unsupported tensor layout for L: %s" % L.layout)
        n = csr.shape[0]
        if csr.shape[0] != csr.shape[1]:
            raise ValueError("L must be square")
        return csr.crow_indices().cpu().to(torch.int64), csr.col_indices().cpu().to(torch.int64), csr.values().cpu(), n
    # scipy sparse / numpy
    try:
        import scipy.sparse as sp
    except ImportError:  # pragma: no cover
        sp = None
    if sp is not None and sp.issparse(L):
        m = L.tocsr().astype(np.float32)
        m.sort_indices()
        m.sum_duplicates()
        return (torch.from_numpy(m.indptr.astype(np.int64)), torch.from_numpy(m.indices.astype(np.int64)),
                torch.from_numpy(m.data.copy()), m.shape[0])
    arr = np.asarray(L, dtype=np.float32)
    return _to_scipy_like(torch.from_numpy(arr))


def _transpose_csr(crow, col, val, n):
    nnz = col.numel()
    rows = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), crow[1:] - crow[:-1])
    # sort by (col, row): stable sort on col keeps rows ascending inside each transposed row
    order = torch.argsort(col, stable=True)
    col_t = rows[order]
    val_t = val[order]
    counts = torch.bincount(col, minlength=n) if nnz else torch.zeros(n, dtype=torch.int64)
    crow_t = torch.zeros(n + 1, dtype=torch.int64)
    crow_t[1:] = torch.cumsum(counts, 0)
    return crow_t, col_t, val_t


def build_csr(L, device):
    crow, col, val, n = _to_scipy_like(L)
    if col.numel() >= 2 ** 31 - 1:
        raise ValueError("nnz(L) does not fit int32")
    crow_t, col_t, val_t = _transpose_csr(crow, col, val, n)
    symmetric = bool(torch.equal(crow, crow_t) and torch.equal(col, col_t) and torch.equal(val, val_t))
    dev = torch.device(device)
    i32 = lambda t: t.to(torch.int32).to(dev)
    rowptr, c, v = i32(crow), i32(col), val.to(dev)
    if symmetric:
        rowptr_t, c_t, v_t = rowptr, c, v
    else:
        rowptr_t, c_t, v_t = i32(crow_t), i32(col_t), val_t.to(dev)
    return LaplacianCSR(n, rowptr, c, v, rowptr_t, c_t, v_t, symmetric, dev)


def build_csr_from_coo(row, col, val, n, device):
    """CSR operand from a COO edge list WITHOUT leaving the device (edge-index operators whose graph changes per
    batch, reference pygeo_hcp.py:272-316): entries sorted by (row, col), duplicates summed (what the reference's
    scatter_add does; more than two duplicates of one entry are added in atomic order), plus the transpose."""
    dev = torch.device(device)
    row = row.to(dev, torch.int64).reshape(-1)
    col = col.to(dev, torch.int64).reshape(-1)
    val = val.to(dev, torch.float32).reshape(-1)
    if row.numel() >= 2 ** 31 - 1:
        raise ValueError("nnz(L) does not fit int32")

    def csr_of(r, c, v):
        key, order = torch.sort(r * n + c, stable=True)
        uniq, inverse = torch.unique_consecutive(key, return_inverse=True)
        vals = torch.zeros(uniq.numel(), dtype=torch.float32, device=dev).index_add_(0, inverse, v[order])
        rr = torch.div(uniq, n, rounding_mode="floor")
        cc = uniq - rr * n
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        if uniq.numel():
            rowptr[1:] = torch.cumsum(torch.bincount(rr, minlength=n), 0)
        return rowptr.to(torch.int32), cc.to(torch.int32), vals, rr

    rowptr, c, v, rr = csr_of(row, col, val)
    rowptr_t, c_t, v_t, _ = csr_of(c.to(torch.int64), rr, v)
    symmetric = bool(torch.equal(rowptr, rowptr_t) and torch.equal(c, c_t) and torch.equal(v, v_t))
    if symmetric:
        rowptr_t, c_t, v_t = rowptr, c, v
    return LaplacianCSR(n, rowptr, c, v, rowptr_t, c_t, v_t, symmetric, dev)


class CSRCache:
    """Per-device cache of the CSR operand; shared by DataParallel replicas (thread-safe).  Pickling / deep-copying
    a module drops the cached plans (they are rebuilt on first use) -- locks and device plans do not travel."""

    def __init__(self):
        self._plans = {}
        self._lock = threading.Lock()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self._plans = {}
        self._lock = threading.Lock()

    def __deepcopy__(self, memo):
        return CSRCache()

    def get(self, L, device):
        key = (device.type, device.index)
        plan = self._plans.get(key)
        if plan is None:
            with self._lock:
                plan = self._plans.get(key)
                if plan is None:
                    plan = build_csr(L, device)
                    self._plans[key] = plan
        return plan

    def clear(self):
        with self._lock:
            self._plans.clear()
