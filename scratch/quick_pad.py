"""Shortest possible A/B of the padded row-tile plan and the SM-contiguous mapping on the 1M-vertex graph."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tgcn_b200 import _lib, workloads as wl
from tgcn_b200.csr import build_csr, make_rowtile_plan
lib = _lib.load(); dev = torch.device("cuda")
L, _ = wl.random_geometric()
base = build_csr(L, dev); host = base._host_arrays(); N, C = base.n, 192
x = torch.randn(2, N, C, device=dev)
st = torch.cuda.current_stream().cuda_stream
ref = None
for pad in (1, 8):
    p = build_csr(L, dev)
    keep = make_rowtile_plan(host[0], host[1], host[2], N, 4, p.col, min_gain=0.0, pad=pad)
    for mode in (1, 16):
        lib.tgcn_set_tuning(b"SPMM_RTILE", mode)
        f = lambda: lib.tgcn_spmm_step(p.rowptr.data_ptr(), p.col.data_ptr(), p.val.data_ptr(), N, x[0].data_ptr(), None, x[1].data_ptr(), C, 1.0, 0.0, st)
        f(); torch.cuda.synchronize()
        if ref is None: ref = x[1].clone()
        err = float((x[1] - ref).abs().max())
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); [f() for _ in range(4)]; b.record(); b.synchronize()
        print("pad=%d mode=%d: %.1f us/step  maxabs diff vs pad1/mode1 %.1e" % (pad, mode, a.elapsed_time(b) / 4 * 1e3, err), flush=True)
