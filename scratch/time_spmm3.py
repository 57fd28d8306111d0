"""Register-tiled SpMM (SPMM_RTILE, row-tile plans) against the default per-entry kernels, one process.
usage: time_spmm3.py [rgg|mesh|both] [n_rgg]"""
import sys, os, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tgcn_b200 import _lib, workloads as wl
from tgcn_b200.csr import build_csr, make_rowtile_plan
lib = _lib.load()
dev = torch.device("cuda")
MODES = (4, 5, 6, 8)


def run(name, L, C, K=4, reps=3, has_prev=False):
    base = build_csr(L, dev)
    N = base.n
    host = base._host_arrays()
    ops = {"default": base}
    for R in (4, 8):
        t0 = time.time()
        p = build_csr(L, dev)
        made = make_rowtile_plan(host[0], host[1], host[2], N, R, p.col, min_gain=0.0)
        p._rowtiles = [made]
        ops["rtile%d" % R] = p
        print("%s: row-tile plan R=%d built in %.1f s, gain %.2f, %d sources" % (name, R, time.time() - t0, made[2]["gain"], made[2]["sources"]), flush=True)
    for R, pad in ((4, 8), (4, 4), (8, 4)):             # padded plans: no one-at-a-time tail loop in the kernel
        p = build_csr(L, dev)
        p._rowtiles = [make_rowtile_plan(host[0], host[1], host[2], N, R, p.col, min_gain=0.0, pad=pad)]
        ops["rtile%dp%d" % (R, pad)] = p
    stack = torch.randn(K, N, C, device=dev)
    byt = 2 * 4 * N * C + 8 * base.nnz + 4 * (N + 1)
    outs = {}
    runs = [("default", ops["default"], 1)]
    for m in MODES:
        runs += [("rtile4/b%d" % m, ops["rtile4"], m)]
    runs += [("rtile8/b3", ops["rtile8"], 1), ("rtile8/b4", ops["rtile8"], 4)]
    runs += [(k, v, 1) for k, v in ops.items() if "p" in k]
    runs += [("rtile4/remap", ops["rtile4"], 16), ("rtile8/remap", ops["rtile8"], 16), ("rtile4p8/remap", ops["rtile4p8"], 16)]
    for vname, plan, mode in runs:
        lib.tgcn_set_tuning(b"SPMM_RTILE", mode)

        def steps():
            st = torch.cuda.current_stream().cuda_stream
            for k in range(1, K):
                pv = stack[k - 2].data_ptr() if (has_prev and k >= 2) else None
                assert lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N,
                                          stack[k - 1].data_ptr(), pv, stack[k].data_ptr(), C, 2.0 if pv else 1.0,
                                          -1.0 if pv else 0.0, st) == 0, _lib.last_error()
        for _ in range(2):
            steps()
        torch.cuda.synchronize()
        outs[vname] = stack[1].clone()
        best, tot = 1e30, 0.0
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); steps(); b.record(); b.synchronize()
            t = a.elapsed_time(b) / (K - 1) * 1e3
            best = min(best, t); tot += t
        us = tot / reps
        err = float((outs[vname] - outs["default"]).abs().max() / outs["default"].abs().max())
        print("%-8s %-8s prev=%d N=%d C=%d nnz=%d: %.1f us/step (best %.1f)  %.0f GB/s of 2S+E (%.1f%% of 6540.8)  max rel diff vs default %.1e"
              % (name, vname, has_prev, N, C, base.nnz, us, best, byt / us / 1e3, byt / us / 1e3 / 65.408, err), flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("mesh", "both"):
    graphs, perm, Ls, n_real = wl.cortical_mesh()
    run("mesh-L1", Ls[0], 240, K=5, reps=5); run("mesh-L2", Ls[2], 256, K=5, reps=5)
if which in ("rgg", "both"):
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
    L, pts = wl.random_geometric(n=n)
    run("rgg%dk" % (n // 1000), L, 192, K=4, reps=3)
