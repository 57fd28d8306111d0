"""Time the streaming SpMM step under several kernel variants in ONE process (tgcn_set_tuning)."""
import sys, os, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tgcn_b200 import _lib, workloads as wl
from tgcn_b200.csr import build_csr
lib = _lib.load()
dev = torch.device("cuda")
KEYS = (b"SPMM_PIPE", b"SPMM_TILE", b"SPMM_WARPROW", b"SPMM_CSM", b"SPMM_STAGED")
VARIANTS = [("plain", {}), ("pipe4", {b"SPMM_PIPE": 4}), ("pipe8", {b"SPMM_PIPE": 8}),
            ("csm6", {b"SPMM_CSM": 6}), ("csm8", {b"SPMM_CSM": 8}), ("csm8p2", {b"SPMM_CSM": 28}), ("csm8p8", {b"SPMM_CSM": 88}),
            ("csm8p6", {b"SPMM_CSM": 68})]


def run(name, L, C, K=5, reps=5, has_prev=False):
    plan = build_csr(L, dev)
    N = plan.n
    stack = torch.randn(K, N, C, device=dev)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

    def steps():
        st = torch.cuda.current_stream().cuda_stream
        for k in range(1, K):
            pv = stack[k - 2].data_ptr() if (has_prev and k >= 2) else None
            assert lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N,
                                      stack[k - 1].data_ptr(), pv, stack[k].data_ptr(), C, 2.0 if pv else 1.0,
                                      -1.0 if pv else 0.0, st) == 0
    byt = 2 * 4 * N * C + 8 * plan.nnz + 4 * (N + 1)
    for vname, kv in VARIANTS:
        for k in KEYS:
            lib.tgcn_set_tuning(k, 0)
        for k, v in kv.items():
            lib.tgcn_set_tuning(k, v)
        for _ in range(2):
            steps()
        torch.cuda.synchronize()
        best = 1e30
        tot = 0.0
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); steps(); b.record(); b.synchronize()
            t = a.elapsed_time(b) / (K - 1) * 1e3
            best = min(best, t); tot += t
        us = tot / reps
        print("%-8s %-6s prev=%d N=%d C=%d nnz=%d: %.1f us/step (best %.1f)  %.0f GB/s of 2S+E (%.1f%% of 6540.8)"
              % (name, vname, has_prev, N, C, plan.nnz, us, best, byt / us / 1e3, byt / us / 1e3 / 65.408), flush=True)
    for k in KEYS:
        lib.tgcn_set_tuning(k, -1)


which = sys.argv[1]
if which == "mesh":
    graphs, perm, Ls, n_real = wl.cortical_mesh()
    run("mesh-L1", Ls[0], 240); run("mesh-L2", Ls[2], 256)
else:
    t0 = time.time(); L, pts = wl.random_geometric(); print("rgg built in %.1f s" % (time.time() - t0), flush=True)
    run("rgg1m", L, 192, K=4, reps=3)
