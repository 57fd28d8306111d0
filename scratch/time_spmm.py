import sys, os, time, torch, numpy as np
sys.path.insert(0, '/root/repo')
from tgcn_b200 import _lib, workloads as wl
from tgcn_b200.csr import build_csr
lib = _lib.load()
dev = torch.device("cuda")
def run(name, L, C, K=5, reps=5):
    plan = build_csr(L, dev)
    N = plan.n
    if os.environ.get("STAGED"):
        info = plan.ensure_block_plans(rows_per_block=int(os.environ["STAGED"]), cap=int(os.environ.get("CAP", "65534")))
        print("block plans:", [i[2] for i in info], flush=True)
    stack = torch.randn(K, N, C, device=dev)
    def steps():
        st = torch.cuda.current_stream().cuda_stream
        for k in range(1, K):
            assert lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N, stack[k-1].data_ptr(), None, stack[k].data_ptr(), C, 1.0, 0.0, st) == 0
    for _ in range(2): steps()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): steps()
    b.record(); b.synchronize()
    us = a.elapsed_time(b) / (reps * (K - 1)) * 1e3
    byt = 2 * 4 * N * C + 8 * plan.nnz + 4 * (N + 1)
    print("%-10s tile=%s staged=%s N=%d C=%d nnz=%d: %.1f us/step  %.0f GB/s algorithmic (%.1f%% of 6540.8)" % (name, os.environ.get("TGCN_SPMM_TILE", "0"), os.environ.get("STAGED", "0"), N, C, plan.nnz, us, byt / us / 1e3, byt / us / 1e3 / 65.408), flush=True)
which = sys.argv[1]
if which == "mesh":
    graphs, perm, Ls, n_real = wl.cortical_mesh()
    run("mesh-L1", Ls[0], 240); run("mesh-L2", Ls[2], 256)
else:
    t0 = time.time(); L, pts = wl.random_geometric(); print("rgg built in %.1f s" % (time.time() - t0), flush=True)
    run("rgg1m", L, 192, K=4, reps=3)
