import sys, os, torch, numpy as np
sys.path.insert(0, '/root/repo')
from tgcn_b200 import _lib, workloads as wl
from tgcn_b200.csr import build_csr
lib = _lib.load()
graphs, perm, Ls, n_real = wl.hcp_parcellation()
dev = torch.device("cuda")
def run(name, L, Q, D, G, K, bias_mode, need_dx, pool_p=4, reps=30):
    plan = build_csr(L, dev)
    N = plan.n
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(Q, N, D, device="cuda", generator=g)
    W = torch.randn(K, D, G, device="cuda", generator=g) * 0.1
    bias = torch.randn(N if bias_mode == 1 else 1, G, device="cuda", generator=g)
    y = torch.empty(Q, N // pool_p, G, device="cuda"); idx = torch.empty(Q, N // pool_p, G, dtype=torch.uint8, device="cuda")
    stack = torch.empty(int(lib.tgcn_resident_stack_bytes(Q, N, D, K)) // 4, device="cuda")
    dy = torch.randn(Q, N // pool_p, G, device="cuda", generator=g)
    dW = torch.empty(K, D, G, device="cuda"); db = torch.empty(N if bias_mode == 1 else 1, G, device="cuda")
    dx = torch.empty(Q, N, D, device="cuda") if need_dx else None
    wimg = torch.empty(int(lib.tgcn_resident_weights_bytes(D, G, K)) // 4, device="cuda")
    ws = torch.empty(int(lib.tgcn_resident_bwd_workspace(Q, N, D, G, K)) // 4, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ri, en, E = plan.packed(lib.tgcn_resident_pack_classes(Q, N, D, 0))
    rit, ent, Et = plan.packed(lib.tgcn_resident_pack_classes(Q, N, D, 1), transpose=True)
    def fwd():
        return lib.tgcn_resident_layer_fwd(ri.data_ptr(), en.data_ptr(), N, E, x.data_ptr(), W.data_ptr(),
            bias.data_ptr(), bias_mode, None, y.data_ptr(), idx.data_ptr(), pool_p, 1, stack.data_ptr(), wimg.data_ptr(), Q, D, G, K, 0, st)
    def bwd():
        return lib.tgcn_resident_layer_bwd(rit.data_ptr(), ent.data_ptr(), N, Et, None, dy.data_ptr(),
            idx.data_ptr(), y.data_ptr(), pool_p, 1, stack.data_ptr(), wimg.data_ptr(), dW.data_ptr(), db.data_ptr(), bias_mode,
            None if dx is None else dx.data_ptr(), ws.data_ptr(), Q, D, G, K, 0, st)
    for fn, nm in ((fwd, "fwd"), (bwd, "bwd")):
        for _ in range(3): assert fn() == 0, _lib.last_error()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); b.synchronize()
        print("%-10s %s Q=%d N=%d D=%d G=%d nnz=%d: %.1f us" % (name, nm, Q, N, D, G, plan.nnz, a.elapsed_time(b) / reps * 1e3), flush=True)
Q = int(os.environ.get("Q", "64"))
run("hcp-L1", Ls[0], Q, 15, 32, int(os.environ.get("KK", "10")), 1, False)
run("hcp-L2", Ls[2], Q, 32, 64, int(os.environ.get("KK", "10")), 2, True)
g2, p2, Lm, nr = wl.mnist_grid()
run("mnist-L1", Lm[0], 100, 12, 15, 10, 1, False, pool_p=2)
