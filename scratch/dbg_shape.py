import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from tgcn_b200 import _lib
lib = _lib.load()
def run(shape):
    Q,N,D,G,K = shape
    g = torch.Generator(device="cuda").manual_seed(0)
    stack = torch.randn(K, N, Q*D, device="cuda", generator=g)
    W = torch.randn(K, D, G, device="cuda", generator=g)*0.2
    bias = torch.randn(N, G, device="cuda", generator=g)
    dout = torch.randn(Q, N, G, device="cuda", generator=g)
    st = torch.cuda.current_stream().cuda_stream
    ref = torch.zeros(Q,N,G, device="cuda", dtype=torch.float64)
    for j in range(K):
        ref += torch.einsum("nqd,dg->qng", stack[j].double().reshape(N,Q,D), W[j].double())
    ref += bias.double()[None]
    for eng in (1,2):
        out = torch.full((Q,N,G), float("nan"), device="cuda")
        scr = torch.empty(int(lib.tgcn_contract_fwd_scratch(Q,N,D,G,K))//4+64, device="cuda")
        rc = lib.tgcn_contract_fwd(stack.data_ptr(), W.data_ptr(), bias.data_ptr(), 1, out.data_ptr(), scr.data_ptr(), Q,N,D,G,K, eng, st)
        torch.cuda.synchronize()
        err = (out.double()-ref).abs()
        rel = float(err.max()/ref.abs().max())
        bad = (err > 1e-3*ref.abs().max()).nonzero()
        print(shape, "eng", eng, "rc", rc, "rel", rel, "nbad", bad.shape[0], bad[:5].tolist() if bad.shape[0] else "")
        if bad.shape[0]:
            # which m rows (n*Q+q) are bad
            m = (bad[:,1]*Q + bad[:,0]).unique()
            print("  bad tiles:", (m//128).unique()[:40].tolist(), "count", (m//128).unique().numel())
for shape in [(8,10388,32,64,10), (8,10464,32,64,10), (8,41856,30,32,10), (64,96,32,64,10), (8, 2000, 32, 64, 10)]:
    run(shape)
