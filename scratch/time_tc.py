import sys, os, torch
sys.path.insert(0, '/root/repo')
from tgcn_b200 import _lib
lib = _lib.load()
def bench(shape, what, reps=20):
    Q,N,D,G,K = shape
    g = torch.Generator(device="cuda").manual_seed(0)
    stack = torch.randn(K, N, Q*D, device="cuda", generator=g)
    W = torch.randn(K, D, G, device="cuda", generator=g)*0.2
    bias = torch.randn(N, G, device="cuda", generator=g)
    dout = torch.randn(Q, N, G, device="cuda", generator=g)
    out = torch.empty(Q,N,G, device="cuda"); dW = torch.empty(K,D,G, device="cuda"); gs = torch.empty(K,N,Q*D, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    scr = torch.empty(int(lib.tgcn_contract_fwd_scratch(Q,N,D,G,K))//4+64, device="cuda")
    ws = torch.empty(int(lib.tgcn_layer_bwd_workspace(Q,N,D,G,K))//4+64, device="cuda")
    flush = torch.empty(64*1024*1024, device="cuda")
    def call():
        if what == "fwd": return lib.tgcn_contract_fwd(stack.data_ptr(), W.data_ptr(), bias.data_ptr(), 1, out.data_ptr(), scr.data_ptr(), Q,N,D,G,K, 2, st)
        if what == "bwd_w": return lib.tgcn_contract_bwd_w(stack.data_ptr(), dout.data_ptr(), dW.data_ptr(), ws.data_ptr(), Q,N,D,G,K, 2, st)
        if what == "bwd_x": return lib.tgcn_contract_bwd_x(dout.data_ptr(), W.data_ptr(), gs.data_ptr(), ws.data_ptr(), Q,N,D,G,K, 2, st)
    for _ in range(3): assert call() == 0, _lib.last_error()
    tot = 0.0
    for r in range(reps):
        flush.fill_(r)
        a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(); b.record(); b.synchronize(); tot += a.elapsed_time(b)
    us = tot/reps*1e3
    byt = 4*(K*N*Q*D + Q*N*G)
    print("%-6s %s R=%s: %.1f us  (%.0f GB/s algorithmic)" % (what, shape, os.environ.get("TGCN_T2_R","max"), us, byt/us/1e3))
L1=(8,41856,30,32,10); L2=(8,10464,32,64,10)
for w in sys.argv[1:]:
    bench(L1, w); bench(L2, w)
