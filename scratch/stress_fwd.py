import sys, os, torch
sys.path.insert(0, '/root/repo')
from tgcn_b200 import _lib
lib = _lib.load()
def run(shape, reps):
    Q,N,D,G,K = shape
    g = torch.Generator(device="cuda").manual_seed(0)
    stack = torch.randn(K, N, Q*D, device="cuda", generator=g)
    W = torch.randn(K, D, G, device="cuda", generator=g)*0.2
    bias = torch.randn(N, G, device="cuda", generator=g)
    st = torch.cuda.current_stream().cuda_stream
    outs = []
    scr = torch.empty(int(lib.tgcn_contract_fwd_scratch(Q,N,D,G,K))//4+64, device="cuda")
    ref = torch.empty(Q,N,G, device="cuda")
    lib.tgcn_contract_fwd(stack.data_ptr(), W.data_ptr(), bias.data_ptr(), 1, ref.data_ptr(), scr.data_ptr(), Q,N,D,G,K, 1, st)
    nbadruns = 0
    for rep in range(reps):
        out = torch.full((Q,N,G), float("nan"), device="cuda")
        rc = lib.tgcn_contract_fwd(stack.data_ptr(), W.data_ptr(), bias.data_ptr(), 1, out.data_ptr(), scr.data_ptr(), Q,N,D,G,K, 2, st)
        torch.cuda.synchronize()
        err = (out-ref).abs()
        bad = (err > 1e-3*ref.abs().max()).nonzero()
        if bad.shape[0]:
            nbadruns += 1
            m = (bad[:,1]*Q + bad[:,0])
            tiles = (m//128).unique().tolist()
            cols = bad[:,2].unique().tolist()
            rows_in_tile = (m%128).unique().numel()
            print("  rep", rep, "bad tiles", tiles[:8], "cols", cols[:16], "rows/tile", rows_in_tile, "maxerr", float(err.max()))
    print(shape, "bad runs", nbadruns, "/", reps)
for shape in [(8,10464,32,64,10), (8,10464,32,48,10), (8,10464,28,64,10), (8,10464,32,56,10), (8,10464,24,64,10)]:
    run(shape, 30)
