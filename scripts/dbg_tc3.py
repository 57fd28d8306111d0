import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, torch
sys.path.insert(0, %r)
from tgcn_b200 import _lib
lib = _lib.load()
Q, N, D, G, K = 8, int(sys.argv[1]), 32, int(sys.argv[2]), 10
stack = torch.randn(K, N, Q * D, device="cuda"); W = torch.randn(K, D, G, device="cuda") * 0.1
bias = torch.randn(N, G, device="cuda"); out = torch.empty(Q, N, G, device="cuda")
scr = torch.empty(max(int(lib.tgcn_contract_fwd_scratch(Q, N, D, G, K)), 16) // 4 + 64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for it in range(int(sys.argv[3])):
    rc = lib.tgcn_contract_fwd(stack.data_ptr(), W.data_ptr(), bias.data_ptr(), 1, out.data_ptr(), scr.data_ptr(), Q, N, D, G, K, 2, st)
    assert rc == 0, _lib.last_error()
torch.cuda.synchronize()
ref = torch.einsum("jnqd,jdg->qng", stack.double().reshape(K, N, Q, D), W.double()) + bias.double()[None]
print("ok err %%.2e" %% float((out.double() - ref).abs().max() / ref.abs().max()))
''' % ROOT
for ni in (4, 3):
    for n, g, reps in ((5000, 32, 1), (12000, 32, 1), (20000, 32, 1), (41856, 32, 1), (41856, 32, 3), (10464, 64, 3)):
        env = dict(os.environ, TGCN_T3_NI=str(ni))
        r = subprocess.run([sys.executable, "-c", CODE, str(n), str(g), str(reps)], env=env, capture_output=True, text=True, timeout=300)
        tail = (r.stdout.strip().splitlines() or ["-"])[-1] if r.returncode == 0 else [l for l in r.stderr.splitlines() if "rror" in l][-1:]
        print("NI=%d N=%d G=%d reps=%d -> rc=%d %s" % (ni, n, g, reps, r.returncode, tail), flush=True)
