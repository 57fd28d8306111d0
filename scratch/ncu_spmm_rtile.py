"""Two register-tiled SpMM steps (4-row tiles) on the 1M-vertex graph for an ncu capture (-k regex:rtile -s 1 -c 1)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tgcn_b200 import _lib, workloads as wl
from tgcn_b200.csr import build_csr, make_rowtile_plan
lib = _lib.load()
dev = torch.device("cuda")
L, _ = wl.random_geometric()
plan = build_csr(L, dev)
host = plan._host_arrays()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 4
keep = make_rowtile_plan(host[0], host[1], host[2], plan.n, R, plan.col, min_gain=0.0)
N, C = plan.n, 192
stack = torch.randn(2, N, C, device=dev)
st = torch.cuda.current_stream().cuda_stream
for k in range(2):
    assert lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N, stack[0].data_ptr(),
                              None, stack[1].data_ptr(), C, 1.0, 0.0, st) == 0
torch.cuda.synchronize()
print("ok", keep[2])
