"""A few SpMM steps on the 1M-vertex graph (default kernel selection) for an ncu capture."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tgcn_b200 import _lib, workloads as wl
from tgcn_b200.csr import build_csr
lib = _lib.load()
dev = torch.device("cuda")
L, _ = wl.random_geometric()
plan = build_csr(L, dev)
N, C = plan.n, 192
stack = torch.randn(3, N, C, device=dev)
st = torch.cuda.current_stream().cuda_stream
for k in range(4):
    assert lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N, stack[k % 2].data_ptr(),
                              None, stack[2].data_ptr(), C, 1.0, 0.0, st) == 0
torch.cuda.synchronize()
print("ok")
