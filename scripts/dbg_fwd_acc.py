#!/usr/bin/env python
"""Accuracy of the tcgen05 forward contraction against float64 (GPU einsum) at the mesh layer-1 shape; run with
TGCN_T3_NI=1|2|4 to compare the issuer configurations (the epilogue's summation order differs, nothing else)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tgcn_b200 import _lib  # noqa: E402

lib = _lib.load()
torch.manual_seed(0)
for (Q, N, D, G, K) in ((8, 11648, 32, 32, 10), (8, 41856, 32, 32, 10), (8, 2912, 32, 64, 10)):
    stack = torch.randn(K, N, Q * D, device="cuda")
    W = torch.randn(K, D, G, device="cuda") * 0.1
    bias = torch.randn(N, G, device="cuda")
    out = torch.empty(Q, N, G, device="cuda")
    scr = torch.empty(max(int(lib.tgcn_contract_fwd_scratch(Q, N, D, G, K)), 16) // 4 + 64, device="cuda")
    rc = lib.tgcn_contract_fwd(stack.data_ptr(), W.data_ptr(), bias.data_ptr(), 1, out.data_ptr(), scr.data_ptr(), Q, N, D, G, K, 2,
                               torch.cuda.current_stream().cuda_stream)
    assert rc == 0, _lib.last_error()
    ref = torch.einsum("knqd,kdg->qng", stack.double().view(K, N, Q, D), W.double()) + bias.double()[None]
    err = (out.double() - ref).abs()
    print("NI=%s Q=%d N=%d G=%d  max err / max ref = %.3e   rms err / rms ref = %.3e   mean signed err / rms ref = %.3e" % (
        os.environ.get("TGCN_T3_NI", "default"), Q, N, G, float(err.max() / ref.abs().max()),
        float((out.double() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()),
        float((out.double() - ref).mean() / ref.pow(2).mean().sqrt())), flush=True)
