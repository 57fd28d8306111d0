#!/usr/bin/env python
"""Device timings of the streaming-path kernels at the BASELINE shapes, straight through the C-ABI (CUDA events around
CUDA-graph replays, L2 flushed between replays).  Usage:
    python scripts/time_kernels.py [contract] [spmm] [head] [--shapes mesh1,mesh2,rgg] [--reps 10]
Prints one line per kernel: microseconds and algorithmic GB/s (bytes as in DESIGN.md section 4)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from tgcn_b200 import _lib  # noqa: E402

SHAPES = {  # Q, N, D, G, K
    "mesh1": (8, 41856, 32, 32, 10),       # cortical mesh layer 1 (D = 30 padded to 32)
    "mesh1raw": (8, 41856, 30, 32, 10),    # unpadded (second-generation kernels)
    "mesh2": (8, 10464, 32, 64, 10),
    "rgg": (1, 1000000, 192, 64, 8),
    "rgg200k": (1, 200000, 192, 64, 8),
}


PLAIN = False     # --plain: direct launches instead of CUDA-graph replays (for ncu: -k <kernel> -s 2 -c 1)


def time_graph(fn, reps, flush):
    if PLAIN:
        tot = 0.0
        for r in range(3 + reps):
            flush.fill_(float(r))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); b.synchronize()
            tot += a.elapsed_time(b) if r >= 3 else 0.0
        return tot / reps * 1e3
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    tot = 0.0
    for r in range(reps):
        flush.fill_(float(r))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3        # us


def contract(lib, name, reps, flush):
    Q, N, D, G, K = SHAPES[name]
    dev = "cuda"
    stack = torch.randn(K, N, Q * D, device=dev)
    W = torch.randn(K, D, G, device=dev) * 0.1
    bias = torch.randn(N, G, device=dev)
    out = torch.empty(Q, N, G, device=dev)
    dout = torch.randn(Q, N, G, device=dev)
    dW = torch.empty(K, D, G, device=dev)
    gs = torch.empty(K, N, Q * D, device=dev)
    scr = torch.empty(max(int(lib.tgcn_contract_fwd_scratch(Q, N, D, G, K)), 16) // 4 + 64, device=dev)
    ws = torch.empty(max(int(lib.tgcn_layer_bwd_workspace(Q, N, D, G, K)), 16) // 4 + 64, device=dev)
    S, Y = 4.0 * N * Q * D, 4.0 * Q * N * G
    st = lambda: torch.cuda.current_stream().cuda_stream   # noqa: E731

    def fwd():
        assert lib.tgcn_contract_fwd(stack.data_ptr(), W.data_ptr(), bias.data_ptr(), 1, out.data_ptr(), scr.data_ptr(), Q, N, D, G, K, 2, st()) == 0, _lib.last_error()

    def bwd_w():
        assert lib.tgcn_contract_bwd_w(stack.data_ptr(), dout.data_ptr(), dW.data_ptr(), ws.data_ptr(), Q, N, D, G, K, 2, st()) == 0, _lib.last_error()

    def bwd_x():
        assert lib.tgcn_contract_bwd_x(dout.data_ptr(), W.data_ptr(), gs.data_ptr(), ws.data_ptr(), Q, N, D, G, K, 2, st()) == 0, _lib.last_error()
    for label, fn, nbytes in (("fwd", fwd, K * S + Y), ("bwd_w", bwd_w, K * S + Y), ("bwd_x", bwd_x, K * S + Y)):
        us = time_graph(fn, reps, flush)
        print("contract %-8s %-6s Q=%d N=%d D=%d G=%d K=%d  %8.1f us  %7.0f GB/s" % (name, label, Q, N, D, G, K, us, nbytes / us / 1e3), flush=True)


def spmm(lib, name, reps, flush, rowtile, mode=-1):
    from tgcn_b200 import workloads as wl
    from tgcn_b200.csr import build_csr
    Q, N, D, G, K = SHAPES[name]
    if name.startswith("mesh"):
        Ls = wl.cortical_mesh()[2]
        L = Ls[0] if name.startswith("mesh1") else Ls[2]
    else:
        L = wl.random_geometric(n=N)[0]
    plan = build_csr(L, torch.device("cuda"))
    if rowtile:
        plan.ensure_rowtile_plans(rows_per_tile=rowtile)
    C = Q * D
    a = torch.randn(N, C, device="cuda")
    b = torch.empty(N, C, device="cuda")
    st = lambda: torch.cuda.current_stream().cuda_stream   # noqa: E731

    def step():
        assert lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N, a.data_ptr(), None, b.data_ptr(), C, 1.0, 0.0, st()) == 0, _lib.last_error()
        assert lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N, b.data_ptr(), None, a.data_ptr(), C, 1.0, 0.0, st()) == 0
    lib.tgcn_set_tuning(b"SPMM_RTILE", mode)
    us = time_graph(step, reps, flush) / 2

    def chain():          # the recursion as the layer runs it: 8 dependent steps, each reading what the last one wrote
        for _ in range(4):
            step()
    us_chain = time_graph(chain, reps, flush) / 8
    lib.tgcn_set_tuning(b"SPMM_RTILE", -1)
    nbytes = 2 * 4.0 * N * C + 8.0 * plan.nnz + 4.0 * (N + 1)
    print("spmm     %-8s rowtile=%d mode=%d chain-of-8 %8.1f us/step  %7.0f GB/s (%.1f %%)" % (name, rowtile, mode, us_chain, nbytes / us_chain / 1e3, nbytes / us_chain / 1e3 / 65.408))
    print("spmm     %-8s rowtile=%d N=%d C=%d nnz=%d  %8.1f us  %7.0f GB/s  (%.1f %% of 6540.8)" % (name, rowtile, N, C, plan.nnz, us, nbytes / us / 1e3, nbytes / us / 1e3 / 65.408), flush=True)


def head(lib, reps, flush):
    import torch.nn as nn
    import torch.nn.functional as F
    from tgcn_b200.nn.head import Fc1FusedSGD, fused_head
    Q, I, Hd, C = 8, 167424, 200, 6
    mods = [nn.Linear(I, Hd).cuda(), nn.BatchNorm1d(Hd).cuda(), nn.Linear(Hd, C).cuda()]
    for m in mods:
        m.train()
    x = torch.randn(Q, I, device="cuda", requires_grad=True)
    y = torch.randint(0, C, (Q,), device="cuda")
    upd = Fc1FusedSGD(mods[0].weight, lr=0.01, momentum=0.5)

    def fwd_bwd():
        lp = fused_head(x, *mods, fc1_update=upd)
        F.nll_loss(lp, y).backward()
    us = time_graph(fwd_bwd, reps, flush)
    print("head     fwd+bwd+fused update Q=%d I=%d Hd=%d  %8.1f us  (weight bytes touched 5 x %.0f MB -> %.0f GB/s)" % (Q, I, Hd, us, I * Hd * 4 / 1e6, 5 * I * Hd * 4 / us / 1e3), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="*", default=["contract", "spmm", "head"])
    ap.add_argument("--shapes", default="mesh1,mesh2")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--rowtile", default="0,4")
    ap.add_argument("--rtmodes", default="-1", help="SPMM_RTILE tuning values to time (1 one-shot, 2/3 persistent)")
    ap.add_argument("--plain", action="store_true", help="direct launches, no CUDA graph (use under ncu)")
    args = ap.parse_args()
    global PLAIN
    PLAIN = args.plain
    lib = _lib.load()
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    for name in args.shapes.split(","):
        if "contract" in args.what:
            contract(lib, name, args.reps, flush)
        if "spmm" in args.what:
            for rt in args.rowtile.split(","):
                for mode in ([int(m) for m in args.rtmodes.split(",")] if int(rt) else [-1]):
                    spmm(lib, name, args.reps, flush, int(rt), mode)
    if "head" in args.what:
        head(lib, args.reps, flush)


if __name__ == "__main__":
    main()
