#!/usr/bin/env python
"""BASELINE.json configs[4]: Chebyshev order / time window / batch sweep for the SpMM-vs-contraction crossover.

For every point it times, on one B200 and through the C-ABI (CUDA events, CUDA-graph replay, L2 flushed):
  * basis   : tgcn_cheb_basis      (K-1 SpMM steps + the layout transposes)         -- algorithmic bytes (3K-4)S+(K-1)E
  * contract: tgcn_contract_fwd    (tcgen05 3xTF32 when the tiles cover the shape)  -- flops 2 Q N K D G
  * resident: tgcn_resident_layer_fwd (whole layer in one launch) when the per-sample slab fits in shared memory
and writes one CSV line per point.   python scripts/sweep.py > profiles/r01/sweep.csv

`--train` sweeps the TRAINING step of one layer instead (forward + backward + SGD through the public module API,
whichever engine `engine="auto"` picks) and runs on 1..8 GPUs:
    python scripts/sweep.py --train [--quick]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/sweep.py --train --quick
With N ranks every rank steps its own batch of Q samples (weak scaling) and the weight / bias gradients are averaged and
applied by the fused peer-memory optimizer (tgcn_b200.parallel.PeerAllreduceSGD); the time is the max over ranks.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tgcn_b200 import _lib, workloads as wl  # noqa: E402
from tgcn_b200.csr import build_csr  # noqa: E402


def time_graph(fn, flush, reps=5):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    tot = 0.0
    for r in range(reps):
        flush.fill_(float(r))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps * 1e3          # microseconds


STEPS_PER_REPLAY = 8


def train_sweep(quick):
    """One layer's training step over the (graph, Q, K, T) grid; CSV on rank 0."""
    import torch.distributed as dist
    from tgcn_b200.nn.gcn import TGCNCheb_H
    from tgcn_b200.parallel import PeerAllreduceSGD
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    graphs = {"hcp360": wl.hcp_parcellation()[2][0], "mesh32k": wl.cortical_mesh()[2][0]}
    Ks = [3, 10, 25] if quick else [3, 5, 8, 10, 15, 20, 25]
    Ts = [5, 30] if quick else [5, 10, 15, 30, 60]
    if rank == 0:
        print("graph,N,gpus,Q_per_gpu,K,T,G,engine,step_us,samples_per_s")
    for gname, L in graphs.items():
        N = L.shape[0]
        for Q in ((8, 64) if gname == "hcp360" else (8,)):
            for K in Ks:
                for T in Ts:
                    G = 32
                    if 4.0 * N * Q * T * K > 12e9:
                        continue
                    torch.manual_seed(1)                      # identical replicas
                    layer = TGCNCheb_H(L, 1, G, K, T).to(dev)
                    opt = PeerAllreduceSGD(list(layer.parameters()), lr=0.01, momentum=0.5)
                    torch.manual_seed(100 + rank)
                    x = torch.randn(Q, N, T, device=dev)
                    dout = torch.randn(Q, N, G, device=dev)

                    def step():
                        for p_ in layer.parameters():
                            p_.grad = None
                        layer(x).backward(dout)
                        opt.step()

                    def steps():              # several steps per replay: the ranks' launch skew is paid once, not per step
                        for _ in range(STEPS_PER_REPLAY):
                            step()
                    if world > 1:
                        dist.barrier()
                    us = time_graph(steps, flush) / STEPS_PER_REPLAY
                    t = torch.tensor([us], device=dev)
                    if world > 1:
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    if rank == 0:
                        print("%s,%d,%d,%d,%d,%d,%d,%s,%.1f,%.0f" % (gname, N, world, Q, K, T, G, layer.last_engine if hasattr(layer, "last_engine") else "auto",
                                                                    float(t), world * Q / float(t) * 1e6), flush=True)
                    del opt, layer
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)            # see DESIGN.md section 6 (teardown with captured graphs)


def main():
    if "--train" in sys.argv:
        return train_sweep("--quick" in sys.argv)
    lib = _lib.load()
    dev = torch.device("cuda")
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    graphs = {"hcp360": wl.hcp_parcellation()[2][0], "mesh32k": wl.cortical_mesh()[2][0]}
    quick = "--quick" in sys.argv
    Ks = [3, 10, 25] if quick else [3, 5, 8, 10, 15, 20, 25]
    Ts = [5, 30] if quick else [5, 10, 15, 30, 60]
    print("graph,N,nnz,Q,K,T,F,G,basis_us,contract_us,resident_us,basis_GBps_alg,contract_TFLOPs,dominant")
    for gname, L in graphs.items():
        plan = build_csr(L, dev)
        N = plan.n
        for (F, G) in ((1, 32), (32, 64)):
            for Q in ((1, 8, 64) if gname == "hcp360" else (1, 8)):
                for K in Ks:
                    for T in (Ts if F == 1 else [1]):
                        D = T * F
                        C = Q * D
                        if 4.0 * N * C * K > 20e9:
                            continue
                        x = torch.randn(Q, N, D, device=dev)
                        W = torch.randn(K, D, G, device=dev) * 0.1
                        bias = torch.randn(N, G, device=dev)
                        stack = torch.empty(K, N, C, device=dev)
                        out = torch.empty(Q, N, G, device=dev)
                        scr = torch.empty(max(int(lib.tgcn_contract_fwd_scratch(Q, N, D, G, K)), 16) // 4 + 64, device=dev)

                        def basis():
                            st = torch.cuda.current_stream().cuda_stream
                            assert lib.tgcn_cheb_basis(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N, x.data_ptr(),
                                                       stack.data_ptr(), Q, D, K, 0, st) == 0, _lib.last_error()

                        def contract():
                            st = torch.cuda.current_stream().cuda_stream
                            assert lib.tgcn_contract_fwd(stack.data_ptr(), W.data_ptr(), bias.data_ptr(), 1, out.data_ptr(), scr.data_ptr(),
                                                         Q, N, D, G, K, 0, st) == 0, _lib.last_error()
                        tb, tc = time_graph(basis, flush), time_graph(contract, flush)
                        tr = float("nan")
                        if lib.tgcn_resident_supported(N, D, G, K, plan.nnz):
                            rowinfo, entries, E = plan.packed(lib.tgcn_resident_pack_classes(Q, N, D, 0))
                            wimg = torch.empty(int(lib.tgcn_resident_weights_bytes(D, G, K)) // 4, device=dev)

                            def resident():
                                st = torch.cuda.current_stream().cuda_stream
                                assert lib.tgcn_resident_layer_fwd(rowinfo.data_ptr(), entries.data_ptr(), N, E, x.data_ptr(), W.data_ptr(),
                                                                   bias.data_ptr(), 1, out.data_ptr(), None, None, 0, 0, None, None, wimg.data_ptr(),
                                                                   Q, D, G, K, 0, st) == 0, _lib.last_error()
                            tr = time_graph(resident, flush)
                        S = 4.0 * N * C
                        E_ = 8.0 * plan.nnz + 4.0 * (N + 1)
                        gb = ((3 * K - 4) * S + (K - 1) * E_) / tb / 1e3
                        tf = 2.0 * Q * N * K * D * G / tc / 1e6
                        print("%s,%d,%d,%d,%d,%d,%d,%d,%.1f,%.1f,%.1f,%.0f,%.2f,%s" % (gname, N, plan.nnz, Q, K, T, F, G, tb, tc, tr, gb, tf,
                                                                                  "spmm" if tb > tc else "contraction"), flush=True)


if __name__ == "__main__":
    main()
