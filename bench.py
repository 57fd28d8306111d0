#!/usr/bin/env python
"""bench.py -- TGCN train samples/s on B200 (+ Chebyshev SpMM roofline, + CPU reference arm).

Contract (see DESIGN.md section "Measurement"):
    python bench.py --gpus N --steps K --warmup W [--workload hcp360|mesh32k|mnist] [--impl reference]
One "step" = forward + loss + backward (+ gradient allreduce for N > 1) + SGD update of the
reference's model for the workload on one batch of synthetic data.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

METRIC = "tgcn_train_samples_per_s"
UNIT = "samples/s"

WORKLOADS = {
    # name: (description, per-GPU batch, horizon, graph builder kwargs)
    "hcp360": dict(desc="BASELINE.json configs[1]: HCP-shaped parcellation graph (360 regions, top-16 sparse "
                        "connectome, 4 coarsening levels), T=15, batch 64 per GPU, NetTGCN_HCP "
                        "(TGCNCheb_H(1,32,K10,H15)->relu->pool4->GCNCheb(32,64,K10)->relu->pool4->fc200->bn->fc6)",
                   batch=64, H=15, model="hcp", classes=6),
    "hcp360dense": dict(desc="configs[1] with the fully dense connectome variant", batch=64, H=15, model="hcp",
                        classes=6),
    "mesh32k": dict(desc="BASELINE.json configs[2]: cortical-surface mesh (32492 vertices -> 41856 padded), T=30, "
                         "batch 8 per GPU, two TGCN layers + pooling", batch=8, H=30, model="hcp", classes=6),
    "mnist": dict(desc="BASELINE.json configs[0]: 28x28 8-NN grid, 4 coarsening levels, batch 100, H=12, "
                       "TGCNCheb_H(1,15,K10)->relu->fc10", batch=100, H=12, model="mnist", classes=10),
}


def build_graph(name):
    from tgcn_b200 import workloads as wl
    if name == "hcp360":
        return wl.hcp_parcellation(dense=False)
    if name == "hcp360dense":
        return wl.hcp_parcellation(dense=True)
    if name == "mesh32k":
        return wl.cortical_mesh()
    if name == "mnist":
        return wl.mnist_grid()
    raise ValueError(name)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def algorithmic_step_bytes(N, C, nnz, has_prev):
    S = 4 * N * C
    E = 8 * nnz + 4 * (N + 1)
    return (3 if has_prev else 2) * S + E


def run_b200(args):
    import torch.distributed as dist
    from tgcn_b200 import _lib, workloads as wl
    from tgcn_b200.parallel import FlatGradients, broadcast_parameters, init_distributed

    rank, world, local = init_distributed("nccl")
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torchrun for N>1)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    if lib.tgcn_device_supported() != 1:
        raise SystemExit("libtgcn_b200 needs an sm_100 device")
    cfg = WORKLOADS[args.workload]
    Q = args.batch or cfg["batch"]
    H = cfg["H"]

    graphs, perm, Ls, n_real = build_graph(args.workload)
    Lt = wl.as_torch_operands(Ls, device=dev)
    torch.manual_seed(0)
    if cfg["model"] == "hcp":
        model = wl.NetTGCN_HCP(Lt, horizon=H, n_classes=cfg["classes"], engine=args.engine).to(dev)
    else:
        model = wl.NetTGCN_MNIST(Lt, horizon=H, n_classes=cfg["classes"], engine=args.engine).to(dev)
    broadcast_parameters(model)
    N0 = Ls[0].shape[0]
    grads = FlatGradients(model.parameters())
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.5)   # pytorch_hcp_tgcn.py defaults

    # synthetic data: a few distinct pinned host batches per rank, cycled
    n_host = 4
    hx = [wl.synthetic_signals(Q, N0, H, n_real, perm, seed=1000 * rank + i).pin_memory() for i in range(n_host)]
    g = torch.Generator().manual_seed(77 + rank)
    hy = [torch.randint(0, cfg["classes"], (Q,), generator=g).pin_memory() for _ in range(n_host)]
    x = hx[0].to(dev)
    y = hy[0].to(dev)
    loss_dev = torch.zeros((), device=dev)
    loss_host = torch.zeros((), pin_memory=True)

    def fwd_bwd():
        grads.zero_()
        out = model(x)
        loss = F.nll_loss(out, y)
        loss.backward()
        loss_dev.copy_(loss.detach())

    def finish_step():
        grads.allreduce_mean()
        opt.step()

    # warm-up (eager, side stream) then capture the forward/backward and the update as two graphs
    model.train()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fwd_bwd(); finish_step()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    use_graph = not args.no_graph
    launches_per_step = None
    if use_graph:
        g1 = torch.cuda.CUDAGraph()
        c0 = lib.tgcn_launch_count()
        with torch.cuda.graph(g1):
            fwd_bwd()
        launches_per_step = lib.tgcn_launch_count() - c0
        g2 = torch.cuda.CUDAGraph()
        if world == 1:
            with torch.cuda.graph(g2):
                opt.step()

        def step():
            g1.replay()
            if world == 1:
                g2.replay()
            else:
                finish_step()
    else:
        def step():
            fwd_bwd(); finish_step()
        c0 = lib.tgcn_launch_count(); step(); launches_per_step = lib.tgcn_launch_count() - c0

    # L2 hygiene: flush between timed iterations when the step's working set fits in L2
    D1 = H
    work_bytes = 4 * (10 * N0 * Q * D1) * 2
    flush = work_bytes < 256e6
    flush_buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev) if flush else None   # 256 MB

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, e2e):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
        barrier()
        t0 = time.perf_counter()
        for i in range(nsteps):
            if flush:
                flush_buf.fill_(float(i))
            ev[i][0].record()
            if e2e:
                x.copy_(hx[i % n_host], non_blocking=True)
                y.copy_(hy[i % n_host], non_blocking=True)
            step()
            if e2e:
                loss_host.copy_(loss_dev, non_blocking=True)
            ev[i][1].record()
            if e2e:
                ev[i][1].synchronize()         # the user reads the loss every step
                _ = float(loss_host)
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), wall

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, wall = timed(args.steps, e2e=False)
    ms_e2e, wall_e2e = timed(args.steps, e2e=True)
    # back-to-back (no flush, no per-step events) for reference
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step()
    b.record(); torch.cuda.synchronize()
    ms_warm = a.elapsed_time(b)
    clocks = sampler.stop() if rank == 0 else None

    roof = measure_spmm_roofline(lib, model, Ls, Q, H, dev, flush_buf) if rank == 0 else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_port_throughput(args.workload, Ls, perm, n_real, cfg, budget_s=args.cpu_budget)

    if rank == 0:
        ms_step = ms_total / args.steps
        line = {
            "metric": METRIC, "value": world * Q / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": cfg["desc"], "per_gpu_batch": Q, "global_batch": world * Q,
                       "N_padded": [int(L.shape[0]) for L in Ls], "nnz_L0": int(Ls[0].nnz), "K": 10, "H": H,
                       "parallelism": "dp%d" % world, "engine": args.engine,
                       "l2": "flushed between timed steps (256 MB write)" if flush else "working set exceeds L2 (K-slab stack > 126 MB)",
                       "cuda_graph": use_graph},
            "e2e": {"value": world * Q / (ms_e2e / args.steps * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(hx[0].numel() * 4 + hy[0].numel() * 8), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "ms_per_step_back_to_back": ms_warm / args.steps,
            "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step),
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "final_loss": float(loss_host),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_spmm_roofline(lib, model, Ls, Q, H, dev, flush_buf):
    """Average launch duration of the dominant kernel (the CSR SpMM recursion step of layer 1) on the
    workload's own operand and slab shapes, CUDA events on the launching stream."""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    plan = model.tgcn1._plan(dev)
    N = plan.n
    C = Q * H
    K = 10
    stack = torch.randn(K, N, C, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    reps = 20

    def steps():
        st = torch.cuda.current_stream().cuda_stream
        for k in range(1, K):
            lib.tgcn_spmm_step(plan.rowptr.data_ptr(), plan.col.data_ptr(), plan.val.data_ptr(), N,
                               stack[k - 1].data_ptr(), None, stack[k].data_ptr(), C, 1.0, 0.0, st)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            steps()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    # the K-1 launches are replayed from a CUDA graph so that host launch latency (ctypes) is not
    # what gets timed on the small workloads
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        st = torch.cuda.current_stream().cuda_stream
        steps()
    total = 0.0
    for r in range(reps):
        if flush_buf is not None:
            flush_buf.fill_(float(r))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); graph.replay(); b.record(); b.synchronize()
        total += a.elapsed_time(b)
    per_launch_ms = total / (reps * (K - 1))
    bytes_per_launch = algorithmic_step_bytes(N, C, plan.nnz, has_prev=False)
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "spmm_step_vec4_kernel (layer-1 recursion step, 2S+E bytes)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "bytes_per_launch": int(bytes_per_launch), "us_per_launch": per_launch_ms * 1e3, "peak_source": src}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port = the reference's ATen call sequence)
# ------------------------------------------------------------------------------------------------
def cpu_port_throughput(workload, Ls, perm, n_real, cfg, budget_s=20.0, steps=None, warmup=1):
    from oracle import model_torch
    from tgcn_b200 import workloads as wl
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    N0 = Ls[0].shape[0]
    dense_ok = N0 <= 4096
    Lt = wl.as_torch_operands(Ls, device=None, dense=dense_ok)
    if not dense_ok:
        Lt = [t.to_sparse_csr() for t in Lt]      # gcn_matmul.py slab path (dense L~ is infeasible at this N)
    torch.manual_seed(0)
    H = cfg["H"]
    Q = cfg["batch"]
    sample_q = Q if dense_ok else max(1, min(Q, 2))
    if cfg["model"] == "hcp":
        model = model_torch.PortNetTGCN_HCP(Lt, horizon=H, n_classes=cfg["classes"])
    else:
        model = model_torch.PortNetTGCN_MNIST(Lt, horizon=H, n_classes=cfg["classes"])
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.5)
    x = wl.synthetic_signals(sample_q, N0, H, n_real, perm, seed=5)
    y = torch.randint(0, cfg["classes"], (sample_q,))
    model.train()

    def one():
        opt.zero_grad()
        loss = F.nll_loss(model(x), y)
        loss.backward()
        opt.step()
        return float(loss.detach())
    for _ in range(warmup):
        one()
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter(); one(); times.append(time.perf_counter() - t0)
        if steps is not None and len(times) >= steps:
            break
        if steps is None and (time.perf_counter() - t_start > budget_s or len(times) >= 50):
            break
    med = float(np.median(times))
    return {"value": sample_q / med, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d timed train steps (fwd+loss+bwd+SGD) of the torch-CPU port of the reference model, batch %d of %d, "
                      "%s L~, %d threads, median %.1f ms/step" % (len(times), sample_q, Q,
                                                                   "dense" if dense_ok else "torch.sparse_csr (gcn_matmul slab path)",
                                                                   cores, med * 1e3),
            "ms_per_step": med * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = WORKLOADS[args.workload]
    graphs, perm, Ls, n_real = build_graph(args.workload)
    cpu = cpu_port_throughput(args.workload, Ls, perm, n_real, cfg, steps=args.steps, warmup=max(1, min(args.warmup, 3)))
    line = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cpu["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": cfg["desc"], "per_gpu_batch": cfg["batch"]},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hcp360", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--engine", default="auto", choices=["auto", "ffma", "tcgen05"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
