"""Generate the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  The tests never import the reference; they read these
files.  Every array is the direct output of reference code:

  graph_*.npz   gcn/graph.py grid/distance_sklearn_metrics/adjacency + gcn/coarsening.py
                coarsen (perm, parents, per-level graphs) + laplacian/rescale_L per level
  layer_*.npz   tgcn/nn/gcn.py TGCNCheb_H / GCNCheb / TGCNCheb forward + autograd backward
  pool.npz      gcn_pool / gcn_pool_4 values, argmax indices and gradient routing
  init.npz      parameter values after torch.manual_seed(s) (RNG draw order)
  perm_data.npz coarsening.perm_data / perm_data_time
"""
import contextlib
import io
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

G, C, NN, MM = ref_loader.load()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def csr_parts(prefix, A):
    A = A.tocsr()
    A.sort_indices()
    return {prefix + "_indptr": A.indptr.astype(np.int64), prefix + "_indices": A.indices.astype(np.int64),
            prefix + "_data": A.data, prefix + "_shape": np.array(A.shape, dtype=np.int64)}


def graph_case(name, A, levels, seed):
    out = {"levels": np.int64(levels), "seed": np.int64(seed)}
    out.update(csr_parts("A", A))
    np.random.seed(seed)
    graphs, parents = C.metis(A, levels)
    for i, p in enumerate(parents):
        out["parents_%d" % i] = np.asarray(p)
    for i, g in enumerate(graphs):
        out.update(csr_parts("metis_graph_%d" % i, g))
    perms = C.compute_perm(parents)
    for i, p in enumerate(perms):
        out["perms_%d" % i] = np.asarray(p, dtype=np.int64)
    np.random.seed(seed)
    cgraphs, perm = quiet(C.coarsen, A, levels=levels, self_connections=False)
    out["perm"] = np.asarray(perm, dtype=np.int64)
    for i, g in enumerate(cgraphs):
        out.update(csr_parts("graph_%d" % i, g))
        L = G.rescale_L(G.laplacian(g, normalized=True), lmax=2)
        out.update(csr_parts("L_%d" % i, L))
    np.savez_compressed(os.path.join(HERE, "graph_%s.npz" % name), **out)
    print(name, [g.shape[0] for g in cgraphs], "nnz0", cgraphs[0].nnz)
    return cgraphs, perm


def grid_graph(m, k):
    z = G.grid(m)
    d, i = G.distance_sklearn_metrics(z, k=k, metric="euclidean")
    return G.adjacency(d, i), z, d, i


def random_sparse_graph(n, deg, seed):
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(n), deg)
    cols = rng.integers(0, n, size=n * deg)
    vals = rng.random(n * deg).astype(np.float32) + np.float32(0.05)
    keep = rows != cols
    W = sp.coo_matrix((vals[keep], (rows[keep], cols[keep])), shape=(n, n)).tocsr()
    W = W.maximum(W.T).tocsr()
    # connect a ring so that no vertex is isolated
    ring = sp.coo_matrix((np.full(n, 0.5, np.float32), (np.arange(n), (np.arange(n) + 1) % n)), shape=(n, n))
    W = W.maximum(ring).maximum(ring.T).tocsr().astype(np.float32)
    W.setdiag(0)
    W.eliminate_zeros()
    return W


def dense_connectome(n, seed):
    rng = np.random.default_rng(seed)
    M = rng.lognormal(0.0, 1.0, size=(n, n)).astype(np.float32)
    M = np.maximum(M, M.T)
    np.fill_diagonal(M, 0)
    return sp.csr_matrix(M)


def layer_case(name, cls_name, L_sp, x, in_ch, out_ch, K, H=None, bias=True, seed=0, module="gcn"):
    mod = NN if module == "gcn" else MM
    Ld = torch.tensor(np.asarray(L_sp.todense()), dtype=torch.float)
    torch.manual_seed(seed)
    cls = getattr(mod, cls_name)
    layer = cls(Ld, in_ch, out_ch, K, H, bias=bias) if cls_name == "TGCNCheb_H" else cls(Ld, in_ch, out_ch, K, bias=bias)
    xt = torch.tensor(x, dtype=torch.float, requires_grad=True)
    out = layer(xt)
    gen = torch.Generator().manual_seed(seed + 1000)
    dout = torch.randn(out.shape, generator=gen)
    out.backward(dout)
    basis = (layer._time_chebyshev(xt) if hasattr(layer, "_time_chebyshev") else layer._chebyshev(xt)).detach()
    rec = {"x": x.astype(np.float32), "W": layer.weight.detach().numpy(), "out": out.detach().numpy(),
           "dout": dout.numpy(), "dW": layer.weight.grad.numpy(), "dx": xt.grad.numpy(),
           "basis": basis.numpy(), "K": np.int64(K), "in_ch": np.int64(in_ch), "out_ch": np.int64(out_ch),
           "H": np.int64(-1 if H is None else H), "cls": np.array(cls_name), "seed": np.int64(seed)}
    if bias:
        rec["b"] = layer.bias.detach().numpy()
        rec["db"] = layer.bias.grad.numpy()
    rec.update(csr_parts("L", L_sp))
    np.savez_compressed(os.path.join(HERE, "layer_%s.npz" % name), **rec)
    print("layer", name, tuple(out.shape), float(out.abs().max()))


def main():
    # ---------------- graphs + coarsening ----------------
    A8, z, d8, i8 = grid_graph(28, 8)
    np.savez_compressed(os.path.join(HERE, "knn_grid28_k8.npz"), z=z, dist=d8, idx=i8.astype(np.int64),
                        **csr_parts("A", A8))
    g_s0, perm_s0 = graph_case("grid28_k8_seed0", A8, 4, 0)
    graph_case("grid28_k8_seed1", A8, 4, 1)
    A12, _, _, _ = grid_graph(28, 12)
    graph_case("grid28_k12_seed0", A12, 4, 0)
    g_rs, perm_rs = graph_case("rand300_seed2", random_sparse_graph(300, 5, 7), 3, 2)
    g_dc, perm_dc = graph_case("dense40_seed3", dense_connectome(40, 11), 2, 3)

    # ---------------- layers ----------------
    def Lof(g):
        return G.rescale_L(G.laplacian(g, normalized=True), lmax=2)

    rng = np.random.default_rng(123)
    L0 = Lof(g_s0[0])      # N=992
    L2 = Lof(g_s0[2])      # N=248
    L4 = Lof(g_s0[4])      # N=62
    Lr = Lof(g_rs[0])      # random graph, padded
    Ld = Lof(g_dc[0])      # dense connectome, padded

    def data(Q, N, *rest, perm=None, n_real=None):
        x = rng.standard_normal((Q, N) + rest).astype(np.float32)
        return x

    # config-1 shape: TGCNCheb_H(L0, 1, 15, K=10, H=12), x [Q,N,H]
    layer_case("tgcnh_c1_q3", "TGCNCheb_H", L0, data(3, 992, 12), 1, 15, 10, 12)
    # 4-D input with F>1, matmul module (bit-identical maths, gcn_matmul.py)
    layer_case("tgcnh_f3_matmul", "TGCNCheb_H", L2, data(2, 248, 5, 3), 3, 7, 6, 5, module="gcn_matmul")
    layer_case("tgcnh_k1", "TGCNCheb_H", L4, data(2, 62, 4), 1, 5, 1, 4)
    layer_case("tgcnh_k2_nobias", "TGCNCheb_H", L4, data(2, 62, 4, 2), 2, 5, 2, 4, bias=False)
    layer_case("tgcnh_k3", "TGCNCheb_H", L4, data(1, 62, 3), 1, 4, 3, 3)
    layer_case("tgcnh_k25_g32", "TGCNCheb_H", L2, data(2, 248, 12), 1, 32, 25, 12)
    layer_case("tgcnh_rand", "TGCNCheb_H", Lr, data(3, Lr.shape[0], 15), 1, 32, 10, 15)
    layer_case("tgcnh_dense", "TGCNCheb_H", Ld, data(4, Ld.shape[0], 15), 1, 32, 10, 15)
    # second-layer shapes: GCNCheb(L2, 32, 64, 10) on [Q,N,F]; and the 2-D input rule
    layer_case("gcn_f32_g64", "GCNCheb", L2, data(3, 248, 32), 32, 64, 10)
    layer_case("gcn_2d_input", "GCNCheb", L4, data(5, 62), 1, 6, 4)
    layer_case("gcn_k1", "GCNCheb", L4, data(2, 62, 3), 3, 2, 1)
    # TGCNCheb (per-vertex bias, no horizon)
    layer_case("tgcn_f4", "TGCNCheb", L4, data(3, 62, 4), 4, 9, 5)

    # ---------------- pooling ----------------
    torch.manual_seed(5)
    xp = torch.randn(3, 48, 7)
    xp = torch.relu(xp)                       # lots of exact-zero ties
    xp[0, 4:8, 0] = float("nan")
    xp[1, 9, 1] = float("nan")
    xp[2, 12:16, 2] = 1.5                     # 4-way tie
    xp[2, 17, 3] = float("inf")
    xp[2, 20:24, 4] = float("-inf")
    rec = {"x": xp.numpy()}
    for p, fn in ((2, NN.gcn_pool), (4, NN.gcn_pool_4)):
        xr = xp.clone().requires_grad_(True)
        y = fn(xr)
        vals, idx = torch.max(xr.reshape(3, 48 // p, p, 7), dim=2)
        assert torch.equal(torch.nan_to_num(vals, nan=-7.0), torch.nan_to_num(y, nan=-7.0))
        gen = torch.Generator().manual_seed(p)
        dy = torch.randn(y.shape, generator=gen)
        y.backward(dy)
        rec.update({"y%d" % p: y.detach().numpy(), "idx%d" % p: idx.numpy(), "dy%d" % p: dy.numpy(),
                    "dx%d" % p: xr.grad.numpy()})
    np.savez_compressed(os.path.join(HERE, "pool.npz"), **rec)

    # ---------------- init (RNG draw order) ----------------
    rec = {}
    Lt = torch.tensor(np.asarray(L4.todense()), dtype=torch.float)
    torch.manual_seed(42)
    a = NN.TGCNCheb_H(Lt, 2, 5, 3, 4)
    b = NN.GCNCheb(Lt, 5, 6, 4)
    c = NN.TGCNCheb(Lt, 3, 2, 5)
    d = NN.TGCNCheb_H(Lt, 1, 3, 2, 4, bias=False)
    e = NN.GCNCheb(Lt, 2, 2, 2)
    for nm, lay in (("a", a), ("b", b), ("c", c), ("d", d), ("e", e)):
        rec[nm + "_weight"] = lay.weight.detach().numpy()
        if lay.bias is not None:
            rec[nm + "_bias"] = lay.bias.detach().numpy()
        rec[nm + "_repr"] = np.array(repr(lay))
    np.savez_compressed(os.path.join(HERE, "init.npz"), **rec)

    # ---------------- perm_data ----------------
    perm = list(perm_s0)
    x2 = rng.standard_normal((3, 784)).astype(np.float32)
    x3 = rng.standard_normal((2, 784, 4)).astype(np.float32)
    sys.path.insert(0, os.path.join(ref_loader.REF_ROOT))
    # perm_data_time lives in the example scripts (not importable: torchvision/autograd deps);
    # its body is the 3-D twin of coarsening.perm_data, so the golden applies perm_data per time slice.
    y2 = C.perm_data(x2, perm)
    y3 = np.stack([C.perm_data(x3[:, :, t], perm) for t in range(x3.shape[2])], axis=2)
    np.savez_compressed(os.path.join(HERE, "perm_data.npz"), perm=np.asarray(perm, dtype=np.int64), x2=x2, y2=y2,
                        x3=x3, y3=y3)
    # KAT from the reference itself (coarsening.py:216-217) is restated verbatim in the tests.
    print("done")


if __name__ == "__main__":
    main()
